// Segmentation / adversarial losses and the IoU / Dice metrics.
// One streaming pass over (logits, target) yields every partial sum BCEDiceLoss + MSELoss need; the
// backward is one more pass.  Metrics: integer counts for IoU; for Dice the float32 sums follow
// NumPy's pairwise-summation tree exactly so the result is bit-identical to metrics.py on identical
// probabilities.
#include "common.cuh"
#include <vector>

namespace ssg {

// sums[b][5] = {sum bce_elem, sum p*t, sum p, sum t, sum (x-t)^2}
__global__ void __launch_bounds__(256) seg_loss_sums_kernel(const float* __restrict__ x, const float* __restrict__ t, long long per_sample,
                                                             double* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* xb = x + (long long)b * per_sample;
    const float* tb = t + (long long)b * per_sample;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += stride) {
        const float xv = xb[i], tv = tb[i];
        // losses.py:135: input.clamp(min=0) - input*target + log(1 + exp(-|input|))
        a0 += fmaxf(xv, 0.f) - xv * tv + logf(1.f + expf(-fabsf(xv)));
        const float p = sigmoidf_(xv);
        a1 = fmaf(p, tv, a1);
        a2 += p;
        a3 += tv;
        const float d = xv - tv;
        a4 = fmaf(d, d, a4);
    }
    __shared__ float red[5][8];
    float v[5] = {a0, a1, a2, a3, a4};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) red[k][wid] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += (double)red[threadIdx.x][w];
        atomicAdd(&sums[b * 5 + threadIdx.x], s);
    }
}

__global__ void seg_loss_finalize_kernel(const double* __restrict__ sums, int batch, long long per_sample, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float smooth = 1e-5f;
    double bce_sum = 0, mse_sum = 0;
    float dice_acc = 0.f;
    for (int b = 0; b < batch; ++b) {
        bce_sum += sums[b * 5 + 0];
        mse_sum += sums[b * 5 + 4];
        const float inter = (float)sums[b * 5 + 1], ps = (float)sums[b * 5 + 2], ts = (float)sums[b * 5 + 3];
        dice_acc += (2.f * inter + smooth) / (ps + ts + smooth);    // losses.py:291
    }
    const double total = (double)batch * (double)per_sample;
    const float bce = (float)(bce_sum / total);
    const float dice = 1.f - dice_acc / (float)batch;               // losses.py:292
    const bool bad = isinf(bce) || isnan(bce);                       // losses.py:297
    out[0] = bad ? 2.0f * dice : 0.5f * bce + dice;
    out[1] = bce;
    out[2] = dice;
    out[3] = (float)(mse_sum / total);
    out[4] = bad ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                            const double* __restrict__ sums, const float* __restrict__ out,
                                                            const float* __restrict__ g, int batch, long long per_sample,
                                                            float* __restrict__ dx) {
    const int b = blockIdx.y;
    const float smooth = 1e-5f;
    const float inter = (float)sums[b * 5 + 1], ps = (float)sums[b * 5 + 2], ts = (float)sums[b * 5 + 3];
    const float den = ps + ts + smooth, num = 2.f * inter + smooth;
    const bool bad = out[4] != 0.f;
    const float total = (float)batch * (float)per_sample;
    const float g_loss = g[0], g_mse = g[1], g_bce = g[2];
    const float k_bce = (bad ? 0.f : 0.5f * g_loss / total) + g_bce / total;
    const float k_dice = (bad ? 2.f : 1.f) * g_loss * (-1.f / (float)batch);
    const float k_mse = 2.f * g_mse / total;
    const float inv_den2 = 1.f / (den * den);
    const float* xb = x + (long long)b * per_sample;
    const float* tb = t + (long long)b * per_sample;
    float* db = dx + (long long)b * per_sample;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += stride) {
        const float xv = xb[i], tv = tb[i];
        const float p = sigmoidf_(xv);
        // d dice_n / d p_i = (2 t_i den - num) / den^2 ;  d p / d x = p (1 - p)
        const float ddice = (2.f * tv * den - num) * inv_den2 * p * (1.f - p);
        db[i] = k_bce * (p - tv) + k_dice * ddice + k_mse * (xv - tv);
    }
}

// tiny (n = batch) BCEWithLogits against a constant target
__global__ void bce_logits_fwd_kernel(const float* __restrict__ x, float tv, int n, float* __restrict__ out) {
    __shared__ float red[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float xv = x[i];
        // ATen binary_cross_entropy_with_logits: (1-t)*x + max(-x,0) + log(exp(-max)+exp(-x-max))  ==  stable softplus form
        a += (1.f - tv) * xv + (fmaxf(-xv, 0.f) + logf(expf(-fmaxf(-xv, 0.f)) + expf(-xv - fmaxf(-xv, 0.f))));
    }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        out[0] = s / (float)n;
    }
}
__global__ void bce_logits_bwd_kernel(const float* __restrict__ x, float tv, int n, const float* __restrict__ g, float* __restrict__ dx) {
    const float gg = g[0] / (float)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dx[i] = gg * (sigmoidf_(x[i]) - tv);
}

__global__ void __launch_bounds__(256) iou_counts_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n,
                                                          unsigned long long* __restrict__ counts) {
    unsigned int inter = 0, uni = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float p = sigmoidf_(x[i]);
        const bool o = p > 0.5f;              // NaN -> false, metrics.py:15
        const bool tt = t[i] > 0.5f;
        inter += (o && tt);
        uni += (o || tt);
    }
    for (int o = 16; o > 0; o >>= 1) {
        inter += __shfl_xor_sync(0xffffffffu, inter, o);
        uni += __shfl_xor_sync(0xffffffffu, uni, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&counts[0], (unsigned long long)inter);
        atomicAdd(&counts[1], (unsigned long long)uni);
    }
}

// NumPy pairwise sum leaf (n <= 128): 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
// then the tail added sequentially; n < 8: plain sequential sum starting from 0.
// __fadd_rn keeps the compiler from contracting / reassociating.
struct Leaf3 {
    float r[3][8];
    float res[3];
};
__global__ void __launch_bounds__(128) dice_leaf_sums_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                              const long long* __restrict__ offs, long long n_leaves,
                                                              float* __restrict__ leaf, float* __restrict__ probs) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_leaves) return;
    const long long lo = offs[li];
    const int n = (int)(offs[li + 1] - lo);
    float res[3];
    auto vals = [&](long long i, float* v) {
        const float p = sigmoidf_(x[i]);
        const float tv = t[i];
        if (probs) probs[i] = p;
        v[0] = __fmul_rn(p, tv);
        v[1] = p;
        v[2] = tv;
    };
    if (n < 8) {
        res[0] = res[1] = res[2] = 0.f;
        for (int i = 0; i < n; ++i) {
            float v[3]; vals(lo + i, v);
#pragma unroll
            for (int k = 0; k < 3; ++k) res[k] = __fadd_rn(res[k], v[k]);
        }
    } else {
        float r[3][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[3]; vals(lo + j, v);
#pragma unroll
            for (int k = 0; k < 3; ++k) r[k][j] = v[k];
        }
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v[3]; vals(lo + i + j, v);
#pragma unroll
                for (int k = 0; k < 3; ++k) r[k][j] = __fadd_rn(r[k][j], v[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            res[k] = __fadd_rn(__fadd_rn(__fadd_rn(r[k][0], r[k][1]), __fadd_rn(r[k][2], r[k][3])),
                               __fadd_rn(__fadd_rn(r[k][4], r[k][5]), __fadd_rn(r[k][6], r[k][7])));
        for (; i < n; ++i) {
            float v[3]; vals(lo + i, v);
#pragma unroll
            for (int k = 0; k < 3; ++k) res[k] = __fadd_rn(res[k], v[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) leaf[k * n_leaves + li] = res[k];
}

static void leaves_rec(long long lo, long long n, std::vector<long long>& out) {
    if (n <= 128) { out.push_back(lo); return; }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    leaves_rec(lo, n2, out);
    leaves_rec(lo + n2, n - n2, out);
}
static float combine_rec(const float*& leaf, long long n) {
    if (n <= 128) return *leaf++;
    long long n2 = n / 2;
    n2 -= n2 % 8;
    volatile float a = combine_rec(leaf, n2);
    volatile float b = combine_rec(leaf, n - n2);
    volatile float r = a + b;
    return r;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_seg_loss_sums(const float* logits, const float* target, int batch, long long per_sample, double* sums, ssg_stream_t s) {
    SSG_CHECK_ARG(batch > 0 && batch <= 65535 && per_sample > 0, "seg_loss: bad shape");
    SSG_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 5 * batch, (cudaStream_t)s));
    long long bx = (per_sample + 256 * 8 - 1) / (256 * 8);
    long long cap = (long long)sm_count_cached() * 8 / batch + 1;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)batch);
    seg_loss_sums_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(logits, target, per_sample, sums);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_seg_loss_finalize(const double* sums, int batch, long long per_sample, float* out, ssg_stream_t s) {
    seg_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)s>>>(sums, batch, per_sample, out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_seg_loss_bwd(const float* logits, const float* target, const double* sums, const float* out, const float* g_loss, int batch,
                     long long per_sample, float* dlogits, ssg_stream_t s) {
    SSG_CHECK_ARG(batch > 0 && batch <= 65535 && per_sample > 0, "seg_loss: bad shape");
    long long bx = (per_sample + 256 * 4 - 1) / (256 * 4);
    long long cap = (long long)sm_count_cached() * 8 / batch + 1;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)batch);
    seg_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(logits, target, sums, out, g_loss, batch, per_sample, dlogits);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_bce_logits_fwd(const float* x, float target_value, int n, float* out, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0, "bce_logits: n");
    bce_logits_fwd_kernel<<<1, 256, 0, (cudaStream_t)s>>>(x, target_value, n, out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_bce_logits_bwd(const float* x, float target_value, int n, const float* g, float* dx, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0, "bce_logits: n");
    bce_logits_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)s>>>(x, target_value, n, g, dx);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_iou_counts(const float* logits, const float* target, long long n, long long* counts, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0, "iou_counts: n");
    SSG_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(long long), (cudaStream_t)s));
    iou_counts_kernel<<<grid_for(n, 256 * 8), 256, 0, (cudaStream_t)s>>>(logits, target, n, (unsigned long long*)counts);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
long long ssg_pairwise_leaves_host(long long n, long long* offsets_host, long long capacity) {
    if (n < 0) return -1;
    std::vector<long long> v;
    if (n > 0) leaves_rec(0, n, v);
    if (offsets_host) {
        if ((long long)v.size() + 1 > capacity) return -(long long)v.size() - 1;
        for (size_t i = 0; i < v.size(); ++i) offsets_host[i] = v[i];
        offsets_host[v.size()] = n;
    }
    return (long long)v.size();
}
int ssg_dice_leaf_sums(const float* logits, const float* target, const long long* offsets_dev, long long n_leaves, float* leaf_sums,
                       float* probs_out, ssg_stream_t s) {
    SSG_CHECK_ARG(n_leaves > 0, "dice_leaf_sums: n_leaves");
    dice_leaf_sums_kernel<<<(unsigned)((n_leaves + 127) / 128), 128, 0, (cudaStream_t)s>>>(logits, target, offsets_dev, n_leaves, leaf_sums, probs_out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
float ssg_pairwise_combine_host(const float* leaf_sums_host, long long n) {
    if (n <= 0) return 0.f;
    const float* p = leaf_sums_host;
    return combine_rec(p, n);
}

}  // extern "C"
