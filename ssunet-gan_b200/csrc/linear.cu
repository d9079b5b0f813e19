// Skinny linear layers of the discriminator head (batch rows << features): weight-streaming kernels.
// The weight (fp32 master, [nout][k]) is read once per 8-row batch chunk; activations are tiny and L2-resident.
#include "common.cuh"

namespace ssg {

constexpr int MCH = 8;

// Weight-streaming forward: a block owns JB output features and walks the whole reduction dimension in 4-element
// vectors (float4 weights, 4 x bf16 / float4 activations), MROWS batch rows at a time, so every weight element is read
// once per MROWS rows and feeds MROWS FMAs; block-level tree reduction at the end.  (The first version gave each warp one
// feature and issued one 2-byte activation load per FMA: 0.58 ms for fc1 at batch 16; the weight read alone is 12 us.)
// JB = 4 features per block: every block re-reads the whole activation matrix from L2 (590 KB for fc1 at batch 16), so with
// JB = 2 the 512 blocks moved 300 MB of activations around 75 MB of weights (0.124 ms); the k loop is unrolled by two so that
// eight 16-byte weight loads per thread are in flight.
constexpr int LIN_JB = 4, LIN_MROWS = 16, LIN_THREADS = 256;

__device__ __forceinline__ void load4f(const float* p, float* f) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load4f(const bf16* p, float* f) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}

template <typename T>
__global__ void __launch_bounds__(LIN_THREADS, 2) linear_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, T* __restrict__ y, int m, int k,
                                                                  int nout, int act, float slope, const float* __restrict__ inv_scale) {
    __shared__ float red[LIN_THREADS / 32][LIN_JB * LIN_MROWS];
    const int j0 = blockIdx.x * LIN_JB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    const bool vec = (k % 4 == 0);
    for (int m0 = 0; m0 < m; m0 += LIN_MROWS) {
        float acc[LIN_JB][LIN_MROWS];
#pragma unroll
        for (int j = 0; j < LIN_JB; ++j)
#pragma unroll
            for (int i = 0; i < LIN_MROWS; ++i) acc[j][i] = 0.f;
        if (vec) {
            auto body = [&](int kk, const float (&wv)[LIN_JB][4]) {
#pragma unroll
                for (int i = 0; i < LIN_MROWS; ++i) {
                    if (m0 + i < m) {
                        float xv[4];
                        load4f(x + (long long)(m0 + i) * k + kk, xv);
#pragma unroll
                        for (int j = 0; j < LIN_JB; ++j)
                            acc[j][i] = fmaf(wv[j][0], xv[0], fmaf(wv[j][1], xv[1], fmaf(wv[j][2], xv[2], fmaf(wv[j][3], xv[3], acc[j][i]))));
                    }
                }
            };
            auto loadw = [&](int kk, float (&wv)[LIN_JB][4]) {
#pragma unroll
                for (int j = 0; j < LIN_JB; ++j) {
                    if (j0 + j < nout) load4f(w + (long long)(j0 + j) * k + kk, wv[j]);
                    else wv[j][0] = wv[j][1] = wv[j][2] = wv[j][3] = 0.f;
                }
            };
            int kk = threadIdx.x * 4;
            for (; kk + LIN_THREADS * 4 < k; kk += LIN_THREADS * 8) {       // two k positions per trip: 2 x JB weight loads in flight
                float wa[LIN_JB][4], wb[LIN_JB][4];
                loadw(kk, wa);
                loadw(kk + LIN_THREADS * 4, wb);
                body(kk, wa);
                body(kk + LIN_THREADS * 4, wb);
            }
            for (; kk < k; kk += LIN_THREADS * 4) {
                float wa[LIN_JB][4];
                loadw(kk, wa);
                body(kk, wa);
            }
        } else {
            for (int kk = threadIdx.x; kk < k; kk += LIN_THREADS) {
#pragma unroll
                for (int j = 0; j < LIN_JB; ++j) {
                    const float wv = j0 + j < nout ? w[(long long)(j0 + j) * k + kk] : 0.f;
#pragma unroll
                    for (int i = 0; i < LIN_MROWS; ++i)
                        if (m0 + i < m) acc[j][i] = fmaf(wv, to_f(x[(long long)(m0 + i) * k + kk]), acc[j][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LIN_JB; ++j)
#pragma unroll
            for (int i = 0; i < LIN_MROWS; ++i) {
                const float v = warp_sum(acc[j][i]);
                if (lane == 0) red[warp][j * LIN_MROWS + i] = v;
            }
        __syncthreads();
        if (threadIdx.x < LIN_JB * LIN_MROWS) {
            float v = 0.f;
#pragma unroll
            for (int wi = 0; wi < LIN_THREADS / 32; ++wi) v += red[wi][threadIdx.x];
            const int j = j0 + threadIdx.x / LIN_MROWS, i = m0 + threadIdx.x % LIN_MROWS;
            if (j < nout && i < m) {
                v = v * sc + (bias ? bias[j] : 0.f);
                y[(long long)i * nout + j] = from_f<T>(apply_act(v, act, slope));
            }
        }
        __syncthreads();
    }
}

// dx32[i][kk] += sum_j dy[i][j] * w[j][kk]: a thread owns 4 consecutive input features (one float4 of every weight row it
// visits) and all MCH16 batch rows, so the weight matrix is streamed ONCE for up to 16 rows with 16-byte loads; the j range
// is split over blockIdx.y and combined with fp32 atomics into dx32 (pre-zeroed).  dy is staged in shared memory.
constexpr int MCH16 = 16, DG_JT = 64;

template <typename T>
__global__ void __launch_bounds__(256) linear_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                            float* __restrict__ dx32, int m, int k, int nout, int j_per_block,
                                                            const float* __restrict__ inv_scale) {
    __shared__ float sdy[MCH16][DG_JT];
    const int kk = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j0 = blockIdx.y * j_per_block;
    const int j1 = min(nout, j0 + j_per_block);
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    const bool live = kk < k;            // k % 4 == 0 is checked by the launcher
    for (int m0 = 0; m0 < m; m0 += MCH16) {
        float acc[MCH16][4];
#pragma unroll
        for (int i = 0; i < MCH16; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int jt = j0; jt < j1; jt += DG_JT) {
            __syncthreads();
            for (int e = threadIdx.x; e < MCH16 * DG_JT; e += blockDim.x) {
                const int i = e / DG_JT, j = jt + e % DG_JT;
                sdy[i][e % DG_JT] = (m0 + i < m && j < j1) ? to_f(dy[(long long)(m0 + i) * nout + j]) : 0.f;
            }
            __syncthreads();
            if (live) {
                const int jn = min(DG_JT, j1 - jt);
                int jj = 0;
                for (; jj + 1 < jn; jj += 2) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + (long long)(jt + jj) * k + kk);
                    const float4 w1 = *reinterpret_cast<const float4*>(w + (long long)(jt + jj + 1) * k + kk);
#pragma unroll
                    for (int i = 0; i < MCH16; ++i) {
                        const float d0 = sdy[i][jj], d1 = sdy[i][jj + 1];
                        acc[i][0] = fmaf(w0.x, d0, fmaf(w1.x, d1, acc[i][0]));
                        acc[i][1] = fmaf(w0.y, d0, fmaf(w1.y, d1, acc[i][1]));
                        acc[i][2] = fmaf(w0.z, d0, fmaf(w1.z, d1, acc[i][2]));
                        acc[i][3] = fmaf(w0.w, d0, fmaf(w1.w, d1, acc[i][3]));
                    }
                }
                for (; jj < jn; ++jj) {
                    const float4 w0 = *reinterpret_cast<const float4*>(w + (long long)(jt + jj) * k + kk);
#pragma unroll
                    for (int i = 0; i < MCH16; ++i) {
                        const float d0 = sdy[i][jj];
                        acc[i][0] = fmaf(w0.x, d0, acc[i][0]); acc[i][1] = fmaf(w0.y, d0, acc[i][1]);
                        acc[i][2] = fmaf(w0.z, d0, acc[i][2]); acc[i][3] = fmaf(w0.w, d0, acc[i][3]);
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < MCH16; ++i)
                if (m0 + i < m) {
                    float* dst = dx32 + (long long)(m0 + i) * k + kk;
                    atomicAdd(dst, acc[i][0] * sc); atomicAdd(dst + 1, acc[i][1] * sc);
                    atomicAdd(dst + 2, acc[i][2] * sc); atomicAdd(dst + 3, acc[i][3] * sc);
                }
        }
    }
}

// scalar fallback (k % 4 != 0): thread per input feature
template <typename T>
__global__ void __launch_bounds__(256) linear_dgrad_scalar_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                                   float* __restrict__ dx32, int m, int k, int nout, int j_per_block,
                                                                   const float* __restrict__ inv_scale) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * j_per_block;
    const int j1 = min(nout, j0 + j_per_block);
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    if (kk >= k) return;
    for (int m0 = 0; m0 < m; m0 += MCH) {
        float acc[MCH];
#pragma unroll
        for (int i = 0; i < MCH; ++i) acc[i] = 0.f;
        for (int j = j0; j < j1; ++j) {
            const float wv = w[(long long)j * k + kk];
#pragma unroll
            for (int i = 0; i < MCH; ++i)
                if (m0 + i < m) acc[i] = fmaf(wv, to_f(dy[(long long)(m0 + i) * nout + j]), acc[i]);
        }
#pragma unroll
        for (int i = 0; i < MCH; ++i)
            if (m0 + i < m) atomicAdd(&dx32[(long long)(m0 + i) * k + kk], acc[i] * sc);
    }
}

// dw[j][kk] = sum_i dy[i][j] * x[i][kk]: a thread owns 4 consecutive kk and WG_J output features, so every activation
// element it loads feeds WG_J weight rows (the first version re-read x once per output feature).
constexpr int WG_J = 8;

template <typename T>
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                                            float* __restrict__ dbias, int m, int k, int nout) {
    const int kk = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j0 = blockIdx.y * WG_J;
    float acc[WG_J][4], accb[WG_J];
#pragma unroll
    for (int j = 0; j < WG_J; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; accb[j] = 0.f; }
    if (kk < k) {
        for (int i = 0; i < m; ++i) {
            float xv[4];
            if (kk + 3 < k && (k & 3) == 0) load4f(x + (long long)i * k + kk, xv);       // rows are 8 / 16-byte aligned only when k % 4 == 0
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) xv[e] = kk + e < k ? to_f(x[(long long)i * k + kk + e]) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < WG_J; ++j) {
                const float d = j0 + j < nout ? to_f(dy[(long long)i * nout + j0 + j]) : 0.f;
                acc[j][0] = fmaf(d, xv[0], acc[j][0]); acc[j][1] = fmaf(d, xv[1], acc[j][1]);
                acc[j][2] = fmaf(d, xv[2], acc[j][2]); acc[j][3] = fmaf(d, xv[3], acc[j][3]);
                accb[j] += d;
            }
        }
#pragma unroll
        for (int j = 0; j < WG_J; ++j) {
            if (j0 + j >= nout) break;
            float* dst = dw + (long long)(j0 + j) * k + kk;
            if (kk + 3 < k && (((long long)(j0 + j) * k + kk) & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (kk + e < k) dst[e] = acc[j][e];
            }
            if (dbias && kk == 0) dbias[j0 + j] = accb[j];
        }
    }
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_linear_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int m, int k, int nout, int act,
                   float slope, const float* inv_scale_dev, ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0, "linear_fwd: bad shape");
    unsigned g = (unsigned)((nout + LIN_JB - 1) / LIN_JB);
    SSG_DISPATCH_DTYPE(dtype, linear_fwd_kernel<T><<<g, LIN_THREADS, 0, (cudaStream_t)s>>>((const T*)x, w, bias, (T*)y, m, k, nout, act, slope, inv_scale_dev));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_linear_dgrad(const void* dy, const float* w, float* dx, int dtype, int m, int k, int nout, const float* inv_scale_dev,
                     ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0, "linear_dgrad: bad shape");
    SSG_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)m * k, (cudaStream_t)s));
    const bool vec = k % 4 == 0;
    int kb = vec ? (k / 4 + 255) / 256 : (k + 255) / 256;
    int splits = (4 * sm_count_cached() + kb - 1) / kb;
    if (splits > nout) splits = nout;
    if (splits < 1) splits = 1;
    int jpb = (nout + splits - 1) / splits;
    dim3 grid((unsigned)kb, (unsigned)((nout + jpb - 1) / jpb));
    if (vec) {
        SSG_DISPATCH_DTYPE(dtype, linear_dgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, w, (float*)dx, m, k, nout, jpb, inv_scale_dev));
    } else {
        SSG_DISPATCH_DTYPE(dtype, linear_dgrad_scalar_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, w, (float*)dx, m, k, nout, jpb, inv_scale_dev));
    }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_linear_wgrad(const void* x, const void* dy, float* dw, float* dbias, int dtype, int m, int k, int nout, ssg_stream_t s) {
    SSG_CHECK_ARG(m > 0 && k > 0 && nout > 0 && nout <= 65535 * WG_J, "linear_wgrad: bad shape");
    dim3 grid((unsigned)(((k + 3) / 4 + 255) / 256), (unsigned)((nout + WG_J - 1) / WG_J));
    SSG_DISPATCH_DTYPE(dtype, linear_wgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, dw, dbias, m, k, nout));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
