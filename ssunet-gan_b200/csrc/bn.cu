// Per-channel statistics and (Sync)BatchNorm forward/backward on NHWC activations.
// HBM-bound: every kernel streams its operands once with 16-byte vector accesses; per-thread
// fp32 partials -> shared-memory fp32 -> one fp64 atomic per channel per block.
#include "common.cuh"
#include <stdlib.h>

namespace ssg {

constexpr int BN_THREADS = 256;

// The forward's per-channel affine y = fmaf(x, sc, sh): sc = inv_std * gamma, sh = beta - mean * sc.  ONE definition, explicit
// rounding, because the backward re-derives the activation mask from x with it and must reproduce the forward bit for bit.
__device__ __forceinline__ float bn_scale(float inv_std, float gamma) { return __fmul_rn(inv_std, gamma); }
__device__ __forceinline__ float bn_shift(float beta, float mean, float sc) { return __fmaf_rn(-mean, sc, beta); }

// Generic "reduce over rows, per channel" skeleton.  F(c, vals...) is applied by the callers.
// REDUCE_MODE 0: stats (sum x, sum x^2); 1: BN backward (sum dz, sum dz*xhat)
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS) channel_reduce_vec_kernel(
    const T* __restrict__ a,      // MODE0: x        MODE1: dy
    const T* __restrict__ yout,   // MODE1: post-activation output (may be null when act == none)
    const T* __restrict__ xin,    // MODE1: BN input x
    long long rows, int C, long long rows_per_block, const float* __restrict__ mean, const float* __restrict__ inv_std,
    int act, float slope, int with_sq, double* __restrict__ sums,
    const float* __restrict__ fsc = nullptr, const float* __restrict__ fsh = nullptr, int recompute = 0) {
    // recompute (MODE 1, no residual): the activation mask is re-derived from x with the forward's own expression
    // fmaf(x, sc, sh) (bn_apply_rows_kernel; sc / sh as bn_finalize_kernel stored them at forward time) instead of being read
    // back from y -- one operand less to stream.
    constexpr int V = Vec<T>::N;
    // Block-level combination without atomics: every thread parks its fp32 partial sums in shared memory ([row group][C]), then
    // thread c adds the row groups of channel c in a fixed order in fp64 and issues ONE fp64 atomic into the result.  The
    // statistics are therefore reproducible run to run (only the cross-block fp64 atomics are unordered: 1e-16).  The first
    // version used fp32 shared-memory atomics: the order in which warps hit them moved mean / inv_std of a few channels by one
    // fp32 ulp between identical launches, which the 8-sample BatchNorm at a U-Net bottleneck occasionally amplified to 1e-3
    // on the logits; it also serialised badly for thin tensors (C = 8: 256 threads on 16 addresses).
    extern __shared__ float part[];  // [2][rpb][C]
    const int vpr = C / V;
    const int lanes = vpr < BN_THREADS ? vpr : BN_THREADS;   // threads cooperating on one row
    const int rpb = BN_THREADS / lanes;                       // rows processed per iteration
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    float* p0 = part;
    float* p1 = part + (size_t)rpb * C;
    const long long r_begin = (long long)blockIdx.x * rows_per_block;
    long long r_end = r_begin + rows_per_block;
    if (r_end > rows) r_end = rows;
    if (rsub < rpb) {
        for (int v = lane; v < vpr; v += lanes) {
            float acc0[V], acc1[V], mu[V], is[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; mu[i] = 0.f; is[i] = 1.f; }
            float rsc[V], rsh[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { rsc[i] = 0.f; rsh[i] = 0.f; }
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < V; ++i) { mu[i] = mean[v * V + i]; is[i] = inv_std[v * V + i]; }
                if (recompute) {
#pragma unroll
                    for (int i = 0; i < V; ++i) { rsc[i] = fsc[v * V + i]; rsh[i] = fsh[v * V + i]; }
                }
            }
            const bool read_y = (MODE == 1) && act != SSG_ACT_NONE && !recompute;
            // U rows per trip with all loads issued before the first use: U independent 16-byte requests per operand
            // in flight per thread (a single dependent load per trip left this kernel latency-bound at < 20 % of HBM)
            constexpr int U = 4;                        // (8 rows in flight measured slower for the single-operand mode)
            long long r = r_begin + rsub;
            for (; r + (long long)(U - 1) * rpb < r_end; r += (long long)U * rpb) {
                Vec<T> va[U], vx[U], vy[U];
#pragma unroll
                for (int u = 0; u < U; ++u) va[u].load(a + (r + (long long)u * rpb) * C + (long long)v * V);
                if (MODE == 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) vx[u].load(xin + (r + (long long)u * rpb) * C + (long long)v * V);
                    if (read_y) {
#pragma unroll
                        for (int u = 0; u < U; ++u) vy[u].load(yout + (r + (long long)u * rpb) * C + (long long)v * V);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float fa[V]; va[u].get(fa);
                    if (MODE == 0) {
#pragma unroll
                        for (int i = 0; i < V; ++i) { acc0[i] += fa[i]; acc1[i] = fmaf(fa[i], fa[i], acc1[i]); }
                    } else {
                        float fx[V]; vx[u].get(fx);
                        if (read_y) {
                            float fy[V]; vy[u].get(fy);
#pragma unroll
                            for (int i = 0; i < V; ++i) fa[i] *= act_grad_from_out(fy[i], act, slope);
                        } else if (recompute) {
#pragma unroll
                            for (int i = 0; i < V; ++i) fa[i] *= act_grad_from_out(fmaf(fx[i], rsc[i], rsh[i]), act, slope);
                        }
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            acc0[i] += fa[i];
                            acc1[i] = fmaf(fa[i], (fx[i] - mu[i]) * is[i], acc1[i]);
                        }
                    }
                }
            }
            for (; r < r_end; r += rpb) {
                const long long off = r * C + (long long)v * V;
                Vec<T> va; va.load(a + off);
                float fa[V]; va.get(fa);
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < V; ++i) { acc0[i] += fa[i]; acc1[i] = fmaf(fa[i], fa[i], acc1[i]); }
                } else {
                    Vec<T> vx; vx.load(xin + off);
                    float fx[V]; vx.get(fx);
                    if (read_y) {
                        Vec<T> vy; vy.load(yout + off);
                        float fy[V]; vy.get(fy);
#pragma unroll
                        for (int i = 0; i < V; ++i) fa[i] *= act_grad_from_out(fy[i], act, slope);
                    } else if (recompute) {
#pragma unroll
                        for (int i = 0; i < V; ++i) fa[i] *= act_grad_from_out(fmaf(fx[i], rsc[i], rsh[i]), act, slope);
                    }
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        acc0[i] += fa[i];
                        acc1[i] = fmaf(fa[i], (fx[i] - mu[i]) * is[i], acc1[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {            // slot (rsub, channel) has exactly one owner in the block
                p0[(size_t)rsub * C + v * V + i] = acc0[i];
                p1[(size_t)rsub * C + v * V + i] = acc1[i];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        double a0 = 0.0, a1 = 0.0;
        for (int r = 0; r < rpb; ++r) { a0 += (double)p0[(size_t)r * C + c]; a1 += (double)p1[(size_t)r * C + c]; }
        atomicAdd(&sums[c], a0);
        if (with_sq) atomicAdd(&sums[C + c], a1);
    }
}

// Any C (tiny or not a multiple of the vector width): element-strided, shared-memory atomics.
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS) channel_reduce_scalar_kernel(
    const T* __restrict__ a, const T* __restrict__ yout, const T* __restrict__ xin, long long rows, int C,
    long long rows_per_block, const float* __restrict__ mean, const float* __restrict__ inv_std, int act, float slope,
    int with_sq, double* __restrict__ sums) {
    extern __shared__ double smd[];      // fp64 block accumulators: see channel_reduce_vec_kernel
    double* s0 = smd;
    double* s1 = smd + C;
    for (int i = threadIdx.x; i < 2 * C; i += BN_THREADS) smd[i] = 0.0;
    __syncthreads();
    const long long e_begin = (long long)blockIdx.x * rows_per_block * C;
    long long e_end = e_begin + rows_per_block * C;
    if (e_end > rows * C) e_end = rows * C;
    for (long long e = e_begin + threadIdx.x; e < e_end; e += BN_THREADS) {
        const int c = (int)(e % C);
        float v = to_f(a[e]);
        if (MODE == 0) {
            atomicAdd(&s0[c], (double)v);
            if (with_sq) atomicAdd(&s1[c], (double)(v * v));
        } else {
            if (act != SSG_ACT_NONE) v *= act_grad_from_out(to_f(yout[e]), act, slope);
            atomicAdd(&s0[c], (double)v);
            atomicAdd(&s1[c], (double)(v * (to_f(xin[e]) - mean[c]) * inv_std[c]));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        atomicAdd(&sums[c], s0[c]);
        if (with_sq) atomicAdd(&sums[C + c], s1[c]);
    }
}

template <typename T, int MODE>
static int launch_channel_reduce(const T* a, const T* y, const T* x, long long rows, int C, const float* mean,
                                 const float* inv_std, int act, float slope, int with_sq, double* sums, cudaStream_t st,
                                 const float* fsc = nullptr, const float* fsh = nullptr, int recompute = 0) {
    constexpr int V = Vec<T>::N;
    SSG_CHECK_ARG(rows > 0 && C > 0 && C <= 8192, "channel reduce: rows=%lld C=%d unsupported", rows, C);
    SSG_CHECK_ARG(!recompute || C % V == 0, "channel reduce: mask recomputation needs C %% %d == 0", V);
    SSG_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    // aim for ~8 waves of blocks, but keep >= 64 rows per block so the fp64 atomics stay negligible
    long long blocks = (long long)sm_count_cached() * 8;
    long long rpb = (rows + blocks - 1) / blocks;
    if (rpb < 64) rpb = 64;
    blocks = (rows + rpb - 1) / rpb;
    const int vpr_h = C / V, lanes_h = vpr_h < BN_THREADS ? vpr_h : BN_THREADS;
    // vec kernel: fp32 partials [2][row groups][C]; scalar kernel (C not a multiple of the vector width): fp64 [2][C]
    size_t smem = (C % V == 0) ? sizeof(float) * 2 * (size_t)(BN_THREADS / (lanes_h > 0 ? lanes_h : 1)) * C : sizeof(double) * 2 * C;
    if (smem > 48 * 1024) {     // C > 3072: opt in to the large dynamic shared-memory window once per instantiation
        static bool big_v = false, big_s = false;
        if (C % V == 0 && !big_v) {
            SSG_CHECK_CUDA(cudaFuncSetAttribute(channel_reduce_vec_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            big_v = true;
        } else if (C % V != 0 && !big_s) {
            SSG_CHECK_CUDA(cudaFuncSetAttribute(channel_reduce_scalar_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            big_s = true;
        }
    }
    if (C % V == 0)
        channel_reduce_vec_kernel<T, MODE><<<(unsigned)blocks, BN_THREADS, smem, st>>>(a, y, x, rows, C, rpb, mean, inv_std, act, slope, with_sq, sums,
                                                                                       fsc, fsh, recompute);
    else
        channel_reduce_scalar_kernel<T, MODE><<<(unsigned)blocks, BN_THREADS, smem, st>>>(a, y, x, rows, C, rpb, mean, inv_std, act, slope, with_sq, sums);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, float eps, float momentum,
                                   int sync_quirk, float* running_mean, float* running_var, float* mean, float* inv_std,
                                   long long* num_batches_tracked = nullptr, const float* gamma = nullptr,
                                   const float* beta = nullptr, float* sc_out = nullptr, float* sh_out = nullptr) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) num_batches_tracked[0] += 1;      // nn.BatchNorm2d's counter (F.batch_norm path)
    if (c >= C) return;
    const double s = sums[c], ss = sums[C + c];
    const double mu = s / count;
    double sumvar = ss - s * mu;              // batchnorm.py:119
    if (sumvar < 0) sumvar = 0;
    const double bias_var = sumvar / count;
    const double unbias_var = count > 1 ? sumvar / (count - 1) : bias_var;
    mean[c] = (float)mu;
    if (sync_quirk) {                         // batchnorm.py:127: bias_var.clamp(eps) ** -0.5
        float bv = (float)bias_var;
        inv_std[c] = 1.0f / sqrtf(bv < eps ? eps : bv);
    } else {
        inv_std[c] = 1.0f / sqrtf((float)bias_var + eps);
    }
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbias_var;
    }
    if (sc_out) {          // the forward's affine y = fma(x, sc, sh), kept for the backward's mask recomputation
        const float sc = bn_scale(inv_std[c], gamma ? gamma[c] : 1.f);
        sc_out[c] = sc;
        sh_out[c] = bn_shift(beta ? beta[c] : 0.f, mean[c], sc);
    }
}

__global__ void bn_eval_prepare_kernel(const float* rm, const float* rv, int C, float eps, float* mean, float* inv_std) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rm[c];
    inv_std[c] = 1.0f / sqrtf(rv[c] + eps);
}

// y = act(x * sc[c] + sh[c] + residual)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                        long long rows, int C, const float* __restrict__ mean,
                                                        const float* __restrict__ inv_std, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int act, float slope) {
    extern __shared__ float sm[];
    float* sc = sm;
    float* sh = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = bn_scale(inv_std[c], gamma ? gamma[c] : 1.f);
        sc[c] = s;
        sh[c] = bn_shift(beta ? beta[c] : 0.f, mean[c], s);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        constexpr int V = Vec<T>::N;
        const int vpr = C / V;
        const long long total = rows * vpr;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c0 = (int)(i % vpr) * V;
            Vec<T> vx; vx.load(x + i * V);
            float f[V]; vx.get(f);
            float r[V];
            if (res) { Vec<T> vr; vr.load(res + i * V); vr.get(r); }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float v = fmaf(f[k], sc[c0 + k], sh[c0 + k]);
                if (res) v += r[k];
                f[k] = apply_act(v, act, slope);
            }
            Vec<T> vo; vo.set(f); vo.store(y + i * V);
        }
    } else {
        const long long total = rows * C;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = (int)(i % C);
            float v = fmaf(to_f(x[i]), sc[c], sh[c]);
            if (res) v += to_f(res[i]);
            y[i] = from_f<T>(apply_act(v, act, slope));
        }
    }
}

// dz = dy*act'(y); dx = g*is*(dz - m0 - xhat*m1) with m0 = sum0/count, m1 = sum1/count (training) or g*is*dz (eval)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ yout,
                                                            const T* __restrict__ x, T* __restrict__ dx, T* __restrict__ dres,
                                                            long long rows, int C, const float* __restrict__ mean,
                                                            const float* __restrict__ inv_std, const float* __restrict__ gamma,
                                                            const double* __restrict__ sums, double count, int act, float slope,
                                                            int training) {
    extern __shared__ float sm[];
    float* k_scale = sm;          // gamma*inv_std
    float* k_m0 = sm + C;         // sum dz / count
    float* k_m1 = sm + 2 * C;     // inv_std * sum(dz xhat) / count
    float* k_mu = sm + 3 * C;
    float* k_is = sm + 4 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        k_scale[c] = inv_std[c] * (gamma ? gamma[c] : 1.f);
        k_m0[c] = training ? (float)(sums[c] / count) : 0.f;
        k_m1[c] = training ? (float)(sums[C + c] / count) : 0.f;
        k_mu[c] = mean[c];
        k_is[c] = inv_std[c];
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        constexpr int V = Vec<T>::N;
        const int vpr = C / V;
        const long long total = rows * vpr;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c0 = (int)(i % vpr) * V;
            Vec<T> vd; vd.load(dy + i * V);
            float d[V]; vd.get(d);
            if (act != SSG_ACT_NONE) {
                Vec<T> vy; vy.load(yout + i * V);
                float fy[V]; vy.get(fy);
#pragma unroll
                for (int k = 0; k < V; ++k) d[k] *= act_grad_from_out(fy[k], act, slope);
            }
            if (dres) { Vec<T> vr; vr.set(d); vr.store(dres + i * V); }
            Vec<T> vx; vx.load(x + i * V);
            float fx[V]; vx.get(fx);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int c = c0 + k;
                const float xhat = (fx[k] - k_mu[c]) * k_is[c];
                d[k] = k_scale[c] * (d[k] - k_m0[c] - xhat * k_m1[c]);
            }
            Vec<T> vo; vo.set(d); vo.store(dx + i * V);
        }
    } else {
        const long long total = rows * C;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = (int)(i % C);
            float d = to_f(dy[i]);
            if (act != SSG_ACT_NONE) d *= act_grad_from_out(to_f(yout[i]), act, slope);
            if (dres) dres[i] = from_f<T>(d);
            const float xhat = (to_f(x[i]) - k_mu[c]) * k_is[c];
            dx[i] = from_f<T>(k_scale[c] * (d - k_m0[c] - xhat * k_m1[c]));
        }
    }
}

// Row-strided variants for C % V == 0 and C / V <= 256: a thread keeps ONE channel vector for its whole life, so the
// per-channel coefficients live in registers (no shared-memory table, no per-element div/mod), and U rows are loaded
// before the first use (U independent 16-byte requests per operand in flight per thread).
//   forward : y  = act(sc * x + sh + residual)
//   backward: dz = dy * act'(y);  dx = A * dz + B * x + K  with A = gamma*inv_std, B = -A*inv_std*m1, K = A*(mean*inv_std*m1 - m0)
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_rows_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                             long long rows, int C, const float* __restrict__ mean,
                                                             const float* __restrict__ inv_std, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int act, float slope) {
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const int lanes = C / V, rpb = 256 / lanes;
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    if (rsub >= rpb) return;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = lane * V + k;
        const float s = bn_scale(inv_std[c], gamma ? gamma[c] : 1.f);
        sc[k] = s;
        sh[k] = bn_shift(beta ? beta[c] : 0.f, mean[c], s);
    }
    const long long rstride = (long long)gridDim.x * rpb;
    const long long col = (long long)lane * V;
    long long r = (long long)blockIdx.x * rpb + rsub;
    for (; r + (U - 1) * rstride < rows; r += U * rstride) {
        Vec<T> vx[U], vr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vx[u].load(x + (r + u * rstride) * C + col);
        if (res) {
#pragma unroll
            for (int u = 0; u < U; ++u) vr[u].load(res + (r + u * rstride) * C + col);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[V], fr[V];
            vx[u].get(f);
            if (res) vr[u].get(fr);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float v = fmaf(f[k], sc[k], sh[k]);
                if (res) v += fr[k];
                f[k] = apply_act(v, act, slope);
            }
            Vec<T> vo; vo.set(f); vo.store(y + (r + u * rstride) * C + col);
        }
    }
    for (; r < rows; r += rstride) {
        Vec<T> vx; vx.load(x + r * C + col);
        float f[V], fr[V];
        vx.get(f);
        if (res) { Vec<T> vr; vr.load(res + r * C + col); vr.get(fr); }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            float v = fmaf(f[k], sc[k], sh[k]);
            if (res) v += fr[k];
            f[k] = apply_act(v, act, slope);
        }
        Vec<T> vo; vo.set(f); vo.store(y + r * C + col);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_rows_kernel(const T* __restrict__ dy, const T* __restrict__ yout,
                                                                 const T* __restrict__ x, T* __restrict__ dx, T* __restrict__ dres,
                                                                 long long rows, int C, const float* __restrict__ mean,
                                                                 const float* __restrict__ inv_std, const float* __restrict__ gamma,
                                                                 const double* __restrict__ sums, double count, int act, float slope,
                                                                 int training, const float* __restrict__ fsc = nullptr,
                                                                 const float* __restrict__ fsh = nullptr, int recompute = 0) {
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    const int lanes = C / V, rpb = 256 / lanes;
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    if (rsub >= rpb) return;
    float ka[V], kb[V], kc[V], rsc[V], rsh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = lane * V + k;
        const float is = inv_std[c], a = bn_scale(is, gamma ? gamma[c] : 1.f);
        const float m0 = training ? (float)(sums[c] / count) : 0.f;
        const float m1 = training ? (float)(sums[C + c] / count) : 0.f;
        ka[k] = a;                 // gamma as it is NOW (the reference's BN backward reads the live parameter, train.py:111-115)
        kb[k] = -a * is * m1;
        kc[k] = a * (mean[c] * is * m1 - m0);
        rsc[k] = recompute ? fsc[c] : 0.f;      // the forward's affine, stored at forward time
        rsh[k] = recompute ? fsh[c] : 0.f;
    }
    const bool has_act = act != SSG_ACT_NONE && !recompute;
    const bool rc_act = act != SSG_ACT_NONE && recompute;
    const long long rstride = (long long)gridDim.x * rpb;
    const long long col = (long long)lane * V;
    long long r = (long long)blockIdx.x * rpb + rsub;
    for (; r + (U - 1) * rstride < rows; r += U * rstride) {
        Vec<T> vd[U], vx[U], vy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) vd[u].load(dy + (r + u * rstride) * C + col);
#pragma unroll
        for (int u = 0; u < U; ++u) vx[u].load(x + (r + u * rstride) * C + col);
        if (has_act) {
#pragma unroll
            for (int u = 0; u < U; ++u) vy[u].load(yout + (r + u * rstride) * C + col);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float d[V], fx[V];
            vd[u].get(d);
            vx[u].get(fx);
            if (has_act) {
                float fy[V]; vy[u].get(fy);
#pragma unroll
                for (int k = 0; k < V; ++k) d[k] *= act_grad_from_out(fy[k], act, slope);
            } else if (rc_act) {
#pragma unroll
                for (int k = 0; k < V; ++k) d[k] *= act_grad_from_out(fmaf(fx[k], rsc[k], rsh[k]), act, slope);
            }
            const long long off = (r + u * rstride) * C + col;
            if (dres) { Vec<T> vr; vr.set(d); vr.store(dres + off); }
#pragma unroll
            for (int k = 0; k < V; ++k) d[k] = fmaf(ka[k], d[k], fmaf(kb[k], fx[k], kc[k]));
            Vec<T> vo; vo.set(d); vo.store(dx + off);
        }
    }
    for (; r < rows; r += rstride) {
        const long long off = r * C + col;
        Vec<T> vd; vd.load(dy + off);
        Vec<T> vx; vx.load(x + off);
        float d[V], fx[V];
        vd.get(d);
        vx.get(fx);
        if (has_act) {
            Vec<T> vy; vy.load(yout + off);
            float fy[V]; vy.get(fy);
#pragma unroll
            for (int k = 0; k < V; ++k) d[k] *= act_grad_from_out(fy[k], act, slope);
        } else if (rc_act) {
#pragma unroll
            for (int k = 0; k < V; ++k) d[k] *= act_grad_from_out(fmaf(fx[k], rsc[k], rsh[k]), act, slope);
        }
        if (dres) { Vec<T> vr; vr.set(d); vr.store(dres + off); }
#pragma unroll
        for (int k = 0; k < V; ++k) d[k] = fmaf(ka[k], d[k], fmaf(kb[k], fx[k], kc[k]));
        Vec<T> vo; vo.set(d); vo.store(dx + off);
    }
}

// grid for the row-strided kernels: enough blocks for ~8 resident per SM, never more blocks than row groups
static inline bool bn_rows_enabled() {
    static const bool off = getenv("SSG_BN_GENERIC") != nullptr;       // debugging aid: force the table-driven kernels
    return !off;
}
static inline unsigned rows_grid(long long rows, int rpb) {
    long long b = (rows + rpb - 1) / rpb;
    long long cap = (long long)sm_count_cached() * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

__global__ void bn_param_grads_kernel(const double* sums, int C, float* dgamma, float* dbeta, int accumulate = 0) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (accumulate) {          // straight into the optimiser's gradient arena (zeroed by zero_grad; D runs two passes per step)
        if (dbeta) dbeta[c] += (float)sums[c];
        if (dgamma) dgamma[c] += (float)sums[C + c];
    } else {
        if (dbeta) dbeta[c] = (float)sums[c];
        if (dgamma) dgamma[c] = (float)sums[C + c];
    }
}
__global__ void accum_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += (float)src[i];
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_channel_stats(const void* x, int dtype, long long rows, int c, double* sums, int with_sq, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, return (launch_channel_reduce<T, 0>((const T*)x, nullptr, nullptr, rows, c, nullptr, nullptr,
                                                                  SSG_ACT_NONE, 0.f, with_sq, sums, (cudaStream_t)s)));
    return SSG_OK;
}

int ssg_bn_finalize(const double* sums, double count, int c, float eps, float momentum, int sync_quirk, float* running_mean,
                    float* running_var, float* mean, float* inv_std, ssg_stream_t s) {
    SSG_CHECK_ARG(c > 0 && count > 0, "bn_finalize: bad args");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)s>>>(sums, count, c, eps, momentum, sync_quirk, running_mean,
                                                                     running_var, mean, inv_std);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_finalize_count(const double* sums, double count, int c, float eps, float momentum, int sync_quirk, float* running_mean,
                          float* running_var, float* mean, float* inv_std, long long* num_batches_tracked, const float* gamma,
                          const float* beta, float* sc_out, float* sh_out, ssg_stream_t s) {
    SSG_CHECK_ARG(c > 0 && count > 0, "bn_finalize: bad args");
    SSG_CHECK_ARG((sc_out == nullptr) == (sh_out == nullptr), "bn_finalize: sc_out and sh_out go together");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)s>>>(sums, count, c, eps, momentum, sync_quirk, running_mean,
                                                                     running_var, mean, inv_std, num_batches_tracked, gamma, beta,
                                                                     sc_out, sh_out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_eval_prepare(const float* rm, const float* rv, int c, float eps, float* mean, float* inv_std, ssg_stream_t s) {
    bn_eval_prepare_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)s>>>(rm, rv, c, eps, mean, inv_std);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_apply(const void* x, const void* residual, void* y, int dtype, long long rows, int c, const float* mean,
                 const float* inv_std, const float* gamma, const float* beta, int act, float slope, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0 && c <= 8192, "bn_apply: bad shape");
    size_t smem = sizeof(float) * 2 * c;
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0 && c / V <= 256 && bn_rows_enabled()) {
            const int rpb = 256 / (c / V);
            bn_apply_rows_kernel<T><<<rows_grid(rows, rpb), 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)residual, (T*)y, rows, c, mean, inv_std, gamma, beta, act, slope);
        } else if (c % V == 0) {
            unsigned g = grid_for(rows * (c / V), 256 * 2);
            bn_apply_kernel<T, true><<<g, 256, smem, (cudaStream_t)s>>>((const T*)x, (const T*)residual, (T*)y, rows, c, mean, inv_std, gamma, beta, act, slope);
        } else {
            unsigned g = grid_for(rows * c, 256 * 4);
            bn_apply_kernel<T, false><<<g, 256, smem, (cudaStream_t)s>>>((const T*)x, (const T*)residual, (T*)y, rows, c, mean, inv_std, gamma, beta, act, slope);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_bwd_reduce(const void* dy, const void* y, const void* x, int dtype, long long rows, int c, const float* mean,
                      const float* inv_std, int act, float slope, double* sums, ssg_stream_t s) {
    SSG_CHECK_ARG(act == SSG_ACT_NONE || y != nullptr, "bn_bwd_reduce: activation needs the forward output");
    SSG_DISPATCH_DTYPE(dtype, return (launch_channel_reduce<T, 1>((const T*)dy, (const T*)y, (const T*)x, rows, c, mean, inv_std,
                                                                  act, slope, 1, sums, (cudaStream_t)s)));
    return SSG_OK;
}

int ssg_bn_bwd_apply(const void* dy, const void* y, const void* x, void* dx, void* dres, int dtype, long long rows, int c,
                     const float* mean, const float* inv_std, const float* gamma, const double* sums, double count, int act,
                     float slope, int training, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0 && c <= 2448, "bn_bwd_apply: bad shape (C <= 2448: the table-driven kernel keeps 5 C floats in 48 KB)");
    SSG_CHECK_ARG(act == SSG_ACT_NONE || y != nullptr, "bn_bwd_apply: activation needs the forward output");
    size_t smem = sizeof(float) * 5 * c;
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0 && c / V <= 256 && bn_rows_enabled()) {
            const int rpb = 256 / (c / V);
            bn_bwd_apply_rows_kernel<T><<<rows_grid(rows, rpb), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)y, (const T*)x, (T*)dx, (T*)dres, rows, c, mean, inv_std, gamma, sums, count, act, slope, training);
        } else if (c % V == 0) {
            unsigned g = grid_for(rows * (c / V), 256 * 2);
            bn_bwd_apply_kernel<T, true><<<g, 256, smem, (cudaStream_t)s>>>((const T*)dy, (const T*)y, (const T*)x, (T*)dx, (T*)dres, rows, c, mean, inv_std, gamma, sums, count, act, slope, training);
        } else {
            unsigned g = grid_for(rows * c, 256 * 4);
            bn_bwd_apply_kernel<T, false><<<g, 256, smem, (cudaStream_t)s>>>((const T*)dy, (const T*)y, (const T*)x, (T*)dx, (T*)dres, rows, c, mean, inv_std, gamma, sums, count, act, slope, training);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_param_grads(const double* sums, int c, float* dgamma, float* dbeta, ssg_stream_t s) {
    bn_param_grads_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)s>>>(sums, c, dgamma, dbeta);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_bn_param_grads_acc(const double* sums, int c, float* dgamma, float* dbeta, ssg_stream_t s) {
    bn_param_grads_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)s>>>(sums, c, dgamma, dbeta, 1);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_accum_f64_f32(const double* src, float* dst, int n, ssg_stream_t s) {
    SSG_CHECK_ARG(src && dst && n > 0, "accum_f64_f32: bad args");
    accum_f64_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)s>>>(src, dst, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

/* BN backward without re-reading the forward output: the activation mask is recomputed from x (no residual). */
int ssg_bn_bwd_reduce_rc(const void* dy, const void* x, int dtype, long long rows, int c, const float* mean, const float* inv_std,
                         const float* fwd_sc, const float* fwd_sh, int act, float slope, double* sums, ssg_stream_t s) {
    SSG_CHECK_ARG(fwd_sc && fwd_sh, "bn_bwd_reduce_rc: the forward's sc / sh are required");
    SSG_DISPATCH_DTYPE(dtype, return (launch_channel_reduce<T, 1>((const T*)dy, nullptr, (const T*)x, rows, c, mean, inv_std, act, slope,
                                                                  1, sums, (cudaStream_t)s, fwd_sc, fwd_sh, 1)));
    return SSG_OK;
}

int ssg_bn_bwd_apply_rc(const void* dy, const void* x, void* dx, int dtype, long long rows, int c, const float* mean,
                        const float* inv_std, const float* gamma, const float* fwd_sc, const float* fwd_sh, const double* sums,
                        double count, int act, float slope, int training, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0 && fwd_sc && fwd_sh, "bn_bwd_apply_rc: bad arguments");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        SSG_CHECK_ARG(c % V == 0 && c / V <= 256, "bn_bwd_apply_rc: C=%d needs the row-strided kernel (C %% %d == 0, C / %d <= 256)", c, V, V);
        const int rpb = 256 / (c / V);
        bn_bwd_apply_rows_kernel<T><<<rows_grid(rows, rpb), 256, 0, (cudaStream_t)s>>>((const T*)dy, nullptr, (const T*)x, (T*)dx, nullptr, rows, c, mean,
                                                                                       inv_std, gamma, sums, count, act, slope, training, fwd_sc, fwd_sh, 1);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
