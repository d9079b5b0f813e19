// Elementwise kernels: SPADE modulation, activations, add, NaN scrub, gradient clamp + Adam, scaling.
#include "common.cuh"

namespace ssg {

// y = x * (1 + gamma) + beta, gb = [rows][2C]
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) spade_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gb, T* __restrict__ y,
                                                         long long rows, int C) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpr = C / V;
    const long long total = rows * vpr, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / vpr;
        const int c0 = (int)(i - r * vpr) * V;
        float fx[V], fg[V], fb[V];
        if (VEC) {
            Vec<T> v;
            v.load(x + r * C + c0); v.get(fx);
            v.load(gb + r * 2 * C + c0); v.get(fg);
            v.load(gb + r * 2 * C + C + c0); v.get(fb);
        } else {
            fx[0] = to_f(x[r * C + c0]); fg[0] = to_f(gb[r * 2 * C + c0]); fb[0] = to_f(gb[r * 2 * C + C + c0]);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) fx[k] = fmaf(fx[k], 1.f + fg[k], fb[k]);
        if (VEC) { Vec<T> v; v.set(fx); v.store(y + r * C + c0); } else { y[r * C + c0] = from_f<T>(fx[0]); }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) spade_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gb,
                                                         T* __restrict__ dx, T* __restrict__ dgb, long long rows, int C) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpr = C / V;
    const long long total = rows * vpr, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / vpr;
        const int c0 = (int)(i - r * vpr) * V;
        float fd[V], fx[V], fg[V], o1[V], o2[V];
        if (VEC) {
            Vec<T> v;
            v.load(dy + r * C + c0); v.get(fd);
            v.load(x + r * C + c0); v.get(fx);
            v.load(gb + r * 2 * C + c0); v.get(fg);
        } else {
            fd[0] = to_f(dy[r * C + c0]); fx[0] = to_f(x[r * C + c0]); fg[0] = to_f(gb[r * 2 * C + c0]);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) { o1[k] = fd[k] * (1.f + fg[k]); o2[k] = fd[k] * fx[k]; }
        if (VEC) {
            Vec<T> v;
            v.set(o1); v.store(dx + r * C + c0);
            v.set(o2); v.store(dgb + r * 2 * C + c0);
            v.set(fd); v.store(dgb + r * 2 * C + C + c0);
        } else {
            dx[r * C + c0] = from_f<T>(o1[0]);
            dgb[r * 2 * C + c0] = from_f<T>(o2[0]);
            dgb[r * 2 * C + C + c0] = from_f<T>(fd[0]);
        }
    }
}

// 16-byte vector body + scalar tail (all buffers come from the torch allocator: 16-byte aligned bases)
// Row-strided SPADE backward that also reduces the bias gradients of the gamma|beta convolution: a thread keeps ONE channel
// vector for its whole life (like bn_*_rows), accumulates sum(dy*x) and sum(dy) of the bf16-ROUNDED values it stores, and the
// block folds them through shared memory into fp64 colsum[2C] (pre-zeroed) -- the separate per-channel reduction pass over
// dgb (2C channels, the largest tensor of the block) disappears.
template <typename T>
__global__ void __launch_bounds__(256) spade_bwd_rows_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gb,
                                                              T* __restrict__ dx, T* __restrict__ dgb, long long rows, int C,
                                                              double* __restrict__ colsum) {
    constexpr int V = Vec<T>::N;
    extern __shared__ float sred[];            // [2C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
    const int lanes = C / V, rpb = 256 / lanes;
    const int lane = threadIdx.x % lanes, rsub = threadIdx.x / lanes;
    if (rsub < rpb) {
        const long long rstride = (long long)gridDim.x * rpb;
        const int c0 = lane * V;
        float s_g[V], s_b[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s_g[k] = s_b[k] = 0.f;
        for (long long r = (long long)blockIdx.x * rpb + rsub; r < rows; r += rstride) {
            Vec<T> vd, vx, vg;
            vd.load(dy + r * C + c0); vx.load(x + r * C + c0); vg.load(gb + r * 2 * C + c0);
            float fd[V], fx[V], fg[V], o1[V], o2[V];
            vd.get(fd); vx.get(fx); vg.get(fg);
#pragma unroll
            for (int k = 0; k < V; ++k) { o1[k] = fd[k] * (1.f + fg[k]); o2[k] = fd[k] * fx[k]; }
            Vec<T> v;
            v.set(o1); v.store(dx + r * C + c0);
            v.set(o2); v.store(dgb + r * 2 * C + c0);
            v.get(o2);                                  // the rounded values, as the reduction pass would have read them
            vd.store(dgb + r * 2 * C + C + c0);
#pragma unroll
            for (int k = 0; k < V; ++k) { s_g[k] += o2[k]; s_b[k] += fd[k]; }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) { atomicAdd(&sred[c0 + k], s_g[k]); atomicAdd(&sred[C + c0 + k], s_b[k]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&colsum[i], (double)sred[i]);
}

template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int act, float slope, int vec_ok) {
    constexpr int V = Vec<T>::N;
    const long long nv = vec_ok ? n / V : 0, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec<T> v; v.load(x + i * V);
        float f[V]; v.get(f);
#pragma unroll
        for (int k = 0; k < V; ++k) f[k] = apply_act(f[k], act, slope);
        v.set(f); v.store(y + i * V);
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = from_f<T>(apply_act(to_f(x[i]), act, slope));
}
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n, int act, float slope, int vec_ok) {
    constexpr int V = Vec<T>::N;
    const long long nv = vec_ok ? n / V : 0, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec<T> vd, vy; vd.load(dy + i * V); vy.load(y + i * V);
        float fd[V], fy[V]; vd.get(fd); vy.get(fy);
#pragma unroll
        for (int k = 0; k < V; ++k) fd[k] *= act_grad_from_out(fy[k], act, slope);
        vd.set(fd); vd.store(dx + i * V);
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dx[i] = from_f<T>(to_f(dy[i]) * act_grad_from_out(to_f(y[i]), act, slope));
}
template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long long n, int vec_ok) {
    constexpr int V = Vec<T>::N;
    const long long nv = vec_ok ? n / V : 0, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec<T> va, vb; va.load(a + i * V); vb.load(b + i * V);
        float fa[V], fb[V]; va.get(fa); vb.get(fb);
#pragma unroll
        for (int k = 0; k < V; ++k) fa[k] += fb[k];
        va.set(fa); va.store(o + i * V);
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        o[i] = from_f<T>(to_f(a[i]) + to_f(b[i]));
}
__global__ void nan_scrub_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = x[i];
        y[i] = (v != v) ? 0.f : v;
    }
}
__global__ void nan_scrub_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = x[i];
        dx[i] = (v != v) ? 0.f : dy[i];
    }
}
__global__ void clamp_kernel(float* __restrict__ g, long long n, float clip) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        g[i] = fminf(fmaxf(g[i], -clip), clip);   // NaN propagates like torch.clamp
}
// one Adam element: g <- clamp(g * gscale) [+ wd * p: torch.optim.Adam's L2 weight decay, train.py:290]; m, v, p updated exactly
// in torch.optim.Adam's operation order
__device__ __forceinline__ void adam_elem(float& pi, float& gi, float& mi, float& vi, float b1, float b2, float eps, float bc2_sqrt,
                                          float step, float clip, float gscale, float wd = 0.f) {
    gi *= gscale;
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
    float ge = gi;                           // .grad keeps the clamped gradient; the decayed one is a temporary in torch too
    if (wd != 0.f) ge = gi + wd * pi;
    mi = b1 * mi + (1.f - b1) * ge;
    vi = b2 * vi + (1.f - b2) * ge * ge;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step * (mi / denom);
}
// float4 body (the four arenas are torch allocations: 16-byte aligned) + scalar tail
__device__ __forceinline__ void adam_span(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                          long long n, long long stride, float b1, float b2, float eps, float bc2_sqrt, float step,
                                          float clip, float gscale, float wd = 0.f) {
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    const long long nv = vec ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 p4 = reinterpret_cast<float4*>(p)[i], g4 = reinterpret_cast<float4*>(g)[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        adam_elem(p4.x, g4.x, m4.x, v4.x, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
        adam_elem(p4.y, g4.y, m4.y, v4.y, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
        adam_elem(p4.z, g4.z, m4.z, v4.z, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
        adam_elem(p4.w, g4.w, m4.w, v4.w, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
        reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(g)[i] = g4;
        reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam_elem(p[i], g[i], m[i], v[i], b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
}
// torch.optim.Adam (no amsgrad, no weight decay) after clip_gradient's element clamp.
__global__ void __launch_bounds__(256) clamp_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, long long n, float lr, float b1, float b2,
                                                          float eps, float bc1, float bc2_sqrt, float clip, float gscale, float wd) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float step = lr / bc1;
    adam_span(p, g, m, v, n, stride, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
}
// Graph-capturable variant: the step count lives on the device (a captured launch cannot carry per-step host scalars).
__global__ void step_inc_kernel(float* step) { step[0] += 1.f; }
__global__ void __launch_bounds__(256) clamp_adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                              float* __restrict__ v, long long n, float lr, float b1, float b2,
                                                              float eps, const float* __restrict__ step_dev, float clip, float gscale,
                                                              float wd) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float t = step_dev[0];
    const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
    const float step = lr / bc1;
    adam_span(p, g, m, v, n, stride, b1, b2, eps, bc2_sqrt, step, clip, gscale, wd);
}
// Hyper-parameters on the device as well: hp = {lr, beta1, beta2, eps, clip, grad_scale, weight_decay, -}.  A captured launch
// then follows an LR schedule / a changed clip: the host rewrites the eight floats between replays.
__global__ void __launch_bounds__(256) clamp_adam_hp_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                             float* __restrict__ v, long long n, const float* __restrict__ hp,
                                                             const float* __restrict__ step_dev) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], clip = hp[4], gscale = hp[5], wd = hp[6];
    const float t = step_dev[0];
    const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
    adam_span(p, g, m, v, n, stride, b1, b2, eps, bc2_sqrt, lr / bc1, clip, gscale, wd);
}
__global__ void scale_by_dev_kernel(const float* __restrict__ w, const float* __restrict__ sc, float* __restrict__ o, long long n) {
    const float s = sc[0];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = w[i] * s;
}

static inline int aligned16(const void* a, const void* b = nullptr, const void* c = nullptr) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0 ? 1 : 0;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_spade_modulate_fwd(const void* x, const void* gb, void* y, int dtype, long long rows, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0, "spade_modulate: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        if (c % Vec<T>::N == 0) spade_fwd_kernel<T, true><<<grid_for(rows * c / Vec<T>::N, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)gb, (T*)y, rows, c);
        else spade_fwd_kernel<T, false><<<grid_for(rows * c, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)gb, (T*)y, rows, c);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_spade_modulate_bwd(const void* dy, const void* x, const void* gb, void* dx, void* dgb, int dtype, long long rows, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0, "spade_modulate: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        if (c % Vec<T>::N == 0) spade_bwd_kernel<T, true><<<grid_for(rows * c / Vec<T>::N, 256), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)gb, (T*)dx, (T*)dgb, rows, c);
        else spade_bwd_kernel<T, false><<<grid_for(rows * c, 256), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)gb, (T*)dx, (T*)dgb, rows, c);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_spade_modulate_bwd_sums(const void* dy, const void* x, const void* gb, void* dx, void* dgb, int dtype, long long rows, int c,
                                double* colsum, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0 && colsum, "spade_modulate_bwd_sums: bad arguments");
    SSG_CHECK_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * 2 * (size_t)c, (cudaStream_t)s));
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        SSG_CHECK_ARG(c % V == 0 && c / V <= 256, "spade_modulate_bwd_sums: C=%d unsupported", c);
        const int rpb = 256 / (c / V);
        long long b = (rows + rpb - 1) / rpb, cap = (long long)sm_count_cached() * 8;
        if (b > cap) b = cap;
        spade_bwd_rows_kernel<T><<<(unsigned)b, 256, sizeof(float) * 2 * c, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)gb, (T*)dx,
                                                                                               (T*)dgb, rows, c, colsum);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_act_fwd(const void* x, void* y, int dtype, long long n, int act, float slope, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, act_fwd_kernel<T><<<grid_for(n, 2048), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, act, slope, aligned16(x, y)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_act_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, int act, float slope, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, act_bwd_kernel<T><<<grid_for(n, 2048), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)y, (T*)dx, n, act, slope, aligned16(dy, y, dx)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_add(const void* a, const void* b, void* out, int dtype, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, add_kernel<T><<<grid_for(n, 2048), 256, 0, (cudaStream_t)s>>>((const T*)a, (const T*)b, (T*)out, n, aligned16(a, b, out)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_nan_scrub_fwd(const float* x, float* y, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    nan_scrub_fwd_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(x, y, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_nan_scrub_bwd(const float* dy, const float* x, float* dx, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    nan_scrub_bwd_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(dy, x, dx, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_(float* g, long long n, float clip, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    clamp_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(g, n, clip);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_adam_wd(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                      float bias_corr1, float bias_corr2, float clip, float grad_scale, float weight_decay, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    clamp_adam_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1,
                                                                      sqrtf(bias_corr2), clip, grad_scale, weight_decay);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_adam(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float bias_corr1, float bias_corr2, float clip, float grad_scale, ssg_stream_t s) {
    return ssg_clamp_adam_wd(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1, bias_corr2, clip, grad_scale, 0.f, s);
}
int ssg_clamp_adam_wd_dev(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                          float* step_dev, float clip, float grad_scale, float weight_decay, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_CHECK_ARG(step_dev != nullptr, "clamp_adam_dev: step counter missing");
    step_inc_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step_dev);
    clamp_adam_dev_kernel<<<grid_for(n, 2048), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, clip, grad_scale,
                                                                          weight_decay);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_adam_hp_dev(float* p, float* g, float* m, float* v, long long n, const float* hp_dev, float* step_dev, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_CHECK_ARG(step_dev != nullptr && hp_dev != nullptr, "clamp_adam_hp_dev: step counter / hyper-parameters missing");
    step_inc_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step_dev);
    clamp_adam_hp_kernel<<<grid_for(n, 2048), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, hp_dev, step_dev);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_adam_dev(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                       float* step_dev, float clip, float grad_scale, ssg_stream_t s) {
    return ssg_clamp_adam_wd_dev(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, clip, grad_scale, 0.f, s);
}
int ssg_scale_by_dev(const float* w, const float* inv_sigma, float* out, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    scale_by_dev_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(w, inv_sigma, out, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
