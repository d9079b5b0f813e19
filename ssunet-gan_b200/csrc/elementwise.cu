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

template <typename T>
__global__ void act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int act, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = from_f<T>(apply_act(to_f(x[i]), act, slope));
}
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long long n, int act, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dx[i] = from_f<T>(to_f(dy[i]) * act_grad_from_out(to_f(y[i]), act, slope));
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
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
// torch.optim.Adam (no amsgrad, no weight decay) after clip_gradient's element clamp.
__global__ void __launch_bounds__(256) clamp_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, long long n, float lr, float b1, float b2,
                                                          float eps, float bc1, float bc2_sqrt, float clip, float gscale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float step = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * gscale;
        if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
        g[i] = gi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step * (mi / denom);
    }
}
// Graph-capturable variant: the step count lives on the device (a captured launch cannot carry per-step host scalars).
__global__ void step_inc_kernel(float* step) { step[0] += 1.f; }
__global__ void __launch_bounds__(256) clamp_adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                              float* __restrict__ v, long long n, float lr, float b1, float b2,
                                                              float eps, const float* __restrict__ step_dev, float clip, float gscale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float t = step_dev[0];
    const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
    const float step = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gi = g[i] * gscale;
        if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);
        g[i] = gi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step * (mi / denom);
    }
}
__global__ void scale_by_dev_kernel(const float* __restrict__ w, const float* __restrict__ sc, float* __restrict__ o, long long n) {
    const float s = sc[0];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = w[i] * s;
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
int ssg_act_fwd(const void* x, void* y, int dtype, long long n, int act, float slope, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, act_fwd_kernel<T><<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, act, slope));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_act_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, int act, float slope, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, act_bwd_kernel<T><<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)y, (T*)dx, n, act, slope));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_add(const void* a, const void* b, void* out, int dtype, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_DISPATCH_DTYPE(dtype, add_kernel<T><<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>((const T*)a, (const T*)b, (T*)out, n));
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
int ssg_clamp_adam(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float bias_corr1, float bias_corr2, float clip, float grad_scale, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    clamp_adam_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1,
                                                                      sqrtf(bias_corr2), clip, grad_scale);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_clamp_adam_dev(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                       float* step_dev, float clip, float grad_scale, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    SSG_CHECK_ARG(step_dev != nullptr, "clamp_adam_dev: step counter missing");
    step_inc_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step_dev);
    clamp_adam_dev_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, clip, grad_scale);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_scale_by_dev(const float* w, const float* inv_sigma, float* out, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    scale_by_dev_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)s>>>(w, inv_sigma, out, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
