// Attention-U-Net pieces of the arch zoo (archs.py:115-142 Attention_block, archs.py:848-861 up_conv):
//   * nearest-neighbour x2 upsampling and its adjoint (nn.Upsample(scale_factor=2), default mode 'nearest'),
//   * the attention gate  y = x * sigmoid(z)  with ONE gate value per pixel (z = BatchNorm2d(1) output of `psi`),
//     and its backward  dx = dy * s,  dz = s (1 - s) * sum_c dy x.
// All four are HBM-bound: every live tensor is read once and written once, 16-byte vectors along the channel axis.
#include "common.cuh"

namespace ssg {

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) nearest2x_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpp = c / V;                                  // vectors per pixel
    const long long total = (long long)n * h * w * vpp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long orow = (long long)2 * w * c;            // elements per output row
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cv = (int)(i % vpp) * V;
        long long p = i / vpp;
        const int px = (int)(p % w); p /= w;
        const int py = (int)(p % h);
        const long long img = p / h;
        const T* src = x + i * V;
        T* dst = y + ((img * 2 * h + 2 * py) * 2 * w + 2 * px) * (long long)c + cv;
        if (VEC) {
            Vec<T> v; v.load(src);
            v.store(dst); v.store(dst + c); v.store(dst + orow); v.store(dst + orow + c);
        } else {
            const T v = *src;
            dst[0] = v; dst[c] = v; dst[orow] = v; dst[orow + c] = v;
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) nearest2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpp = c / V;
    const long long total = (long long)n * h * w * vpp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long orow = (long long)2 * w * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cv = (int)(i % vpp) * V;
        long long p = i / vpp;
        const int px = (int)(p % w); p /= w;
        const int py = (int)(p % h);
        const long long img = p / h;
        const T* src = dy + ((img * 2 * h + 2 * py) * 2 * w + 2 * px) * (long long)c + cv;
        if (VEC) {
            Vec<T> a, b, d, e;
            a.load(src); b.load(src + c); d.load(src + orow); e.load(src + orow + c);
            float fa[V], fb[V], fd[V], fe[V];
            a.get(fa); b.get(fb); d.get(fd); e.get(fe);
#pragma unroll
            for (int k = 0; k < V; ++k) fa[k] = (fa[k] + fb[k]) + (fd[k] + fe[k]);
            Vec<T> o; o.set(fa); o.store(dx + i * V);
        } else {
            dx[i] = from_f<T>((to_f(src[0]) + to_f(src[c])) + (to_f(src[orow]) + to_f(src[orow + c])));
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) pixel_gate_fwd_kernel(const T* __restrict__ x, const T* __restrict__ z, T* __restrict__ y,
                                                             long long rows, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpp = c / V;
    const long long total = rows * vpp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / vpp;
        const float s = sigmoidf_(to_f(z[r]));
        if (VEC) {
            Vec<T> v; v.load(x + i * V);
            float f[V]; v.get(f);
#pragma unroll
            for (int k = 0; k < V; ++k) f[k] *= s;
            Vec<T> o; o.set(f); o.store(y + i * V);
        } else {
            y[i] = from_f<T>(to_f(x[i]) * s);
        }
    }
}

// G lanes (a power of two <= 32) own one pixel row: they stride over its channel vectors, then reduce sum_c dy*x with
// xor shuffles that stay inside the aligned G-lane group.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) pixel_gate_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ z,
                                                             T* __restrict__ dx, T* __restrict__ dz, long long rows, int c, int G) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpp = c / V;
    const int lane = threadIdx.x & (G - 1);
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = (long long)gridDim.x * blockDim.x / G;
    // every lane of a warp runs the same number of iterations (rows are rounded up per warp), so the shuffles are convergent
    const long long iters = (rows + ngroups - 1) / ngroups;
    for (long long it = 0; it < iters; ++it) {
        const long long r = group + it * ngroups;
        float acc = 0.f, s = 0.f;
        if (r < rows) {
            s = sigmoidf_(to_f(z[r]));
            for (int v = lane; v < vpp; v += G) {
                const long long off = r * c + (long long)v * V;
                if (VEC) {
                    Vec<T> a, b; a.load(dy + off); b.load(x + off);
                    float fa[V], fb[V]; a.get(fa); b.get(fb);
#pragma unroll
                    for (int k = 0; k < V; ++k) { acc = fmaf(fa[k], fb[k], acc); fa[k] *= s; }
                    Vec<T> o; o.set(fa); o.store(dx + off);
                } else {
                    const float d = to_f(dy[off]);
                    acc = fmaf(d, to_f(x[off]), acc);
                    dx[off] = from_f<T>(d * s);
                }
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (r < rows && lane == 0) dz[r] = from_f<T>(acc * s * (1.f - s));
    }
}

static inline int gate_group(int vpp) {
    int g = 1;
    while (g < vpp && g < 32) g <<= 1;
    return g;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_upsample_nearest2x_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample_nearest2x_fwd: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0) {
            unsigned g = grid_for((long long)n * h * w * (c / V), 256);
            nearest2x_fwd_kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, h, w, c);
        } else {
            unsigned g = grid_for((long long)n * h * w * c, 256);
            nearest2x_fwd_kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, h, w, c);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_upsample_nearest2x_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample_nearest2x_bwd: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0) {
            unsigned g = grid_for((long long)n * h * w * (c / V), 256);
            nearest2x_bwd_kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (T*)dx, n, h, w, c);
        } else {
            unsigned g = grid_for((long long)n * h * w * c, 256);
            nearest2x_bwd_kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (T*)dx, n, h, w, c);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_pixel_gate_fwd(const void* x, const void* z, void* y, int dtype, long long rows, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0, "pixel_gate_fwd: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0) {
            unsigned g = grid_for(rows * (c / V), 256);
            pixel_gate_fwd_kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)z, (T*)y, rows, c);
        } else {
            unsigned g = grid_for(rows * c, 256);
            pixel_gate_fwd_kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)z, (T*)y, rows, c);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_pixel_gate_bwd(const void* dy, const void* x, const void* z, void* dx, void* dz, int dtype, long long rows, int c,
                       ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && c > 0, "pixel_gate_bwd: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (c % V == 0) {
            const int G = gate_group(c / V);
            unsigned g = grid_for(rows * G, 256);
            pixel_gate_bwd_kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)z, (T*)dx, (T*)dz, rows, c, G);
        } else {
            const int G = gate_group(c);
            unsigned g = grid_for(rows * G, 256);
            pixel_gate_bwd_kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)z, (T*)dx, (T*)dz, rows, c, G);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
