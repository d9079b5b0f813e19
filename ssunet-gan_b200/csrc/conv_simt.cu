// CUDA-core implicit-GEMM convolution (forward / dgrad / wgrad), NHWC, fp32 accumulation.
// General in every dimension; used for the tiny-channel layers (Cin = 3 stems, SPADE's 3/h-channel
// convs, the 64->num_classes head) where a tensor-core tile cannot be filled, for the fp32 parity
// mode, and as the on-device cross-check of the tcgen05 kernels.
#include "common.cuh"

namespace ssg {

struct ConvGeom {
    int n, h, w, cin, cout, kh, kw, stride, pad, oh, ow;
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float* f, int valid, bool vec_ok);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* f, int valid, bool vec_ok) {
    if (vec_ok && valid == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = (i < valid) ? p[i] : 0.f;
    }
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float* f, int valid, bool vec_ok) {
    if (vec_ok && valid == 4) {
        uint2 v = *reinterpret_cast<const uint2*>(p);
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = (i < valid) ? __bfloat162float(p[i]) : 0.f;
    }
}

constexpr int TM = 64, TN = 64, TK = 16;

// MODE 0: forward gather (rows = output pixels, Kd = cin, Nd = cout, src = x, weights [tap][cin][cout])
// MODE 1: dgrad gather   (rows = input pixels,  Kd = cout, Nd = cin, src = dy, weights [tap][cout][cin])
template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const T* __restrict__ src, const T* __restrict__ wgt,
                                                          const float* __restrict__ bias, T* __restrict__ out,
                                                          ConvGeom g, int act, float slope) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int Kd = MODE == 0 ? g.cin : g.cout;
    const int Nd = MODE == 0 ? g.cout : g.cin;
    const int rh = MODE == 0 ? g.oh : g.h, rw = MODE == 0 ? g.ow : g.w;     // row-space spatial dims
    const int sh = MODE == 0 ? g.h : g.oh, sw = MODE == 0 ? g.w : g.ow;     // source spatial dims
    const long long M = (long long)g.n * rh * rw;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // A loader role: one row, 4 consecutive k
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const long long am = m0 + a_row;
    int a_n = 0, a_y = 0, a_x = 0;
    const bool a_row_ok = am < M;
    if (a_row_ok) {
        a_x = (int)(am % rw);
        long long t = am / rw;
        a_y = (int)(t % rh);
        a_n = (int)(t / rh);
    }
    // B loader role: one k, 4 consecutive n
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    const bool a_vec = (Kd % 4 == 0), b_vec = (Nd % 4 == 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int r = 0; r < g.kh; ++r) {
        for (int s = 0; s < g.kw; ++s) {
            long long a_off = -1;
            if (a_row_ok) {
                int sy, sx;
                bool ok;
                if (MODE == 0) {
                    sy = a_y * g.stride - g.pad + r;
                    sx = a_x * g.stride - g.pad + s;
                    ok = sy >= 0 && sy < sh && sx >= 0 && sx < sw;
                } else {
                    int ty_ = a_y + g.pad - r, tx_ = a_x + g.pad - s;
                    ok = ty_ >= 0 && tx_ >= 0 && (ty_ % g.stride == 0) && (tx_ % g.stride == 0);
                    sy = ty_ / g.stride;
                    sx = tx_ / g.stride;
                    ok = ok && sy < sh && sx < sw;
                }
                if (ok) a_off = (((long long)a_n * sh + sy) * sw + sx) * Kd;
            }
            const T* wtap = wgt + (long long)(r * g.kw + s) * Kd * Nd;
            for (int k0 = 0; k0 < Kd; k0 += TK) {
                float av[4] = {0.f, 0.f, 0.f, 0.f};
                int kval = Kd - (k0 + a_k);
                kval = kval > 4 ? 4 : kval;
                if (a_off >= 0 && kval > 0) load4<T>(src + a_off + k0 + a_k, av, kval, a_vec);
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                int nval = Nd - (n0 + b_n);
                nval = nval > 4 ? 4 : nval;
                if (k0 + b_k < Kd && nval > 0) load4<T>(wtap + (long long)(k0 + b_k) * Nd + n0 + b_n, bv, nval, b_vec);
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 4; ++i) As[a_k + i][a_row] = av[i];
                *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < TK; ++kk) {
                    float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                    float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                    const float aa[4] = {a.x, a.y, a.z, a.w};
                    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int nn = n0 + tx * 4 + j;
            if (nn >= Nd) continue;
            float v = acc[i][j] + (bias ? bias[nn] : 0.f);
            out[m * Nd + nn] = from_f<T>(apply_act(v, act, slope));
        }
    }
}

// dW[k][c][r][s] += sum_m x[pix(m,r,s)][c] * dy[m][k], split over m with fp32 atomics.
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                          float* __restrict__ dw, ConvGeom g, long long m_per_split) {
    __shared__ __align__(16) float As[TK][TM + 4];   // [pixel][c]
    __shared__ __align__(16) float Bs[TK][TN + 4];   // [pixel][k]
    const int c_tiles = (g.cin + TM - 1) / TM;
    const int tap = blockIdx.x / c_tiles;
    const int c0 = (blockIdx.x % c_tiles) * TM;
    const int k0 = blockIdx.y * TN;
    const int r = tap / g.kw, s = tap % g.kw;
    const long long M = (long long)g.n * g.oh * g.ow;
    const long long m_begin = (long long)blockIdx.z * m_per_split;
    long long m_end = m_begin + m_per_split;
    if (m_end > M) m_end = M;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int l_p = tid >> 4, l_c = (tid & 15) * 4;   // loader: pixel l_p of the chunk, 4 channels
    const bool a_vec = (g.cin % 4 == 0), b_vec = (g.cout % 4 == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long mb = m_begin; mb < m_end; mb += TK) {
        const long long m = mb + l_p;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_end) {
            int ox = (int)(m % g.ow);
            long long t = m / g.ow;
            int oy = (int)(t % g.oh);
            int nn = (int)(t / g.oh);
            int iy = oy * g.stride - g.pad + r, ix = ox * g.stride - g.pad + s;
            int cval = g.cin - (c0 + l_c);
            cval = cval > 4 ? 4 : cval;
            if (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w && cval > 0)
                load4<T>(x + (((long long)nn * g.h + iy) * g.w + ix) * g.cin + c0 + l_c, av, cval, a_vec);
            int kval = g.cout - (k0 + l_c);
            kval = kval > 4 ? 4 : kval;
            if (kval > 0) load4<T>(dy + m * g.cout + k0 + l_c, bv, kval, b_vec);
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[l_p][l_c]) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4*>(&Bs[l_p][l_c]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty * 4 + i;
        if (c >= g.cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= g.cout) continue;
            atomicAdd(dw + (((long long)k * g.cin + c) * g.kh + r) * g.kw + s, acc[i][j]);
        }
    }
}

static int make_geom(ConvGeom& g, int n, int h, int w, int cin, int cout, int kh, int kw, int stride, int pad) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
                  "conv: bad geometry n=%d h=%d w=%d cin=%d cout=%d k=%dx%d s=%d p=%d", n, h, w, cin, cout, kh, kw, stride, pad);
    g = {n, h, w, cin, cout, kh, kw, stride, pad, (h + 2 * pad - kh) / stride + 1, (w + 2 * pad - kw) / stride + 1};
    SSG_CHECK_ARG(g.oh > 0 && g.ow > 0, "conv: empty output");
    return SSG_OK;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_conv2d_fwd_simt(const void* x, const void* w, const float* bias, void* y, int dtype, int n, int h, int w_, int cin,
                        int cout, int kh, int kw, int stride, int pad, int act, float slope, ssg_stream_t s) {
    ConvGeom g;
    int rc = make_geom(g, n, h, w_, cin, cout, kh, kw, stride, pad);
    if (rc) return rc;
    long long M = (long long)n * g.oh * g.ow;
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((cout + TN - 1) / TN));
    SSG_DISPATCH_DTYPE(dtype, conv_igemm_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)w, bias, (T*)y, g, act, slope));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_conv2d_dgrad_simt(const void* dy, const void* w, void* dx, int dtype, int n, int h, int w_, int cin, int cout, int kh,
                          int kw, int stride, int pad, ssg_stream_t s) {
    ConvGeom g;
    int rc = make_geom(g, n, h, w_, cin, cout, kh, kw, stride, pad);
    if (rc) return rc;
    long long M = (long long)n * h * w_;
    dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((cin + TN - 1) / TN));
    SSG_DISPATCH_DTYPE(dtype, conv_igemm_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)w, nullptr, (T*)dx, g, SSG_ACT_NONE, 0.f));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_conv2d_wgrad_simt(const void* x, const void* dy, float* dw, int dtype, int n, int h, int w_, int cin, int cout, int kh,
                          int kw, int stride, int pad, ssg_stream_t s) {
    ConvGeom g;
    int rc = make_geom(g, n, h, w_, cin, cout, kh, kw, stride, pad);
    if (rc) return rc;
    SSG_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cout * cin * kh * kw, (cudaStream_t)s));
    long long M = (long long)n * g.oh * g.ow;
    int tiles = kh * kw * ((cin + TM - 1) / TM) * ((cout + TN - 1) / TN);
    long long want = (4LL * sm_count_cached() + tiles - 1) / tiles;
    long long max_split = (M + 255) / 256;
    long long nsplit = want < 1 ? 1 : (want > max_split ? max_split : want);
    if (nsplit > 65535) nsplit = 65535;
    long long per = ((M + nsplit - 1) / nsplit + TK - 1) / TK * TK;
    nsplit = (M + per - 1) / per;
    dim3 grid((unsigned)(kh * kw * ((cin + TM - 1) / TM)), (unsigned)((cout + TN - 1) / TN), (unsigned)nsplit);
    SSG_DISPATCH_DTYPE(dtype, conv_wgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, dw, g, per));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
