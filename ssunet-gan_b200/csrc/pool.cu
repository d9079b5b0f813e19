// Max-pool 2x2 with a 2-bit argmax code (instead of ATen's int64 flat indices), max-unpool,
// bilinear x2 (align_corners=True) and adaptive average pooling.  All HBM-bound, NHWC.
#include "common.cuh"
#include <stdlib.h>

namespace ssg {

// ---- max pool / unpool -----------------------------------------------------------------------
// One thread per (output pixel, channel vector).  Tie rule = ATen max_pool2d: scan the window in
// row-major order, replace when (val > max) || isnan(val).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ code,
                                                          int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int oh = h / 2, ow = w / 2, vpr = c / V;
    const long long total = (long long)n * oh * ow * vpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cv = (int)(i % vpr);
        long long p = i / vpr;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const int nn = (int)(p / oh);
        const long long base = (((long long)nn * h + 2 * oy) * w + 2 * ox) * c + (long long)cv * V;
        float best[V]; int arg[V];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long off = base + ((long long)(q >> 1) * w + (q & 1)) * c;
            float f[V];
            if (VEC) { Vec<T> v; v.load(x + off); v.get(f); } else { f[0] = to_f(x[off]); }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                if (q == 0 || f[k] > best[k] || f[k] != f[k]) { best[k] = f[k]; arg[k] = q; }
            }
        }
        const long long o = i * V;
        if (VEC) {
            Vec<T> v; v.set(best); v.store(y + o);
            if (V == 8) {
                uint2 cc;
                cc.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
                cc.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
                *reinterpret_cast<uint2*>(code + o) = cc;
            } else {
                *reinterpret_cast<uint32_t*>(code + o) = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
            }
        } else {
            y[o] = from_f<T>(best[0]);
            code[o] = (uint8_t)arg[0];
        }
    }
}

// dst[n, 2y+dy, 2x+dx, c] = src[n,y,x,c] if code == 2*dy+dx else 0.  One thread writes all four.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) scatter2x2_kernel(const T* __restrict__ src, const uint8_t* __restrict__ code,
                                                          T* __restrict__ dst, int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int vpr = c / V;
    const long long total = (long long)n * h * w * vpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cv = (int)(i % vpr);
        long long p = i / vpr;
        const int x_ = (int)(p % w); p /= w;
        const int y_ = (int)(p % h);
        const int nn = (int)(p / h);
        float f[V]; uint8_t cd[V];
        if (VEC) { Vec<T> v; v.load(src + i * V); v.get(f); } else { f[0] = to_f(src[i]); }
#pragma unroll
        for (int k = 0; k < V; ++k) cd[k] = code[i * V + k];
        const long long base = (((long long)nn * 2 * h + 2 * y_) * 2 * w + 2 * x_) * c + (long long)cv * V;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float o[V];
#pragma unroll
            for (int k = 0; k < V; ++k) o[k] = (cd[k] == q) ? f[k] : 0.f;
            const long long off = base + ((long long)(q >> 1) * 2 * w + (q & 1)) * c;
            if (VEC) { Vec<T> v; v.set(o); v.store(dst + off); } else { dst[off] = from_f<T>(o[0]); }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gather2x2_kernel(const T* __restrict__ src, const uint8_t* __restrict__ code,
                                                         T* __restrict__ dst, int n, int h, int w, int c) {
    const long long total = (long long)n * h * w * c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cc = (int)(i % c);
        long long p = i / c;
        const int x_ = (int)(p % w); p /= w;
        const int y_ = (int)(p % h);
        const int nn = (int)(p / h);
        const int q = code[i];
        dst[i] = src[(((long long)nn * 2 * h + 2 * y_ + (q >> 1)) * 2 * w + 2 * x_ + (q & 1)) * c + cc];
    }
}

// ---- bilinear x2, align_corners=True (ATen upsample_bilinear2d: scale = (in-1)/(out-1)) ----------
__device__ __forceinline__ void src_index(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
    const float s = scale * (float)o;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = s - (float)i0;
}

// One block row per output row (blockIdx.y = n*oh + oy): the vertical stencil is computed once per thread and all
// index arithmetic is 32-bit (the earlier flat-index version spent its time in 64-bit div/mod).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int oh = 2 * h, ow = 2 * w, vpr = c / V;
    const float sy = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f;
    const float sx = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
    const int row_vecs = ow * vpr;
    for (int row = blockIdx.y; row < n * oh; row += gridDim.y) {
        const int nn = row / oh, oy = row - nn * oh;
        int y0, y1; float ly;
        src_index(oy, sy, h, y0, y1, ly);
        const float hy = 1.f - ly;
        const T* r0 = x + ((long long)nn * h + y0) * w * c;
        const T* r1 = x + ((long long)nn * h + y1) * w * c;
        T* yo = y + (long long)row * ow * c;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_vecs; i += gridDim.x * blockDim.x) {
            const int ox = i / vpr, cv = i - ox * vpr;
            int x0, x1; float lx;
            src_index(ox, sx, w, x0, x1, lx);
            const float hx = 1.f - lx;
            const int o0 = x0 * c + cv * V, o1 = x1 * c + cv * V;
            float f00[V], f01[V], f10[V], f11[V], o[V];
            if (VEC) {
                Vec<T> v00, v01, v10, v11;
                v00.load(r0 + o0); v01.load(r0 + o1); v10.load(r1 + o0); v11.load(r1 + o1);
                v00.get(f00); v01.get(f01); v10.get(f10); v11.get(f11);
            } else {
                f00[0] = to_f(r0[o0]); f01[0] = to_f(r0[o1]); f10[0] = to_f(r1[o0]); f11[0] = to_f(r1[o1]);
            }
#pragma unroll
            for (int k = 0; k < V; ++k) o[k] = hy * (hx * f00[k] + lx * f01[k]) + ly * (hx * f10[k] + lx * f11[k]);
            if (VEC) { Vec<T> v; v.set(o); v.store(yo + (long long)i * V); } else { yo[i] = from_f<T>(o[0]); }
        }
    }
}

// weights with which input index `i` receives from the candidate outputs lo .. lo+7 (scale < 0.5: every output whose
// stencil touches i lies in [2i-3, 2i+4]); recomputes the forward's index arithmetic exactly.
__device__ __forceinline__ void adjoint_weights(int i, int in_size, int out_size, float scale, int& lo, float (&wt)[8]) {
    lo = 2 * i - 3;
    if (in_size == 1) lo = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = lo + j;
        float wv = 0.f;
        if (o >= 0 && o < out_size) {
            int i0, i1; float l;
            src_index(o, scale, in_size, i0, i1, l);
            if (i0 == i) wv += 1.f - l;
            if (i1 == i) wv += l;
        }
        wt[j] = wv;
    }
}

// Adjoint in gather form, one block row per input row: the stencil is separable, so a thread evaluates 8 + 8 candidate
// weights (not 64) and loads only the outputs whose weight product is non-zero.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c) {
    constexpr int V = VEC ? Vec<T>::N : 1;
    const int oh = 2 * h, ow = 2 * w, vpr = c / V;
    const float sy = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f;
    const float sx = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
    const int row_vecs = w * vpr;
    for (int row = blockIdx.y; row < n * h; row += gridDim.y) {
        const int nn = row / h, iy = row - nn * h;
        int oy_lo; float wy[8];
        adjoint_weights(iy, h, oh, sy, oy_lo, wy);
        const T* b = dy + (long long)nn * oh * ow * c;
        T* xo = dx + (long long)row * w * c;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_vecs; i += gridDim.x * blockDim.x) {
            const int ix = i / vpr, cv = i - ix * vpr;
            int ox_lo; float wx[8];
            adjoint_weights(ix, w, ow, sx, ox_lo, wx);
            float acc[V];
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
            for (int jy = 0; jy < 8; ++jy) {
                if (wy[jy] == 0.f) continue;
                const T* rp = b + (long long)(oy_lo + jy) * ow * c + cv * V;
#pragma unroll
                for (int jx = 0; jx < 8; ++jx) {
                    const float wgt = wy[jy] * wx[jx];
                    if (wgt == 0.f) continue;
                    float f[V];
                    if (VEC) { Vec<T> v; v.load(rp + (long long)(ox_lo + jx) * c); v.get(f); }
                    else { f[0] = to_f(rp[(long long)(ox_lo + jx) * c]); }
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[k] = fmaf(wgt, f[k], acc[k]);
                }
            }
            if (VEC) { Vec<T> v; v.set(acc); v.store(xo + (long long)i * V); } else { xo[i] = from_f<T>(acc[0]); }
        }
    }
}

// ---- structured x2 kernels ------------------------------------------------------------------------------------------
// With align_corners=True and an exact x2 factor the stencil is regular: outputs 2b-1 and 2b read inputs {b-1, b}
// (clamped), and input i is read by outputs 2i-1 .. 2i+2 only (upsample2x_structure_ok() verifies this on the host for
// the given size by replaying src_index in fp32).  The forward therefore produces a 2 x 2 output block per thread from
// four input vectors (4 loads / 4 stores instead of 16 loads / 4 stores), the adjoint gathers from a 4 x 4 block.
// The arithmetic (products and their order) is the same as in the generic kernels above.
__device__ __forceinline__ void pair_weights(int o, float scale, int in_size, int ra, int rb, float& wa, float& wb) {
    int i0, i1; float l;
    src_index(o, scale, in_size, i0, i1, l);
    wa = (i0 == ra ? 1.f - l : 0.f) + (i1 == ra ? l : 0.f);
    wb = rb != ra ? (i0 == rb ? 1.f - l : 0.f) + (i1 == rb ? l : 0.f) : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_fwd_blk_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c) {
    constexpr int V = Vec<T>::N;
    const int oh = 2 * h, ow = 2 * w, vpr = c / V;
    const float sy = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f;
    const float sx = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
    const int row_items = (w + 1) * vpr;
    for (int row = blockIdx.y; row < n * (h + 1); row += gridDim.y) {
        const int nn = row / (h + 1), bi = row - nn * (h + 1);
        const int ra = bi > 0 ? bi - 1 : 0, rb = bi < h ? bi : h - 1;
        const int oy0 = 2 * bi - 1, oy1 = 2 * bi;
        float wy[2][2];
        pair_weights(oy0 < 0 ? 0 : oy0, sy, h, ra, rb, wy[0][0], wy[0][1]);
        pair_weights(oy1 > oh - 1 ? oh - 1 : oy1, sy, h, ra, rb, wy[1][0], wy[1][1]);
        const T* pa = x + ((long long)nn * h + ra) * w * c;
        const T* pb = x + ((long long)nn * h + rb) * w * c;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_items; i += gridDim.x * blockDim.x) {
            const int bk = i / vpr, cv = i - bk * vpr;
            const int ca = bk > 0 ? bk - 1 : 0, cb = bk < w ? bk : w - 1;
            const int ox0 = 2 * bk - 1, ox1 = 2 * bk;
            float wx[2][2];
            pair_weights(ox0 < 0 ? 0 : ox0, sx, w, ca, cb, wx[0][0], wx[0][1]);
            pair_weights(ox1 > ow - 1 ? ow - 1 : ox1, sx, w, ca, cb, wx[1][0], wx[1][1]);
            Vec<T> vaa, vab, vba, vbb;
            vaa.load(pa + (long long)ca * c + cv * V); vab.load(pa + (long long)cb * c + cv * V);
            vba.load(pb + (long long)ca * c + cv * V); vbb.load(pb + (long long)cb * c + cv * V);
            float faa[V], fab[V], fba[V], fbb[V];
            vaa.get(faa); vab.get(fab); vba.get(fba); vbb.get(fbb);
#pragma unroll
            for (int jy = 0; jy < 2; ++jy) {
                const int oy = jy ? oy1 : oy0;
                if (oy < 0 || oy >= oh) continue;
#pragma unroll
                for (int jx = 0; jx < 2; ++jx) {
                    const int ox = jx ? ox1 : ox0;
                    if (ox < 0 || ox >= ow) continue;
                    float o[V];
#pragma unroll
                    for (int k = 0; k < V; ++k)
                        o[k] = wy[jy][0] * (wx[jx][0] * faa[k] + wx[jx][1] * fab[k]) + wy[jy][1] * (wx[jx][0] * fba[k] + wx[jx][1] * fbb[k]);
                    Vec<T> vo; vo.set(o);
                    vo.store(y + (((long long)nn * oh + oy) * ow + ox) * c + cv * V);
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_blk_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c) {
    constexpr int V = Vec<T>::N;
    const int oh = 2 * h, ow = 2 * w, vpr = c / V;
    const float sy = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f;
    const float sx = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
    const int row_vecs = w * vpr;
    for (int row = blockIdx.y; row < n * h; row += gridDim.y) {
        const int nn = row / h, iy = row - nn * h;
        float wy[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = 2 * iy - 1 + j;
            float wa = 0.f, wb;
            if (o >= 0 && o < oh) pair_weights(o, sy, h, iy, iy, wa, wb);
            wy[j] = wa;
        }
        const T* b = dy + (long long)nn * oh * ow * c;
        T* xo = dx + (long long)row * w * c;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_vecs; i += gridDim.x * blockDim.x) {
            const int ix = i / vpr, cv = i - ix * vpr;
            float wx[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = 2 * ix - 1 + j;
                float wa = 0.f, wb;
                if (o >= 0 && o < ow) pair_weights(o, sx, w, ix, ix, wa, wb);
                wx[j] = wa;
            }
            float acc[V];
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
            for (int jy = 0; jy < 4; ++jy) {
                if (wy[jy] == 0.f) continue;                       // also skips rows outside the image
                const T* rp = b + (long long)(2 * iy - 1 + jy) * ow * c + cv * V;
                Vec<T> v[4];
#pragma unroll
                for (int jx = 0; jx < 4; ++jx)
                    if (wx[jx] != 0.f) v[jx].load(rp + (long long)(2 * ix - 1 + jx) * c);
#pragma unroll
                for (int jx = 0; jx < 4; ++jx) {
                    if (wx[jx] == 0.f) continue;
                    float f[V]; v[jx].get(f);
                    const float wgt = wy[jy] * wx[jx];
#pragma unroll
                    for (int k = 0; k < V; ++k) acc[k] = fmaf(wgt, f[k], acc[k]);
                }
            }
            Vec<T> vo; vo.set(acc); vo.store(xo + (long long)i * V);
        }
    }
}

// host check of the stencil structure the *_blk kernels rely on (same fp32 arithmetic as src_index); cached per size
static bool upsample2x_structure_ok(int in_size) {
    static const bool generic_only = getenv("SSG_UPSAMPLE_GENERIC") != nullptr;      // debugging aid
    if (generic_only) return false;
    static int cache_size[16];
    static int cache_ok[16];
    static int cache_n = 0;
    for (int i = 0; i < cache_n; ++i)
        if (cache_size[i] == in_size) return cache_ok[i] != 0;
    const int out = 2 * in_size;
    const float scale = out > 1 ? (float)(in_size - 1) / (float)(out - 1) : 0.f;
    bool ok = true;
    for (int o = 0; o < out && ok; ++o) {
        const float sf = scale * (float)o;
        int i0 = (int)sf;
        if (i0 > in_size - 1) i0 = in_size - 1;
        const int i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
        const float l = sf - (float)i0;
        const int b = (o + 1) / 2;
        const int ra = b > 0 ? b - 1 : 0, rb = b < in_size ? b : in_size - 1;
        if (i0 != ra && i0 != rb) ok = false;
        if (i1 != ra && i1 != rb && l != 0.f) ok = false;          // a neighbour outside the pair may only carry weight 0
    }
    if (cache_n < 16) { cache_size[cache_n] = in_size; cache_ok[cache_n] = ok ? 1 : 0; ++cache_n; }
    return ok;
}

// ---- adaptive average pool to (oh, ow), flattened channel-major like NCHW .view(batch, -1) -----------
// ATen window: [floor(o*in/out), ceil((o+1)*in/out))
template <typename T>
__global__ void __launch_bounds__(256) adaptive_avgpool_flat_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h,
                                                                         int w, int c, int oh, int ow) {
    const long long total = (long long)n * oh * ow * c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cc = (int)(i % c);       // threads adjacent in c -> coalesced reads
        long long p = i / c;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const int nn = (int)(p / oh);
        const int y0 = (oy * h) / oh, y1 = ((oy + 1) * h + oh - 1) / oh;
        const int x0 = (ox * w) / ow, x1 = ((ox + 1) * w + ow - 1) / ow;
        float acc = 0.f;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) acc += to_f(x[(((long long)nn * h + yy) * w + xx) * c + cc]);
        y[(long long)nn * c * oh * ow + (long long)cc * oh * ow + oy * ow + ox] = from_f<T>(acc / (float)((y1 - y0) * (x1 - x0)));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) adaptive_avgpool_flat_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h,
                                                                         int w, int c, int oh, int ow) {
    const long long total = (long long)n * h * w * c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int cc = (int)(i % c);
        long long p = i / c;
        const int xx = (int)(p % w); p /= w;
        const int yy = (int)(p % h);
        const int nn = (int)(p / h);
        float acc = 0.f;
        // windows may overlap when in % out != 0: scan candidate outputs
        int oy_lo = (yy * oh) / h - 1, oy_hi = ((yy + 1) * oh + h - 1) / h;
        int ox_lo = (xx * ow) / w - 1, ox_hi = ((xx + 1) * ow + w - 1) / w;
        oy_lo = oy_lo < 0 ? 0 : oy_lo; ox_lo = ox_lo < 0 ? 0 : ox_lo;
        oy_hi = oy_hi > oh - 1 ? oh - 1 : oy_hi; ox_hi = ox_hi > ow - 1 ? ow - 1 : ox_hi;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            const int y0 = (oy * h) / oh, y1 = ((oy + 1) * h + oh - 1) / oh;
            if (yy < y0 || yy >= y1) continue;
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                const int x0 = (ox * w) / ow, x1 = ((ox + 1) * w + ow - 1) / ow;
                if (xx < x0 || xx >= x1) continue;
                acc += to_f(dy[(long long)nn * c * oh * ow + (long long)cc * oh * ow + oy * ow + ox]) / (float)((y1 - y0) * (x1 - x0));
            }
        }
        dx[i] = from_f<T>(acc);
    }
}

// Same result, staged: one block per (sample, 64-channel group).  dy is [n][c][oh][ow] (the flatten order fc1 expects), so the
// gather above reads it with a stride of oh*ow elements between adjacent channels (uncoalesced) and redoes the window
// arithmetic with integer divisions per element: 0.18 ms for a 17 MB gradient.  Here the block loads its 64 x (oh*ow) slice of dy
// with contiguous reads, pre-divides by the window areas, tabulates each input row's / column's window range once, and then
// streams dx out with 128-byte rows.
template <typename T>
__global__ void __launch_bounds__(256) adaptive_avgpool_flat_bwd_staged_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h,
                                                                                int w, int c, int oh, int ow) {
    extern __shared__ float sm_pool[];
    const int cells = oh * ow;
    float* g = sm_pool;                         // [64][cells], already divided by the window area
    int* oy_lo = reinterpret_cast<int*>(g + 64 * cells);      // per input row: first window, number of windows
    int* oy_n = oy_lo + h;
    int* ox_lo = oy_n + h;
    int* ox_n = ox_lo + w;
    const int groups = c / 64;
    const int nn = blockIdx.x / groups, c0 = (blockIdx.x % groups) * 64;
    for (int i = threadIdx.x; i < 64 * cells; i += blockDim.x) {
        const int cell = i % cells, oy = cell / ow, ox = cell % ow;
        const int y0 = (oy * h) / oh, y1 = ((oy + 1) * h + oh - 1) / oh;
        const int x0 = (ox * w) / ow, x1 = ((ox + 1) * w + ow - 1) / ow;
        g[i] = to_f(dy[(long long)nn * c * cells + (long long)c0 * cells + i]) / (float)((y1 - y0) * (x1 - x0));
    }
    for (int yy = threadIdx.x; yy < h; yy += blockDim.x) {
        int lo = -1, cnt = 0;
        for (int oy = 0; oy < oh; ++oy) {
            const int y0 = (oy * h) / oh, y1 = ((oy + 1) * h + oh - 1) / oh;
            if (yy >= y0 && yy < y1) { if (lo < 0) lo = oy; ++cnt; }
        }
        oy_lo[yy] = lo; oy_n[yy] = cnt;
    }
    for (int xx = threadIdx.x; xx < w; xx += blockDim.x) {
        int lo = -1, cnt = 0;
        for (int ox = 0; ox < ow; ++ox) {
            const int x0 = (ox * w) / ow, x1 = ((ox + 1) * w + ow - 1) / ow;
            if (xx >= x0 && xx < x1) { if (lo < 0) lo = ox; ++cnt; }
        }
        ox_lo[xx] = lo; ox_n[xx] = cnt;
    }
    __syncthreads();
    const int cc = threadIdx.x & 63, sub = threadIdx.x >> 6;       // 4 pixels x 64 channels per pass
    const float* gc = g + cc * cells;
    // blockIdx.y splits the pixels of the plane (each block re-stages the small dy slice: the first version ran n * c / 64 = 128
    // long-lived blocks and took 0.1 ms regardless of the batch)
    const int per_blk = (h * w + gridDim.y - 1) / gridDim.y;
    const int p_end = min(h * w, (int)(blockIdx.y + 1) * per_blk);
    for (int p = blockIdx.y * per_blk + sub; p < p_end; p += 4) {
        const int yy = p / w, xx = p - yy * w;
        float acc = 0.f;
        for (int a = 0; a < oy_n[yy]; ++a)
            for (int b = 0; b < ox_n[xx]; ++b) acc += gc[(oy_lo[yy] + a) * ow + ox_lo[xx] + b];
        dx[(((long long)nn * h + yy) * w + xx) * c + c0 + cc] = from_f<T>(acc);
    }
}

}  // namespace ssg
using namespace ssg;

#define SSG_VEC_LAUNCH(kernel, total_expr, ...)                                                         \
    SSG_DISPATCH_DTYPE(dtype, {                                                                         \
        constexpr int V = Vec<T>::N;                                                                    \
        if (c % V == 0) {                                                                               \
            unsigned g = grid_for((total_expr) / V, 256);                                               \
            kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>(__VA_ARGS__);                               \
        } else {                                                                                        \
            unsigned g = grid_for((total_expr), 256);                                                   \
            kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>(__VA_ARGS__);                              \
        }                                                                                               \
    });                                                                                                 \
    SSG_CHECK_LAUNCH();                                                                                 \
    return SSG_OK

// 2-D launch for the row-structured kernels: grid.x covers one row's vectors, grid.y the rows (capped at 65535)
#define SSG_ROW_LAUNCH(kernel, row_pixels, n_rows, ...)                                                 \
    SSG_DISPATCH_DTYPE(dtype, {                                                                         \
        constexpr int V = Vec<T>::N;                                                                    \
        const bool vec = c % V == 0;                                                                    \
        const long long per_row = (long long)(row_pixels) * (vec ? c / V : c);                          \
        dim3 g((unsigned)((per_row + 255) / 256 > 64 ? 64 : (per_row + 255) / 256), (unsigned)((n_rows) > 65535 ? 65535 : (n_rows))); \
        if (vec) kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>(__VA_ARGS__);                          \
        else kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>(__VA_ARGS__);                             \
    });                                                                                                 \
    SSG_CHECK_LAUNCH();                                                                                 \
    return SSG_OK

extern "C" {

int ssg_maxpool2x2_fwd(const void* x, void* y, uint8_t* code, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h >= 2 && w >= 2 && c > 0, "maxpool2x2: bad shape");
    SSG_VEC_LAUNCH(maxpool2x2_kernel, (long long)n * (h / 2) * (w / 2) * c, (const T*)x, (T*)y, code, n, h, w, c);
}
int ssg_scatter2x2(const void* src, const uint8_t* code, void* dst, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "scatter2x2: bad shape");
    SSG_VEC_LAUNCH(scatter2x2_kernel, (long long)n * h * w * c, (const T*)src, code, (T*)dst, n, h, w, c);
}
int ssg_gather2x2(const void* src, const uint8_t* code, void* dst, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "gather2x2: bad shape");
    unsigned g = grid_for((long long)n * h * w * c, 256);
    SSG_DISPATCH_DTYPE(dtype, gather2x2_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)src, code, (T*)dst, n, h, w, c));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_upsample2x_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample2x: bad shape");
    if (c % 8 == 0 && upsample2x_structure_ok(h) && upsample2x_structure_ok(w)) {
        SSG_DISPATCH_DTYPE(dtype, {
            if (c % Vec<T>::N == 0) {
                const long long per_row = (long long)(w + 1) * (c / Vec<T>::N);
                dim3 g((unsigned)((per_row + 255) / 256 > 64 ? 64 : (per_row + 255) / 256), (unsigned)(n * (h + 1) > 65535 ? 65535 : n * (h + 1)));
                upsample2x_fwd_blk_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, h, w, c);
                SSG_CHECK_LAUNCH();
                return SSG_OK;
            }
        });
    }
    SSG_ROW_LAUNCH(upsample2x_fwd_kernel, 2 * w, n * 2 * h, (const T*)x, (T*)y, n, h, w, c);
}
int ssg_upsample2x_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample2x: bad shape");
    if (c % 8 == 0 && upsample2x_structure_ok(h) && upsample2x_structure_ok(w)) {
        SSG_DISPATCH_DTYPE(dtype, {
            if (c % Vec<T>::N == 0) {
                const long long per_row = (long long)w * (c / Vec<T>::N);
                dim3 g((unsigned)((per_row + 255) / 256 > 64 ? 64 : (per_row + 255) / 256), (unsigned)(n * h > 65535 ? 65535 : n * h));
                upsample2x_bwd_blk_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (T*)dx, n, h, w, c);
                SSG_CHECK_LAUNCH();
                return SSG_OK;
            }
        });
    }
    SSG_ROW_LAUNCH(upsample2x_bwd_kernel, w, n * h, (const T*)dy, (T*)dx, n, h, w, c);
}
int ssg_adaptive_avgpool_flat_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "adaptive_avgpool: bad shape");
    unsigned g = grid_for((long long)n * oh * ow * c, 256);
    SSG_DISPATCH_DTYPE(dtype, adaptive_avgpool_flat_fwd_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, h, w, c, oh, ow));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_adaptive_avgpool_flat_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "adaptive_avgpool: bad shape");
    const size_t smem = sizeof(float) * 64 * oh * ow + sizeof(int) * 2 * (h + w);
    if (c % 64 == 0 && smem <= 40 * 1024) {
        const int groups = n * (c / 64);
        int ysplit = (4 * sm_count_cached() + groups - 1) / groups;          // ~4 blocks per SM
        if (ysplit > (h * w + 63) / 64) ysplit = (h * w + 63) / 64;
        if (ysplit < 1) ysplit = 1;
        SSG_DISPATCH_DTYPE(dtype, adaptive_avgpool_flat_bwd_staged_kernel<T><<<dim3((unsigned)groups, (unsigned)ysplit), 256, smem, (cudaStream_t)s>>>(
                                      (const T*)dy, (T*)dx, n, h, w, c, oh, ow));
        SSG_CHECK_LAUNCH();
        return SSG_OK;
    }
    unsigned g = grid_for((long long)n * h * w * c, 256);
    SSG_DISPATCH_DTYPE(dtype, adaptive_avgpool_flat_bwd_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>((const T*)dy, (T*)dx, n, h, w, c, oh, ow));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
