// Layout / dtype plumbing kernels: NCHW<->NHWC, casts, concat/split, weight packing.
#include "common.cuh"

namespace ssg {

// [n][C][HW] (float) -> [n][HW][C] (T): 32x32 smem tile transpose, coalesced on both sides.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int CD, long long HW) {
    // CD >= C: destination channel count (storage padding); channels [C, CD) are written as zeros
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* s = src + (long long)n * C * HW;
    T* d = dst + (long long)n * CD * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i;
        long long p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? s[(long long)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        long long p = p0 + i;
        int c = c0 + threadIdx.x;
        if (c < CD && p < HW) d[p * CD + c] = from_f<T>(tile[threadIdx.x][i]);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int CS, long long HW) {
    // CS >= C: source channel count (storage padding); only the first C channels are copied
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const T* s = src + (long long)n * CS * HW;
    float* d = dst + (long long)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        long long p = p0 + i;
        int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? to_f(s[p * CS + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i;
        long long p = p0 + threadIdx.x;
        if (c < C && p < HW) d[(long long)c * HW + p] = tile[threadIdx.x][i];
    }
}

// Thin tensors (C <= 8 source channels: images, masks, logits): one thread per pixel, C coalesced plane reads and one
// 16-byte (CD == 8, bf16) or element-wise store.  The 32 x 32 tile transpose above wastes 29/32 of its lanes here.
template <typename T, int CD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_thin_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, long long HW) {
    const int n = blockIdx.y;
    const float* s = src + (long long)n * C * HW;
    T* d = dst + (long long)n * CD * HW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        float f[CD];
#pragma unroll
        for (int c = 0; c < CD; ++c) f[c] = c < C ? s[(long long)c * HW + p] : 0.f;
        if (CD == Vec<T>::N) {
            Vec<T> v; v.set(f); v.store(d + p * CD);
        } else {
#pragma unroll
            for (int c = 0; c < CD; ++c) d[p * CD + c] = from_f<T>(f[c]);
        }
    }
}
template <typename T, int CS>
__global__ void __launch_bounds__(256) nhwc_to_nchw_thin_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, long long HW) {
    const int n = blockIdx.y;
    const T* s = src + (long long)n * CS * HW;
    float* d = dst + (long long)n * C * HW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
        float f[CS];
        if (CS == Vec<T>::N) {
            Vec<T> v; v.load(s + p * CS); v.get(f);
        } else {
#pragma unroll
            for (int c = 0; c < CS; ++c) f[c] = to_f(s[p * CS + c]);
        }
#pragma unroll
        for (int c = 0; c < CS; ++c)
            if (c < C) d[(long long)c * HW + p] = f[c];
    }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = from_f<D>(to_f(src[i]));
}

// out[row][0:ca] = a[row], out[row][ca:ca+cb] = b[row]; 16-byte vectors when channels allow.
template <typename T, bool SPLIT>
__global__ void concat2_kernel(T* __restrict__ a, int ca, T* __restrict__ b, int cb, T* __restrict__ cat, long long rows) {
    constexpr int V = Vec<T>::N;
    const int ct = ca + cb;
    const int vpr = ct / V;  // vectors per row
    const long long total = rows * vpr;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        long long r = i / vpr;
        int cv = (int)(i - r * vpr) * V;
        T* part = (cv < ca) ? (a + r * ca + cv) : (b + r * cb + (cv - ca));
        T* whole = cat + r * ct + cv;
        Vec<T> v;
        if (SPLIT) { v.load(whole); v.store(part); }
        else { v.load(part); v.store(whole); }
    }
}
template <typename T, bool SPLIT>
__global__ void concat2_scalar_kernel(T* __restrict__ a, int ca, T* __restrict__ b, int cb, T* __restrict__ cat, long long rows) {
    const int ct = ca + cb;
    const long long total = rows * ct;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        long long r = i / ct;
        int c = (int)(i - r * ct);
        T* part = (c < ca) ? (a + r * ca + c) : (b + r * cb + (c - ca));
        if (SPLIT) *part = cat[i]; else cat[i] = *part;
    }
}

// OIHW fp32 -> packed layouts.  One thread per destination element.
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int layout, int cout, int cin, int kh,
                                   int kw, const float* __restrict__ inv_scale, int cout_p, int cin_p) {
    // destination dims are the padded channel counts (cout_p >= cout, cin_p >= cin); padding is zero
    const long long total = (long long)cout_p * cin_p * kh * kw;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float sc = inv_scale ? inv_scale[0] : 1.f;
    int r, s, c, k;
    long long t = i;
    if (layout == SSG_W_RSKC) {  // [r][s][k][c]
        c = (int)(t % cin_p); t /= cin_p; k = (int)(t % cout_p); t /= cout_p; s = (int)(t % kw); r = (int)(t / kw);
    } else {                     // [r][s][c][k] (optionally flipped taps)
        k = (int)(t % cout_p); t /= cout_p; c = (int)(t % cin_p); t /= cin_p; s = (int)(t % kw); r = (int)(t / kw);
        if (layout == SSG_W_RSCK_FLIP) { r = kh - 1 - r; s = kw - 1 - s; }
    }
    dst[i] = (k < cout && c < cin) ? from_f<T>(w[(((long long)k * cin + c) * kh + r) * kw + s] * sc) : from_f<T>(0.f);
}

// Every packed operand of a network in ONE launch (the per-tensor kernel above costs one launch per weight and layout: ~190 per
// step).  A block finds its record by binary search over first_block and converts ONE tile of SSG_PACK_TILE x SSG_PACK_TILE
// (output x input) channels with all ks^2 taps through shared memory: the OIHW source is read in contiguous runs of
// 32 * ks^2 floats per output channel and both destination layouts are written in 32-element runs.  (The first version computed
// one destination element per thread straight from the source: a 36-byte read stride, i.e. nine L2 sectors fetched per 128
// useful bytes -- 0.40 ms per step for the generator's 2 x 34 M packed elements; the tile version moves the same bytes once.)
struct __align__(8) PackDesc {
    const float* src;
    void* dst;
    int layout, cout, cin, ks, cout_p, cin_p;
    long long first_block;
};
static_assert(sizeof(PackDesc) == SSG_PACK_DESC_BYTES, "PackDesc layout is part of the C ABI");

template <typename T, int TAPS>
__device__ __forceinline__ void pack_tile(const PackDesc& d, int tb, float (*tile)[SSG_PACK_TILE * 9 + 1]) {
    constexpr int TL = SSG_PACK_TILE, ROW = TL * TAPS;
    const int tiles_c = (d.cin_p + TL - 1) / TL;
    const int k0 = (tb / tiles_c) * TL, c0 = (tb % tiles_c) * TL;
    const int run = (d.cin - c0 < TL ? (d.cin - c0 > 0 ? d.cin - c0 : 0) : TL) * TAPS;     // live floats per source row of the tile
    const float* src = d.src + ((long long)k0 * d.cin + c0) * TAPS;
    for (int e = threadIdx.x; e < TL * ROW; e += 256) {
        const int k = e / ROW, off = e - k * ROW;
        tile[k][off] = (k0 + k < d.cout && off < run) ? src[(long long)k * d.cin * TAPS + off] : 0.f;
    }
    __syncthreads();
    T* dst = reinterpret_cast<T*>(d.dst);
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;        // a warp writes one 32-element run; 8 runs per pass
    if (d.layout == SSG_W_RSKC) {           // [tap][k][c], c fastest
        if (c0 + lane < d.cin_p) {
            for (int r = grp; r < TL * TAPS; r += 8) {
                const int k = r % TL, tap = r / TL;
                if (k0 + k < d.cout_p) dst[((long long)tap * d.cout_p + k0 + k) * d.cin_p + c0 + lane] = from_f<T>(tile[k][lane * TAPS + tap]);
            }
        }
    } else {                                // [tap][c][k], k fastest (optionally with the taps mirrored)
        const bool flip = d.layout == SSG_W_RSCK_FLIP;
        if (k0 + lane < d.cout_p) {
            for (int r = grp; r < TL * TAPS; r += 8) {
                const int c = r % TL, tap = r / TL;
                if (c0 + c < d.cin_p)
                    dst[((long long)tap * d.cin_p + c0 + c) * d.cout_p + k0 + lane] = from_f<T>(tile[lane][c * TAPS + (flip ? TAPS - 1 - tap : tap)]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackDesc* __restrict__ descs, int n_descs) {
    __shared__ float tile[SSG_PACK_TILE][SSG_PACK_TILE * 9 + 1];
    const long long b = blockIdx.x;
    int lo = 0, hi = n_descs - 1;
    while (lo < hi) {                       // last record with first_block <= b
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    if (d.ks == 3) pack_tile<T, 9>(d, (int)(b - d.first_block), tile);      // ks is 1 or 3 (checked by the caller)
    else pack_tile<T, 1>(d, (int)(b - d.first_block), tile);
}

}  // namespace ssg
using namespace ssg;

template <bool SPLIT>
static int concat_impl(void* a, int ca, void* b, int cb, void* cat, int dtype, long long rows, ssg_stream_t s) {
    SSG_CHECK_ARG(ca > 0 && cb > 0 && rows > 0, "concat2: bad shape");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec<T>::N;
        if (ca % V == 0 && cb % V == 0) {
            unsigned g = grid_for(rows * ((ca + cb) / V), 256);
            concat2_kernel<T, SPLIT><<<g, 256, 0, (cudaStream_t)s>>>((T*)a, ca, (T*)b, cb, (T*)cat, rows);
        } else {
            unsigned g = grid_for(rows * (ca + cb), 256);
            concat2_scalar_kernel<T, SPLIT><<<g, 256, 0, (cudaStream_t)s>>>((T*)a, ca, (T*)b, cb, (T*)cat, rows);
        }
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" {

int ssg_nchw_to_nhwc_pad(const float* src, void* dst, int dtype, int n, int c, int c_dst, int h, int w, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && c > 0 && c_dst >= c && h > 0 && w > 0 && n <= 65535, "nchw_to_nhwc: bad shape");
    long long hw = (long long)h * w;
    if (c_dst == 8 && dtype == SSG_BF16) {
        dim3 tg((unsigned)((hw + 1023) / 1024), (unsigned)n);
        nchw_to_nhwc_thin_kernel<bf16, 8><<<tg, 256, 0, (cudaStream_t)s>>>(src, (bf16*)dst, c, hw);
        SSG_CHECK_LAUNCH();
        return SSG_OK;
    }
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c_dst + 31) / 32), (unsigned)n), block(32, 8);
    SSG_DISPATCH_DTYPE(dtype, nchw_to_nhwc_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>(src, (T*)dst, c, c_dst, hw));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_nchw_to_nhwc(const float* src, void* dst, int dtype, int n, int c, int h, int w, ssg_stream_t s) {
    return ssg_nchw_to_nhwc_pad(src, dst, dtype, n, c, c, h, w, s);
}

int ssg_nhwc_to_nchw_pad(const void* src, int dtype, float* dst, int n, int c, int c_src, int h, int w, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && c > 0 && c_src >= c && h > 0 && w > 0 && n <= 65535, "nhwc_to_nchw: bad shape");
    long long hw = (long long)h * w;
    if (c_src == 8 && dtype == SSG_BF16) {
        dim3 tg((unsigned)((hw + 1023) / 1024), (unsigned)n);
        nhwc_to_nchw_thin_kernel<bf16, 8><<<tg, 256, 0, (cudaStream_t)s>>>((const bf16*)src, dst, c, hw);
        SSG_CHECK_LAUNCH();
        return SSG_OK;
    }
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n), block(32, 8);
    SSG_DISPATCH_DTYPE(dtype, nhwc_to_nchw_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>((const T*)src, dst, c, c_src, hw));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_nhwc_to_nchw(const void* src, int dtype, float* dst, int n, int c, int h, int w, ssg_stream_t s) {
    return ssg_nhwc_to_nchw_pad(src, dtype, dst, n, c, c, h, w, s);
}

int ssg_cast(const void* src, int sd, void* dst, int dd, long long n, ssg_stream_t s) {
    if (n <= 0) return SSG_OK;
    unsigned g = grid_for(n, 256 * 4);
    cudaStream_t st = (cudaStream_t)s;
    if (sd == SSG_F32 && dd == SSG_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)src, (bf16*)dst, n);
    else if (sd == SSG_BF16 && dd == SSG_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)src, (float*)dst, n);
    else if (sd == SSG_F32 && dd == SSG_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
    else if (sd == SSG_BF16 && dd == SSG_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)src, (bf16*)dst, n);
    else { set_error("cast: bad dtypes"); return SSG_ERR_INVALID_ARG; }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_concat2(const void* a, int ca, const void* b, int cb, void* out, int dtype, long long rows, ssg_stream_t s) {
    return concat_impl<false>((void*)a, ca, (void*)b, cb, out, dtype, rows, s);
}
int ssg_split2(const void* in, void* a, int ca, void* b, int cb, int dtype, long long rows, ssg_stream_t s) {
    return concat_impl<true>(a, ca, b, cb, (void*)in, dtype, rows, s);
}

int ssg_pack_conv_weight_pad(const float* w, void* dst, int dtype, int layout, int cout, int cin, int kh, int kw, int cout_p,
                             int cin_p, const float* inv_scale_dev, ssg_stream_t s) {
    SSG_CHECK_ARG(cout > 0 && cin > 0 && kh > 0 && kw > 0 && layout >= 0 && layout <= 2 && cout_p >= cout && cin_p >= cin,
                  "pack_conv_weight: bad args");
    long long total = (long long)cout_p * cin_p * kh * kw;
    unsigned g = (unsigned)((total + 255) / 256);
    SSG_DISPATCH_DTYPE(dtype, pack_weight_kernel<T><<<g, 256, 0, (cudaStream_t)s>>>(w, (T*)dst, layout, cout, cin, kh, kw, inv_scale_dev,
                                                                                    cout_p, cin_p));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_pack_conv_weights_multi(const void* descs_dev, int n_descs, long long total_blocks, int dtype, ssg_stream_t s) {
    SSG_CHECK_ARG(descs_dev && n_descs > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "pack_conv_weights_multi: bad args");
    // (records with ksize other than 1 or 3 are not accepted: the caller packs those with ssg_pack_conv_weight_pad)
    SSG_DISPATCH_DTYPE(dtype, pack_weights_multi_kernel<T><<<(unsigned)total_blocks, 256, 0, (cudaStream_t)s>>>((const PackDesc*)descs_dev, n_descs));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int ssg_pack_conv_weight(const float* w, void* dst, int dtype, int layout, int cout, int cin, int kh, int kw,
                         const float* inv_scale_dev, ssg_stream_t s) {
    return ssg_pack_conv_weight_pad(w, dst, dtype, layout, cout, cin, kh, kw, cout, cin, inv_scale_dev, s);
}

}  // extern "C"
