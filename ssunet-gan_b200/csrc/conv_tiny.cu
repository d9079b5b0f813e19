// 3x3 / stride 1 / pad 1 convolutions between THIN activations: both tensors stored with 8 channels (16-byte NHWC pixels), at
// most 8 real channels on either side -- SPADE's mlp_shared (label_nc = 3 -> nhidden = 4 or 8, normalization.py:93-95) at the two
// finest U-Net levels.  On the tcgen05 path these ran at 4 TFLOP/s: a 128-row MMA tile for 108 multiply-adds per pixel is all issue
// and drain overhead (0.21-0.24 ms per direction for a 67 MB tensor whose read + write takes 0.02 ms).  Here one thread owns one
// pixel: nine 16-byte neighbour loads (L1-resident after the first touch), the 9 x CI x CO weights broadcast from shared memory,
// fp32 accumulation, one 16-byte store.  HBM-bound: algorithmic bytes = 16 B read + 16 B written per pixel.
//   forward : y  = act(conv(x, W) + b)
//   dgrad   : dx = conv(dy, W^T flipped)                  (the same loop nest with the roles of CI / CO swapped)
//   wgrad   : dW[co][ci][tap] += sum_px dy[px][co] x[px + tap][ci],  db[co] += sum_px dy[px][co]
#include "common.cuh"

namespace ssg {

__device__ __forceinline__ void load8(const bf16* __restrict__ p, bool ok, float* f) {
    if (ok) {
        Vec<bf16> v;
        v.load(p);
        v.get(f);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
}

// DGRAD == false: w_s[tap][ci][co] = W[co][ci][tap]              (CI = conv input channels, CO = conv output channels)
// DGRAD == true : the tensor read is dy (CI = the convolution's OUTPUT channels), the tensor written is dx (CO = its input
//                 channels): w_s[tap][ci][co] = W[ci][co][8 - tap]
template <int CI, int CO, bool DGRAD, int PX>
__global__ void __launch_bounds__(128) conv3x3_tiny_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            bf16* __restrict__ y, int n, int h, int wd, int w_cout, int w_cin, int act,
                                                            float slope) {
    // One thread computes PX (4, or 2 for the wider variants) consecutive pixels of a row: the 3 x (PX + 2) neighbourhood is
    // loaded once (18 16-byte loads instead of 36 at PX = 4) and every weight read from shared memory feeds PX multiply-adds (the
    // first version, one pixel per thread, was bound by its 9 x CI x CO shared-memory weight loads per pixel).
    __shared__ float w_s[9 * CI * CO];
    __shared__ float b_s[CO];
    for (int i = threadIdx.x; i < 9 * CI * CO; i += blockDim.x) {
        const int co = i % CO, ci = (i / CO) % CI, tap = i / (CO * CI);
        float v = 0.f;
        if (!DGRAD) { if (co < w_cout && ci < w_cin) v = w[((long long)co * w_cin + ci) * 9 + tap]; }
        else { if (ci < w_cout && co < w_cin) v = w[((long long)ci * w_cin + co) * 9 + (8 - tap)]; }
        w_s[i] = v;
    }
    if (threadIdx.x < CO) b_s[threadIdx.x] = (bias != nullptr && threadIdx.x < w_cout) ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    const int groups_per_row = (wd + PX - 1) / PX;
    const int total = n * h * groups_per_row;                    // < 2^31 (checked by the launcher): 32-bit index arithmetic
    const int stride = (int)(gridDim.x * blockDim.x);
    for (int gi = (int)(blockIdx.x * blockDim.x + threadIdx.x); gi < total; gi += stride) {
        const int rowi = gi / groups_per_row;                    // n * h + y
        const int x0 = (gi - rowi * groups_per_row) * PX;
        const int yy = rowi % h;
        float acc[PX][CO];
#pragma unroll
        for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int co = 0; co < CO; ++co) acc[q][co] = b_s[co];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int sy = yy + r - 1;
            const bool row_ok = sy >= 0 && sy < h;
            const bf16* xr = x + ((long long)(rowi + (r - 1)) * wd) * 8;
            float f[PX + 2][CI];
#pragma unroll
            for (int c = 0; c < PX + 2; ++c) {
                const int sx = x0 + c - 1;
                float t[8];
                load8(xr + (long long)sx * 8, row_ok && sx >= 0 && sx < wd, t);
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) f[c][ci] = t[ci];
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float* wt = w_s + (r * 3 + s) * CI * CO;
#pragma unroll
                for (int ci = 0; ci < CI; ++ci)
#pragma unroll
                    for (int co = 0; co < CO; ++co) {
                        const float wv = wt[ci * CO + co];
#pragma unroll
                        for (int q = 0; q < PX; ++q) acc[q][co] = fmaf(f[q + s][ci], wv, acc[q][co]);
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < PX; ++q) {
            if (x0 + q >= wd) break;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
            for (int co = 0; co < CO; ++co) o[co] = apply_act(acc[q][co], act, slope);
            Vec<bf16> vo;
            vo.set(o);
            vo.store(y + ((long long)rowi * wd + x0 + q) * 8);
        }
    }
}

// blockIdx.y = filter row r: one thread accumulates the 3 x CI x CO partial products of that row's taps (<= 96 registers) over
// runs of PXW = 8 consecutive pixels of an image row -- one index decomposition and PXW + 2 sliding x loads per run (the first
// version decomposed a 64-bit pixel index per pixel: two emulated 64-bit divisions cost more than the arithmetic) -- then the block
// combines the partials through shared memory in a fixed order and adds once per block into dW (fp32 atomics: a few hundred
// blocks).  Row 0's blocks also reduce the bias gradient.
template <int CI, int CO>
__global__ void __launch_bounds__(128) conv3x3_tiny_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                                                                  float* __restrict__ db, int n, int h, int wd, int w_cout, int w_cin) {
    constexpr int NW = 3 * CI * CO;
    constexpr int PXW = 8;
    const int r = blockIdx.y;
    float acc[NW];
    float accb[CO];
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < CO; ++i) accb[i] = 0.f;
    const int runs_per_row = (wd + PXW - 1) / PXW;
    const int rows = n * h;                                  // < 2^31 / runs_per_row (checked by the launcher)
    const int total = rows * runs_per_row;
    const int stride = (int)(gridDim.x * blockDim.x);
    for (int gi = (int)(blockIdx.x * blockDim.x + threadIdx.x); gi < total; gi += stride) {
        const int rowi = gi / runs_per_row;
        const int x0 = (gi - rowi * runs_per_row) * PXW;
        const int yy = rowi % h;
        const int sy = yy + r - 1;
        const bool row_ok = sy >= 0 && sy < h;
        const bf16* xr = x + ((long long)(rowi + (r - 1)) * wd) * 8;
        const bf16* gr = dy + ((long long)rowi * wd) * 8;
        float f0[CI], f1[CI], f2[CI];                        // sliding window: x at columns q - 1, q, q + 1
        {
            float t[8];
            load8(xr + (long long)(x0 - 1) * 8, row_ok && x0 - 1 >= 0, t);
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) f0[ci] = t[ci];
            load8(xr + (long long)x0 * 8, row_ok && x0 < wd, t);
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) f1[ci] = t[ci];
        }
#pragma unroll
        for (int q = 0; q < PXW; ++q) {
            const int xc = x0 + q;
            float t[8];
            load8(xr + (long long)(xc + 1) * 8, row_ok && xc + 1 < wd, t);
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) f2[ci] = t[ci];
            float g[8];
            load8(gr + (long long)xc * 8, xc < wd, g);
#pragma unroll
            for (int co = 0; co < CO; ++co) accb[co] += g[co];
#pragma unroll
            for (int ci = 0; ci < CI; ++ci)
#pragma unroll
                for (int co = 0; co < CO; ++co) {
                    acc[(0 * CI + ci) * CO + co] = fmaf(f0[ci], g[co], acc[(0 * CI + ci) * CO + co]);
                    acc[(1 * CI + ci) * CO + co] = fmaf(f1[ci], g[co], acc[(1 * CI + ci) * CO + co]);
                    acc[(2 * CI + ci) * CO + co] = fmaf(f2[ci], g[co], acc[(2 * CI + ci) * CO + co]);
                }
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) { f0[ci] = f1[ci]; f1[ci] = f2[ci]; }
        }
    }
    // warp tree (shuffles), then the four warps' results through shared memory
    __shared__ float red[4][NW + CO];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[wp][i] = v;
    }
#pragma unroll
    for (int i = 0; i < CO; ++i) {
        const float v = warp_sum(accb[i]);
        if (lane == 0) red[wp][NW + i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NW + CO; i += blockDim.x) {
        const float v = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
        if (i < NW) {
            const int co = i % CO, ci = (i / CO) % CI, s = i / (CO * CI);
            if (co < w_cout && ci < w_cin) atomicAdd(dw + ((long long)co * w_cin + ci) * 9 + r * 3 + s, v);
        } else if (db != nullptr && r == 0 && i - NW < w_cout) {
            atomicAdd(db + (i - NW), v);
        }
    }
}

template <int CI, int CO, bool DGRAD>
static int launch_tiny(const void* x, const float* w, const float* bias, void* y, int n, int h, int wd, int w_cout, int w_cin, int act,
                       float slope, cudaStream_t st) {
    constexpr int PX = (CI * CO <= 16) ? 4 : 2;
    const long long total = (long long)n * h * ((wd + PX - 1) / PX);
    SSG_CHECK_ARG(total < (1ll << 31), "conv3x3_tiny: tensor too large for 32-bit group indexing");
    conv3x3_tiny_kernel<CI, CO, DGRAD, PX><<<grid_for(total, 128, 16), 128, 0, st>>>((const bf16*)x, w, bias, (bf16*)y, n, h, wd, w_cout, w_cin, act, slope);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

static inline int round48(int c) { return c <= 4 ? 4 : 8; }

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_conv3x3_tiny_supported(int cin_stored, int cout_stored, int cin, int cout) {
    return (cin_stored == 8 && cout_stored == 8 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8) ? 1 : 0;
}

int ssg_conv3x3_tiny_fwd(const void* x, const float* w_oihw, const float* bias, void* y, int n, int h, int w, int cin, int cout, int act,
                         float slope, ssg_stream_t s) {
    SSG_CHECK_ARG(x && w_oihw && y && n > 0 && h > 0 && w > 0 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8, "conv3x3_tiny_fwd: bad args");
    const int ci = round48(cin), co = round48(cout);
    cudaStream_t st = (cudaStream_t)s;
    if (ci == 4 && co == 4) return launch_tiny<4, 4, false>(x, w_oihw, bias, y, n, h, w, cout, cin, act, slope, st);
    if (ci == 4 && co == 8) return launch_tiny<4, 8, false>(x, w_oihw, bias, y, n, h, w, cout, cin, act, slope, st);
    if (ci == 8 && co == 4) return launch_tiny<8, 4, false>(x, w_oihw, bias, y, n, h, w, cout, cin, act, slope, st);
    return launch_tiny<8, 8, false>(x, w_oihw, bias, y, n, h, w, cout, cin, act, slope, st);
}

int ssg_conv3x3_tiny_dgrad(const void* dy, const float* w_oihw, void* dx, int n, int h, int w, int cin, int cout, ssg_stream_t s) {
    SSG_CHECK_ARG(dy && w_oihw && dx && n > 0 && h > 0 && w > 0 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8, "conv3x3_tiny_dgrad: bad args");
    const int ci = round48(cout), co = round48(cin);      // roles swap: dy (cout channels) is read, dx (cin channels) written
    cudaStream_t st = (cudaStream_t)s;
    if (ci == 4 && co == 4) return launch_tiny<4, 4, true>(dy, w_oihw, nullptr, dx, n, h, w, cout, cin, SSG_ACT_NONE, 0.f, st);
    if (ci == 4 && co == 8) return launch_tiny<4, 8, true>(dy, w_oihw, nullptr, dx, n, h, w, cout, cin, SSG_ACT_NONE, 0.f, st);
    if (ci == 8 && co == 4) return launch_tiny<8, 4, true>(dy, w_oihw, nullptr, dx, n, h, w, cout, cin, SSG_ACT_NONE, 0.f, st);
    return launch_tiny<8, 8, true>(dy, w_oihw, nullptr, dx, n, h, w, cout, cin, SSG_ACT_NONE, 0.f, st);
}

int ssg_conv3x3_tiny_wgrad(const void* x, const void* dy, float* dw_oihw, float* db, int n, int h, int w, int cin, int cout, ssg_stream_t s) {
    SSG_CHECK_ARG(x && dy && dw_oihw && n > 0 && h > 0 && w > 0 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8, "conv3x3_tiny_wgrad: bad args");
    const int ci = round48(cin), co = round48(cout);
    const long long total = (long long)n * h * ((w + 7) / 8);                  // 8-pixel runs
    SSG_CHECK_ARG(total < (1ll << 31), "conv3x3_tiny_wgrad: tensor too large for 32-bit run indexing");
    const dim3 g(grid_for(total, 128 * 4, 4), 3);          // >= 4 runs (32 pixels) per thread: the block reduction stays a small share
    cudaStream_t st = (cudaStream_t)s;
    if (ci == 4 && co == 4) conv3x3_tiny_wgrad_kernel<4, 4><<<g, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, dw_oihw, db, n, h, w, cout, cin);
    else if (ci == 4 && co == 8) conv3x3_tiny_wgrad_kernel<4, 8><<<g, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, dw_oihw, db, n, h, w, cout, cin);
    else if (ci == 8 && co == 4) conv3x3_tiny_wgrad_kernel<8, 4><<<g, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, dw_oihw, db, n, h, w, cout, cin);
    else conv3x3_tiny_wgrad_kernel<8, 8><<<g, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, dw_oihw, db, n, h, w, cout, cin);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
