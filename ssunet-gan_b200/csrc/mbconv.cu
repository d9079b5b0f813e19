// Bandwidth-bound kernels of the EfficientNet encoder (efficientnet_pytorch/model.py:18-99) and of xResidualBlock
// (xresidualblock.py:9-33): depthwise convolution (any odd/even k, stride 1/2, TF-style asymmetric "same" padding),
// squeeze-and-excitation (pool -> gate MLP -> scale, 3 launches per direction), swish, the Gaussian gate, zero padding
// and bilinear resize.  NHWC storage; every thread owns one pixel x one 16-byte channel vector so that warps read and
// write full 128-byte lines along C.
#include "common.cuh"

namespace ssg {

constexpr int DW_THREADS = 256;
constexpr int DW_LANES = 8;                      // channel vectors per block row
constexpr int DW_PIX = DW_THREADS / DW_LANES;    // pixels per block iteration

__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float swish_gradf_(float x) {
    const float s = sigmoidf_(x);
    return s * (1.f + x * (1.f - s));            // utils.py:47-48 (SwishImplementation.backward)
}

// ---------------------------------------------------------------------------------------------------------------------
// depthwise convolution.  Weights arrive in the parameter's own layout [C][1][k][k] fp32; each block stages the slice of
// its channel chunk transposed ([tap][chunk]) in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void dw_stage_weights(float* sw, const float* __restrict__ w, int C, int kk, int c_base, int chunk) {
    for (int i = threadIdx.x; i < kk * chunk; i += blockDim.x) {
        const int cl = i / kk, tap = i - cl * kk;          // consecutive threads walk one channel's taps: coalesced reads
        const int c = c_base + cl;
        sw[tap * chunk + cl] = c < C ? w[(long long)c * kk + tap] : 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(DW_THREADS) dwconv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, T* __restrict__ y, int N, int H,
                                                                int W, int C, int k, int stride, int pad_t, int pad_l, int OH,
                                                                int OW) {
    constexpr int V = Vec<T>::N;
    constexpr int CHUNK = DW_LANES * V;
    extern __shared__ float sw[];
    const int kk = k * k;
    const int c_base = blockIdx.y * CHUNK;
    dw_stage_weights<T>(sw, w, C, kk, c_base, CHUNK);
    __syncthreads();
    const int lane = threadIdx.x % DW_LANES, prow = threadIdx.x / DW_LANES;
    const int c0 = c_base + lane * V;
    if (c0 >= C) return;
    const long long pixels = (long long)N * OH * OW;
    float fb[V];
#pragma unroll
    for (int j = 0; j < V; ++j) fb[j] = bias ? bias[c0 + j] : 0.f;
    for (long long p = (long long)blockIdx.x * DW_PIX + prow; p < pixels; p += (long long)gridDim.x * DW_PIX) {
        const int ox = (int)(p % OW);
        const int oy = (int)((p / OW) % OH);
        const int n = (int)(p / ((long long)OW * OH));
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fb[j];
        const int iy0 = oy * stride - pad_t, ix0 = ox * stride - pad_l;
        for (int ky = 0; ky < k; ++ky) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= H) continue;
            const T* xrow = x + ((long long)(n * H + iy) * W) * C + c0;
            for (int kx = 0; kx < k; ++kx) {
                const int ix = ix0 + kx;
                if (ix < 0 || ix >= W) continue;
                Vec<T> vx; vx.load(xrow + (long long)ix * C);
                float fx[V]; vx.get(fx);
                const float* wt = sw + (ky * k + kx) * CHUNK + lane * V;
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(fx[j], wt[j], acc[j]);
            }
        }
        Vec<T> vo; vo.set(acc); vo.store(y + p * C + c0);
    }
}

// Sliding-window forward for the kernel sizes / strides the encoder uses: a thread produces DW_OW consecutive output
// pixels of one row for its 16-byte channel vector, so each input vector it loads feeds up to K / S outputs
// (3x3 s1: 4.5 loads per output instead of 9; 5x5 s1: 10 instead of 25).  FLIP stages the weights mirrored, which turns
// the same kernel into the stride-1 data gradient (dx = dy (*) flipped w with pad' = K - 1 - pad).
constexpr int DW_OW = 4;

template <typename T, int K, int S, bool FLIP>
__global__ void __launch_bounds__(DW_THREADS) dwconv_slide_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, T* __restrict__ y, int N, int H,
                                                                  int W, int C, int pad_t, int pad_l, int OH, int OW) {
    constexpr int V = Vec<T>::N;
    constexpr int CHUNK = DW_LANES * V;
    constexpr int COLS = (DW_OW - 1) * S + K;           // input columns one thread touches per row
    extern __shared__ float sw[];
    const int c_base = blockIdx.y * CHUNK;
    for (int i = threadIdx.x; i < K * K * CHUNK; i += blockDim.x) {
        const int cl = i / (K * K), tap = i - cl * (K * K);
        const int c = c_base + cl;
        sw[(FLIP ? K * K - 1 - tap : tap) * CHUNK + cl] = c < C ? w[(long long)c * K * K + tap] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x % DW_LANES, prow = threadIdx.x / DW_LANES;
    const int c0 = c_base + lane * V;
    if (c0 >= C) return;
    const int groups_x = (OW + DW_OW - 1) / DW_OW;
    const long long groups = (long long)N * OH * groups_x;
    float fb[V];
#pragma unroll
    for (int j = 0; j < V; ++j) fb[j] = bias ? bias[c0 + j] : 0.f;
    for (long long gidx = (long long)blockIdx.x * DW_PIX + prow; gidx < groups; gidx += (long long)gridDim.x * DW_PIX) {
        const int gx = (int)(gidx % groups_x);
        const int oy = (int)((gidx / groups_x) % OH);
        const int n = (int)(gidx / ((long long)groups_x * OH));
        const int ox0 = gx * DW_OW;
        float acc[DW_OW][V];
#pragma unroll
        for (int o = 0; o < DW_OW; ++o)
#pragma unroll
            for (int j = 0; j < V; ++j) acc[o][j] = fb[j];
        const int iy0 = oy * S - pad_t, ix0 = ox0 * S - pad_l;
#pragma unroll(K <= 3 ? K : 1)
        for (int ky = 0; ky < K; ++ky) {            // wide filters: one row of loads in flight at a time (register budget)
            const int iy = iy0 + ky;
            const bool row_ok = iy >= 0 && iy < H;
            const T* xrow = x + ((long long)(n * H + (row_ok ? iy : 0)) * W) * C + c0;
            // branch-free row: every load goes to a clamped (always valid) address and is zeroed by a select when the tap
            // falls outside the image, so the COLS loads of the row are independent and can all be in flight
            Vec<T> vx[COLS];
#pragma unroll
            for (int jx = 0; jx < COLS; ++jx) {
                const int ix = ix0 + jx;
                vx[jx].load(xrow + (long long)(ix < 0 ? 0 : (ix >= W ? W - 1 : ix)) * C);
            }
#pragma unroll
            for (int jx = 0; jx < COLS; ++jx) {
                const int ix = ix0 + jx;
                const bool keep = row_ok && ix >= 0 && ix < W;
                float fx[V]; vx[jx].get(fx);
#pragma unroll
                for (int j = 0; j < V; ++j) fx[j] = keep ? fx[j] : 0.f;
#pragma unroll
                for (int o = 0; o < DW_OW; ++o) {
                    const int kx = jx - o * S;              // compile-time after unrolling
                    if (kx < 0 || kx >= K) continue;
                    const float* wt = sw + (ky * K + kx) * CHUNK + lane * V;
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[o][j] = fmaf(fx[j], wt[j], acc[o][j]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < DW_OW; ++o) {
            if (ox0 + o >= OW) break;
            Vec<T> vo; vo.set(acc[o]);
            vo.store(y + (((long long)n * OH + oy) * OW + ox0 + o) * C + c0);
        }
    }
}

template <typename T, int K, int S, bool FLIP>
static int launch_dw_slide(const T* x, const float* w, const float* bias, T* y, int n, int h, int w_, int c, int pad_t, int pad_l, int oh,
                           int ow, cudaStream_t st) {
    constexpr int CHUNK = DW_LANES * Vec<T>::N;
    const long long groups = (long long)n * oh * ((ow + DW_OW - 1) / DW_OW);
    // few fat blocks: every block first stages its weight slice, so it must amortise that over many pixel groups
    const int chunks = (c + CHUNK - 1) / CHUNK;
    dim3 grid(grid_for(groups, DW_PIX * 2, chunks >= 4 ? 1 : (chunks >= 2 ? 2 : 4)), chunks);
    dwconv_slide_kernel<T, K, S, FLIP><<<grid, DW_THREADS, sizeof(float) * K * K * CHUNK, st>>>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// returns SSG_ERR_UNSUPPORTED when no sliding-window instance covers (k, stride): the caller falls back to the generic kernel
template <typename T, bool FLIP>
static int dispatch_dw_slide(const T* x, const float* w, const float* bias, T* y, int n, int h, int w_, int c, int k, int stride, int pad_t,
                             int pad_l, int oh, int ow, cudaStream_t st) {
    if (k == 3 && stride == 1) return launch_dw_slide<T, 3, 1, FLIP>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow, st);
    if (k == 5 && stride == 1) return launch_dw_slide<T, 5, 1, FLIP>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow, st);
    if (k == 9 && stride == 1) return launch_dw_slide<T, 9, 1, FLIP>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow, st);
    if (!FLIP && k == 3 && stride == 2) return launch_dw_slide<T, 3, 2, false>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow, st);
    if (!FLIP && k == 5 && stride == 2) return launch_dw_slide<T, 5, 2, false>(x, w, bias, y, n, h, w_, c, pad_t, pad_l, oh, ow, st);
    return SSG_ERR_UNSUPPORTED;
}

// dx[n,iy,ix,c] = sum_taps dy[n,(iy+pad_t-ky)/s,(ix+pad_l-kx)/s,c] * w[c,ky,kx]   (only where the division is exact)
template <typename T>
__global__ void __launch_bounds__(DW_THREADS) dwconv_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w,
                                                                  T* __restrict__ dx, int N, int H, int W, int C, int k, int stride,
                                                                  int pad_t, int pad_l, int OH, int OW) {
    constexpr int V = Vec<T>::N;
    constexpr int CHUNK = DW_LANES * V;
    extern __shared__ float sw[];
    const int kk = k * k;
    const int c_base = blockIdx.y * CHUNK;
    dw_stage_weights<T>(sw, w, C, kk, c_base, CHUNK);
    __syncthreads();
    const int lane = threadIdx.x % DW_LANES, prow = threadIdx.x / DW_LANES;
    const int c0 = c_base + lane * V;
    if (c0 >= C) return;
    const long long pixels = (long long)N * H * W;
    for (long long p = (long long)blockIdx.x * DW_PIX + prow; p < pixels; p += (long long)gridDim.x * DW_PIX) {
        const int ix = (int)(p % W);
        const int iy = (int)((p / W) % H);
        const int n = (int)(p / ((long long)W * H));
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        for (int ky = 0; ky < k; ++ky) {
            const int ty = iy + pad_t - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= OH) continue;
            const T* drow = dy + ((long long)(n * OH + oy) * OW) * C + c0;
            for (int kx = 0; kx < k; ++kx) {
                const int tx = ix + pad_l - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox >= OW) continue;
                Vec<T> vd; vd.load(drow + (long long)ox * C);
                float fd[V]; vd.get(fd);
                const float* wt = sw + (ky * k + kx) * CHUNK + lane * V;
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(fd[j], wt[j], acc[j]);
            }
        }
        Vec<T> vo; vo.set(acc); vo.store(dx + p * C + c0);
    }
}

// dw[c,ky,kx] = sum_{n,oy,ox} dy[n,oy,ox,c] * x[n, oy*s - pad_t + ky, ox*s - pad_l + kx, c]
// grid (pixel slabs, channel chunks); tap-outer loop (the slab's dy / x lines stay in L1/L2 across taps); the DW_PIX
// partial sums of a block are folded in shared memory and added to dw (pre-zeroed) with one fp32 atomic per (c, tap).
template <typename T>
__global__ void __launch_bounds__(DW_THREADS) dwconv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  float* __restrict__ dw, int N, int H, int W, int C, int k, int stride,
                                                                  int pad_t, int pad_l, int OH, int OW, long long pix_per_block) {
    constexpr int V = Vec<T>::N;
    constexpr int CHUNK = DW_LANES * V;
    __shared__ float red[DW_PIX][CHUNK + 1];
    const int lane = threadIdx.x % DW_LANES, prow = threadIdx.x / DW_LANES;
    const int c0 = blockIdx.y * CHUNK + lane * V;
    const bool live = c0 < C;
    const long long pixels = (long long)N * OH * OW;
    const long long p_lo = (long long)blockIdx.x * pix_per_block;
    const long long p_hi = p_lo + pix_per_block < pixels ? p_lo + pix_per_block : pixels;
    for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
            float acc[V];
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = 0.f;
            if (live)
                for (long long p = p_lo + prow; p < p_hi; p += DW_PIX) {
                    const int ox = (int)(p % OW);
                    const int oy = (int)((p / OW) % OH);
                    const int n = (int)(p / ((long long)OW * OH));
                    const int iy = oy * stride - pad_t + ky, ix = ox * stride - pad_l + kx;
                    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
                    Vec<T> vd; vd.load(dy + p * C + c0);
                    Vec<T> vx; vx.load(x + ((long long)(n * H + iy) * W + ix) * C + c0);
                    float fd[V], fx[V]; vd.get(fd); vx.get(fx);
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] = fmaf(fd[j], fx[j], acc[j]);
                }
#pragma unroll
            for (int j = 0; j < V; ++j) red[prow][lane * V + j] = acc[j];
            __syncthreads();
            if (threadIdx.x < CHUNK) {
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < DW_PIX; ++r) s += red[r][threadIdx.x];
                const int c = blockIdx.y * CHUNK + threadIdx.x;
                if (c < C) atomicAdd(dw + (long long)c * k * k + ky * k + kx, s);
            }
            __syncthreads();
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// squeeze-and-excitation (model.py:78-82): pooled = mean_hw x; gate = sigmoid(W2 swish(W1 pooled + b1) + b2); y = x * gate
// ---------------------------------------------------------------------------------------------------------------------
// sums[n][c] += sum over this block's pixel slab of a[n,p,c] * (b ? b[n,p,c] : 1)   (fp32, pre-zeroed)
template <typename T, bool PROD>
__global__ void __launch_bounds__(DW_THREADS) plane_sum_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ sums,
                                                               int HW, int C, int pix_per_block) {
    constexpr int V = Vec<T>::N;
    constexpr int CHUNK = DW_LANES * V;
    __shared__ float red[DW_PIX][CHUNK + 1];
    const int lane = threadIdx.x % DW_LANES, prow = threadIdx.x / DW_LANES;
    const int n = blockIdx.z;
    const int c0 = blockIdx.y * CHUNK + lane * V;
    const int p_lo = blockIdx.x * pix_per_block;
    const int p_hi = p_lo + pix_per_block < HW ? p_lo + pix_per_block : HW;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    if (c0 < C)
        for (int p = p_lo + prow; p < p_hi; p += DW_PIX) {
            const long long off = ((long long)n * HW + p) * C + c0;
            Vec<T> va; va.load(a + off);
            float fa[V]; va.get(fa);
            if (PROD) {
                Vec<T> vb; vb.load(b + off);
                float fb[V]; vb.get(fb);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(fa[j], fb[j], acc[j]);
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] += fa[j];
            }
        }
#pragma unroll
    for (int j = 0; j < V; ++j) red[prow][lane * V + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < CHUNK) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < DW_PIX; ++r) s += red[r][threadIdx.x];
        const int c = blockIdx.y * CHUNK + threadIdx.x;
        if (c < C) atomicAdd(sums + (long long)n * C + c, s);
    }
}

// one block per sample.  pooled_sum holds sum_hw x (divided by HW here).  Writes s_pre [N][S] and gate [N][C].
__global__ void __launch_bounds__(256) se_gate_fwd_kernel(const float* __restrict__ pooled_sum, float inv_hw, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, int C, int S, float* __restrict__ pooled,
                                                           float* __restrict__ s_pre, float* __restrict__ gate) {
    extern __shared__ float sm[];
    float* p = sm;            // [C]
    float* s = sm + C;        // [S]
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = pooled_sum[(long long)n * C + c] * inv_hw;
        p[c] = v;
        pooled[(long long)n * C + c] = v;
    }
    __syncthreads();
    for (int j = warp; j < S; j += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(w1[(long long)j * C + c], p[c], a);
        a = warp_sum(a);
        if (lane == 0) {
            a += b1[j];
            s_pre[(long long)n * S + j] = a;
            s[j] = swishf_(a);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = b2[c];
        for (int j = 0; j < S; ++j) a = fmaf(w2[(long long)c * S + j], s[j], a);
        gate[(long long)n * C + c] = sigmoidf_(a);
    }
}

// one block per sample: dgate [N][C] (= sum_hw dy*x) -> dpooled [N][C] (already divided by HW), and the MLP's parameter
// gradients accumulated over samples with fp32 atomics (dw1/db1/dw2/db2 pre-zeroed).
__global__ void __launch_bounds__(256) se_gate_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                           const float* __restrict__ s_pre, const float* __restrict__ pooled,
                                                           const float* __restrict__ w1, const float* __restrict__ w2, int C, int S,
                                                           float inv_hw, float* __restrict__ dpooled, float* __restrict__ dw1,
                                                           float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
    extern __shared__ float sm[];
    float* de = sm;               // [C]  dL/d(pre-sigmoid)
    float* sv = sm + C;           // [S]  swish(s_pre)
    float* dsp = sm + C + S;      // [S]  dL/d(s_pre)
    float* p = sm + C + 2 * S;    // [C]
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float g = gate[(long long)n * C + c];
        const float d = dgate[(long long)n * C + c] * g * (1.f - g);
        de[c] = d;
        p[c] = pooled[(long long)n * C + c];
        atomicAdd(db2 + c, d);
    }
    for (int j = threadIdx.x; j < S; j += blockDim.x) sv[j] = swishf_(s_pre[(long long)n * S + j]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float d = de[c];
        for (int j = 0; j < S; ++j) atomicAdd(dw2 + (long long)c * S + j, d * sv[j]);
    }
    for (int j = warp; j < S; j += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(w2[(long long)c * S + j], de[c], a);
        a = warp_sum(a);
        if (lane == 0) {
            const float d = a * swish_gradf_(s_pre[(long long)n * S + j]);
            dsp[j] = d;
            atomicAdd(db1 + j, d);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        const float pc = p[c];
        for (int j = 0; j < S; ++j) {
            const float d = dsp[j];
            a = fmaf(w1[(long long)j * C + c], d, a);
            atomicAdd(dw1 + (long long)j * C + c, d * pc);
        }
        dpooled[(long long)n * C + c] = a * inv_hw;
    }
}

// y[n,p,c] = a[n,p,c] * mul[n][c] + (add ? add[n][c] : 0): the SE scale (forward), and dx = dy*gate + dpooled/HW (backward)
template <typename T>
__global__ void __launch_bounds__(256) plane_scale_kernel(const T* __restrict__ a, const float* __restrict__ mul, const float* __restrict__ add,
                                                           T* __restrict__ y, long long HW, int C, long long total_vec) {
    constexpr int V = Vec<T>::N;
    const int vpr = C / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
        const long long row = i / vpr;
        const int c0 = (int)(i - row * vpr) * V;
        const long long n = row / HW;
        Vec<T> va; va.load(a + i * V);
        float f[V]; va.get(f);
        const float* m = mul + n * C + c0;
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] *= m[j];
        if (add) {
            const float* ad = add + n * C + c0;
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] += ad[j];
        }
        Vec<T> vo; vo.set(f); vo.store(y + i * V);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// elementwise: swish (utils.py:36-48), Gaussian gate x1 * exp(-z^2) (xresidualblock.py:4-6,19-23)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int OP>   // OP 0: swish fwd (a = x); 1: swish bwd (a = dy, b = x)
__global__ void __launch_bounds__(256) swish_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long long n) {
    constexpr int V = Vec<T>::N;
    const long long nv = n / V, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec<T> va; va.load(a + i * V);
        float fa[V]; va.get(fa);
        if (OP == 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) fa[j] = swishf_(fa[j]);
        } else {
            Vec<T> vb; vb.load(b + i * V);
            float fb[V]; vb.get(fb);
#pragma unroll
            for (int j = 0; j < V; ++j) fa[j] *= swish_gradf_(fb[j]);
        }
        Vec<T> vo; vo.set(fa); vo.store(o + i * V);
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        o[i] = from_f<T>(OP == 0 ? swishf_(to_f(a[i])) : to_f(a[i]) * swish_gradf_(to_f(b[i])));
}

template <typename T>
__global__ void __launch_bounds__(256) gauss_gate_fwd_kernel(const T* __restrict__ x1, const T* __restrict__ z, T* __restrict__ y, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float zz = to_f(z[i]);
        y[i] = from_f<T>(to_f(x1[i]) * expf(-zz * zz));
    }
}
template <typename T>
__global__ void __launch_bounds__(256) gauss_gate_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x1, const T* __restrict__ z,
                                                              T* __restrict__ dx1, T* __restrict__ dz, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float zz = to_f(z[i]), d = to_f(dy[i]), e = expf(-zz * zz);
        dx1[i] = from_f<T>(d * e);
        dz[i] = from_f<T>(d * to_f(x1[i]) * e * (-2.f * zz));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// nn.ZeroPad2d / crop (utils.py:133-136): y[n,oy,ox,:] = x[n,oy-pad_t,ox-pad_l,:] or 0 (negative pads crop)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pad2d_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int pad_t,
                                                     int pad_l, int OH, int OW) {
    const long long total = (long long)N * OH * OW * C, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int ox = (int)(p % OW), oy = (int)((p / OW) % OH), n = (int)(p / ((long long)OW * OH));
        const int iy = oy - pad_t, ix = ox - pad_l;
        y[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[((long long)(n * H + iy) * W + ix) * C + c] : from_f<T>(0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// F.interpolate(mode='bilinear', align_corners=False) to an arbitrary size (archs.py:459)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, float scale, int in, int& i0, int& i1, float& l1) {
    float src = ((float)o + 0.5f) * scale - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                                                                   int OH, int OW, float sh, float sw) {
    const long long total = (long long)N * OH * OW * C, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int ox = (int)(p % OW), oy = (int)((p / OW) % OH), n = (int)(p / ((long long)OW * OH));
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src(oy, sh, H, y0, y1, ly);
        bilinear_src(ox, sw, W, x0, x1, lx);
        const T* b = x + (long long)n * H * W * C + c;
        const float v00 = to_f(b[((long long)y0 * W + x0) * C]), v01 = to_f(b[((long long)y0 * W + x1) * C]);
        const float v10 = to_f(b[((long long)y1 * W + x0) * C]), v11 = to_f(b[((long long)y1 * W + x1) * C]);
        y[i] = from_f<T>((1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11));
    }
}

// adjoint: scatter with fp32 atomics into dx32 (pre-zeroed)
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx32, int N, int H, int W,
                                                                   int C, int OH, int OW, float sh, float sw) {
    const long long total = (long long)N * OH * OW * C, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int ox = (int)(p % OW), oy = (int)((p / OW) % OH), n = (int)(p / ((long long)OW * OH));
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src(oy, sh, H, y0, y1, ly);
        bilinear_src(ox, sw, W, x0, x1, lx);
        const float d = to_f(dy[i]);
        float* b = dx32 + (long long)n * H * W * C + c;
        atomicAdd(b + ((long long)y0 * W + x0) * C, d * (1.f - ly) * (1.f - lx));
        atomicAdd(b + ((long long)y0 * W + x1) * C, d * (1.f - ly) * lx);
        atomicAdd(b + ((long long)y1 * W + x0) * C, d * ly * (1.f - lx));
        atomicAdd(b + ((long long)y1 * W + x1) * C, d * ly * lx);
    }
}

template <typename T>
static int dw_geometry_ok(int n, int h, int w, int c, int k, int stride, int pad_t, int pad_l, int oh, int ow) {
    constexpr int V = Vec<T>::N;
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "dwconv: empty tensor");
    SSG_CHECK_ARG(c % V == 0, "dwconv: C=%d must be a multiple of %d (16-byte channel vectors)", c, V);
    SSG_CHECK_ARG(k >= 1 && k <= 11 && (stride == 1 || stride == 2), "dwconv: k=%d stride=%d unsupported", k, stride);
    SSG_CHECK_ARG(pad_t >= 0 && pad_l >= 0 && pad_t < k && pad_l < k, "dwconv: bad padding");
    SSG_CHECK_ARG((oh - 1) * stride - pad_t < h && (ow - 1) * stride - pad_l < w, "dwconv: output larger than the padded input");
    return SSG_OK;
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_dwconv2d_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int n, int h, int w_, int c, int k, int stride,
                     int pad_t, int pad_l, int oh, int ow, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, {
        int rc = dw_geometry_ok<T>(n, h, w_, c, k, stride, pad_t, pad_l, oh, ow);
        if (rc) return rc;
        rc = dispatch_dw_slide<T, false>((const T*)x, w, bias, (T*)y, n, h, w_, c, k, stride, pad_t, pad_l, oh, ow, (cudaStream_t)s);
        if (rc != SSG_ERR_UNSUPPORTED) return rc;
        constexpr int CHUNK = DW_LANES * Vec<T>::N;
        dim3 grid(grid_for((long long)n * oh * ow, DW_PIX * 4, 4), (c + CHUNK - 1) / CHUNK);
        dwconv_fwd_kernel<T><<<grid, DW_THREADS, sizeof(float) * k * k * CHUNK, (cudaStream_t)s>>>(
            (const T*)x, w, bias, (T*)y, n, h, w_, c, k, stride, pad_t, pad_l, oh, ow);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_dwconv2d_dgrad(const void* dy, const float* w, void* dx, int dtype, int n, int h, int w_, int c, int k, int stride, int pad_t,
                       int pad_l, int oh, int ow, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, {
        int rc = dw_geometry_ok<T>(n, h, w_, c, k, stride, pad_t, pad_l, oh, ow);
        if (rc) return rc;
        if (stride == 1) {      // dx = dy (*) mirrored weights, leading pad k - 1 - pad: the forward kernel with FLIP
            rc = dispatch_dw_slide<T, true>((const T*)dy, w, nullptr, (T*)dx, n, oh, ow, c, k, 1, k - 1 - pad_t, k - 1 - pad_l, h, w_, (cudaStream_t)s);
            if (rc != SSG_ERR_UNSUPPORTED) return rc;
        }
        constexpr int CHUNK = DW_LANES * Vec<T>::N;
        dim3 grid(grid_for((long long)n * h * w_, DW_PIX * 4, 4), (c + CHUNK - 1) / CHUNK);
        dwconv_dgrad_kernel<T><<<grid, DW_THREADS, sizeof(float) * k * k * CHUNK, (cudaStream_t)s>>>(
            (const T*)dy, w, (T*)dx, n, h, w_, c, k, stride, pad_t, pad_l, oh, ow);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_dwconv2d_wgrad(const void* x, const void* dy, float* dw, int dtype, int n, int h, int w_, int c, int k, int stride, int pad_t,
                       int pad_l, int oh, int ow, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, {
        int rc = dw_geometry_ok<T>(n, h, w_, c, k, stride, pad_t, pad_l, oh, ow);
        if (rc) return rc;
        constexpr int CHUNK = DW_LANES * Vec<T>::N;
        SSG_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)c * k * k, (cudaStream_t)s));
        const long long pixels = (long long)n * oh * ow;
        const int chunks = (c + CHUNK - 1) / CHUNK;
        long long slabs = ((long long)sm_count_cached() * 4 + chunks - 1) / chunks;
        long long ppb = (pixels + slabs - 1) / slabs;
        if (ppb < 4 * DW_PIX) ppb = 4 * DW_PIX;
        slabs = (pixels + ppb - 1) / ppb;
        dim3 grid((unsigned)slabs, chunks);
        dwconv_wgrad_kernel<T><<<grid, DW_THREADS, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, dw, n, h, w_, c, k, stride, pad_t,
                                                                         pad_l, oh, ow, ppb);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_plane_sums(const void* a, const void* b, float* sums, int dtype, int n, int hw, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && hw > 0 && c > 0, "plane_sums: empty tensor");
    SSG_DISPATCH_DTYPE(dtype, {
        constexpr int CHUNK = DW_LANES * Vec<T>::N;
        SSG_CHECK_ARG(c % Vec<T>::N == 0, "plane_sums: C=%d must be a multiple of %d", c, Vec<T>::N);
        SSG_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)n * c, (cudaStream_t)s));
        const int chunks = (c + CHUNK - 1) / CHUNK;
        long long slabs = ((long long)sm_count_cached() * 4 + (long long)chunks * n - 1) / ((long long)chunks * n);
        int ppb = (int)((hw + slabs - 1) / slabs);
        if (ppb < 2 * DW_PIX) ppb = 2 * DW_PIX;
        slabs = (hw + ppb - 1) / ppb;
        dim3 grid((unsigned)slabs, chunks, n);
        if (b)
            plane_sum_kernel<T, true><<<grid, DW_THREADS, 0, (cudaStream_t)s>>>((const T*)a, (const T*)b, sums, hw, c, ppb);
        else
            plane_sum_kernel<T, false><<<grid, DW_THREADS, 0, (cudaStream_t)s>>>((const T*)a, nullptr, sums, hw, c, ppb);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_se_gate_fwd(const float* pooled_sum, int n, int hw, int c, int sq, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* pooled, float* s_pre, float* gate, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && c > 0 && sq > 0 && (size_t)(c + sq) * 4 <= 48 * 1024, "se_gate: C=%d S=%d unsupported", c, sq);
    se_gate_fwd_kernel<<<n, 256, sizeof(float) * (c + sq), (cudaStream_t)s>>>(pooled_sum, 1.0f / (float)hw, w1, b1, w2, b2, c, sq, pooled,
                                                                             s_pre, gate);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_se_gate_bwd(const float* dgate, const float* gate, const float* s_pre, const float* pooled, const float* w1, const float* w2,
                    int n, int hw, int c, int sq, float* dpooled, float* dw1, float* db1, float* dw2, float* db2, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && c > 0 && sq > 0 && (size_t)(2 * c + 2 * sq) * 4 <= 48 * 1024, "se_gate_bwd: C=%d S=%d unsupported", c, sq);
    cudaStream_t st = (cudaStream_t)s;
    SSG_CHECK_CUDA(cudaMemsetAsync(dw1, 0, sizeof(float) * (size_t)c * sq, st));
    SSG_CHECK_CUDA(cudaMemsetAsync(dw2, 0, sizeof(float) * (size_t)c * sq, st));
    SSG_CHECK_CUDA(cudaMemsetAsync(db1, 0, sizeof(float) * (size_t)sq, st));
    SSG_CHECK_CUDA(cudaMemsetAsync(db2, 0, sizeof(float) * (size_t)c, st));
    se_gate_bwd_kernel<<<n, 256, sizeof(float) * (2 * c + 2 * sq), st>>>(dgate, gate, s_pre, pooled, w1, w2, c, sq, 1.0f / (float)hw,
                                                                          dpooled, dw1, db1, dw2, db2);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_plane_scale(const void* a, const float* mul, const float* add, void* y, int dtype, int n, int hw, int c, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && hw > 0 && c > 0, "plane_scale: empty tensor");
    SSG_DISPATCH_DTYPE(dtype, {
        SSG_CHECK_ARG(c % Vec<T>::N == 0, "plane_scale: C=%d must be a multiple of %d", c, Vec<T>::N);
        const long long tv = (long long)n * hw * (c / Vec<T>::N);
        plane_scale_kernel<T><<<grid_for(tv, 256 * 2), 256, 0, (cudaStream_t)s>>>((const T*)a, mul, add, (T*)y, hw, c, tv);
    });
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_swish_fwd(const void* x, void* y, int dtype, long long n, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, (swish_kernel<T, 0><<<grid_for(n / Vec<T>::N + 1, 256 * 2), 256, 0, (cudaStream_t)s>>>((const T*)x, nullptr, (T*)y, n)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_swish_bwd(const void* dy, const void* x, void* dx, int dtype, long long n, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, (swish_kernel<T, 1><<<grid_for(n / Vec<T>::N + 1, 256 * 2), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (T*)dx, n)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_gauss_gate_fwd(const void* x1, const void* z, void* y, int dtype, long long n, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, (gauss_gate_fwd_kernel<T><<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)s>>>((const T*)x1, (const T*)z, (T*)y, n)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_gauss_gate_bwd(const void* dy, const void* x1, const void* z, void* dx1, void* dz, int dtype, long long n, ssg_stream_t s) {
    SSG_DISPATCH_DTYPE(dtype, (gauss_gate_bwd_kernel<T><<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x1, (const T*)z,
                                                                                                      (T*)dx1, (T*)dz, n)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_pad2d(const void* x, void* y, int dtype, int n, int h, int w, int c, int pad_t, int pad_l, int oh, int ow, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "pad2d: empty tensor");
    SSG_DISPATCH_DTYPE(dtype, (pad2d_kernel<T><<<grid_for((long long)n * oh * ow * c, 256 * 4), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, n, h, w, c,
                                                                                                                     pad_t, pad_l, oh, ow)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_resize_bilinear_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "resize: empty tensor");
    SSG_DISPATCH_DTYPE(dtype, (resize_bilinear_fwd_kernel<T><<<grid_for((long long)n * oh * ow * c, 256 * 4), 256, 0, (cudaStream_t)s>>>(
                                   (const T*)x, (T*)y, n, h, w, c, oh, ow, (float)h / (float)oh, (float)w / (float)ow)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_resize_bilinear_bwd(const void* dy, float* dx32, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s) {
    SSG_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "resize: empty tensor");
    SSG_CHECK_CUDA(cudaMemsetAsync(dx32, 0, sizeof(float) * (size_t)n * h * w * c, (cudaStream_t)s));
    SSG_DISPATCH_DTYPE(dtype, (resize_bilinear_bwd_kernel<T><<<grid_for((long long)n * oh * ow * c, 256 * 4), 256, 0, (cudaStream_t)s>>>(
                                   (const T*)dy, dx32, n, h, w, c, oh, ow, (float)h / (float)oh, (float)w / (float)ow)));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
