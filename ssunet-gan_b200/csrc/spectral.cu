// Spectral-norm power iteration on W [rows, cols] (fp32 master weights): two streaming mat-vecs,
// two normalisations, sigma.  HBM-bound: W is read twice per forward (2*|W|*4 bytes).
#include "common.cuh"

namespace ssg {

// out[c] = sum_r W[r][c] * u[r]   (coalesced over c; rows split over blockIdx.y, fp32 atomics)
__global__ void __launch_bounds__(256) gemv_t_kernel(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ out,
                                                      int rows, int cols, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (c >= cols) return;
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc = fmaf(w[(long long)r * cols + c], u[r], acc);
    atomicAdd(&out[c], acc);
}
// out[r] = sum_c W[r][c] * v[c]   (one warp per row)
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ w, const float* __restrict__ v, float* __restrict__ out,
                                                    int rows, int cols) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* wr = w + (long long)r * cols;
    float acc = 0.f;
    for (int c = lane; c < cols; c += 32) acc = fmaf(wr[c], v[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
}
// single block: dst = src / max(||src||, eps)  (F.normalize);  optionally sigma = dot(dst, src), inv_sigma[0]=1/sigma, [1]=sigma
__global__ void __launch_bounds__(1024) normalize_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, float eps,
                                                          float* __restrict__ inv_sigma) {
    __shared__ float red[32];
    __shared__ float s_norm;
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fmaf(src[i], src[i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float b = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        b = warp_sum(b);
        if (threadIdx.x == 0) s_norm = sqrtf(b);
    }
    __syncthreads();
    const float nrm = s_norm;
    const float d = fmaxf(nrm, eps);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] / d;
    if (inv_sigma && threadIdx.x == 0) {
        // sigma = u . (W v) with u = (W v)/max(||W v||, eps)  ==  ||W v||^2 / max(||W v||, eps)
        const float sigma = nrm * nrm / d;
        inv_sigma[0] = 1.f / sigma;
        inv_sigma[1] = sigma;
    }
}
// sigma = dot(u, wv) without touching u (eval mode)
__global__ void __launch_bounds__(1024) dot_sigma_kernel(const float* __restrict__ u, const float* __restrict__ wv, int n,
                                                          float* __restrict__ inv_sigma) {
    __shared__ float red[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fmaf(u[i], wv[i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        float b = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        b = warp_sum(b);
        if (threadIdx.x == 0) { inv_sigma[0] = 1.f / b; inv_sigma[1] = b; }
    }
}
__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ out) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc = fmaf(a[i], b[i], acc);
    acc = warp_sum(acc);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += (double)red[w];
        atomicAdd(out, s);
    }
}
// dW_orig = (dW_sn - dot * inv_sigma * u v^T) * inv_sigma
__global__ void __launch_bounds__(256) spectral_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ u, const float* __restrict__ v,
                                                            int rows, int cols, const float* __restrict__ inv_sigma,
                                                            const double* __restrict__ dot, float* __restrict__ out) {
    const float is = inv_sigma[0];
    const float k = (float)dot[0] * is;
    const long long n = (long long)rows * cols, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        out[i] = (dw[i] - k * u[r] * v[c]) * is;
    }
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_spectral_sigma(const float* w, float* u, float* v, int rows, int cols, float eps, int do_power_iteration, float* inv_sigma,
                       float* workspace, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && cols > 0, "spectral_sigma: bad shape");
    cudaStream_t st = (cudaStream_t)s;
    float* wv = workspace;          // [rows]
    float* wtu = workspace + rows;  // [cols]
    if (do_power_iteration) {
        SSG_CHECK_CUDA(cudaMemsetAsync(wtu, 0, sizeof(float) * cols, st));
        int cb = (cols + 255) / 256;
        int splits = (2 * sm_count_cached() + cb - 1) / cb;
        splits = splits < 1 ? 1 : (splits > rows ? rows : splits);
        int rpb = (rows + splits - 1) / splits;
        dim3 grid((unsigned)cb, (unsigned)((rows + rpb - 1) / rpb));
        gemv_t_kernel<<<grid, 256, 0, st>>>(w, u, wtu, rows, cols, rpb);
        normalize_kernel<<<1, 1024, 0, st>>>(wtu, v, cols, eps, nullptr);
        gemv_kernel<<<(unsigned)(((long long)rows * 32 + 255) / 256), 256, 0, st>>>(w, v, wv, rows, cols);
        normalize_kernel<<<1, 1024, 0, st>>>(wv, u, rows, eps, inv_sigma);
    } else {
        gemv_kernel<<<(unsigned)(((long long)rows * 32 + 255) / 256), 256, 0, st>>>(w, v, wv, rows, cols);
        dot_sigma_kernel<<<1, 1024, 0, st>>>(u, wv, rows, inv_sigma);
    }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_spectral_weight_bwd(const float* dw_sn, const float* w_orig, const float* u, const float* v, int rows, int cols,
                            const float* inv_sigma, double* dot_ws, float* dw_orig, ssg_stream_t s) {
    SSG_CHECK_ARG(rows > 0 && cols > 0, "spectral_weight_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)s;
    long long n = (long long)rows * cols;
    SSG_CHECK_CUDA(cudaMemsetAsync(dot_ws, 0, sizeof(double), st));
    dot_kernel<<<grid_for(n, 256 * 8), 256, 0, st>>>(dw_sn, w_orig, n, dot_ws);
    spectral_bwd_kernel<<<grid_for(n, 256 * 4), 256, 0, st>>>(dw_sn, u, v, rows, cols, inv_sigma, dot_ws, dw_orig);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
