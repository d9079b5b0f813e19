// Device side of the data feed (dataset.py:95-144 + the Normalize / Flip steps of the albumentations
// pipelines at train_seg_gan.py:366-382): uint8 HWC rasters, exactly as cv2.imread delivers them, are copied to the GPU as
// BYTES (a quarter of the fp32 traffic the reference sends through .cuda()) and one kernel normalises, optionally flips,
// and lays them out either as the reference's NCHW fp32 tensor or directly as the channel-padded NHWC activation the
// first convolution reads.  Byte / elementwise work, HBM-bound: one thread per pixel, bytes read once, output written once.
#include "common.cuh"

namespace ssg {

// flip code per sample: bit 0 = reverse x (cv2.flip code 1), bit 1 = reverse y (cv2.flip code 0); both = cv2.flip code -1.
__device__ __forceinline__ long long src_pixel(long long img, int y, int x, int h, int w, const int* flip) {
    if (flip) {
        const int f = flip[img];
        if (f & 1) x = w - 1 - x;
        if (f & 2) y = h - 1 - y;
    }
    return (img * h + y) * (long long)w + x;
}

__device__ __forceinline__ float feed_value(unsigned char s, float sub, float mul, float post_div) {
    const float v = ((float)s - sub) * mul;
    return post_div != 0.f ? v / post_div : v;          // IEEE float32 division, as numpy's `img.astype('float32') / 255`
}

// albumentations.augmentations.functional.normalize: mean *= max_pixel; std *= max_pixel; denom = 1 / std (float32);
// img = (float32(img) - mean) * denom.   sub[c] = mean[c] * max_pixel, mul[c] = 1 / (std[c] * max_pixel) arrive precomputed
// in float32 by the host, so the kernel performs the same two float32 operations per element.
template <typename T, bool NCHW>
__global__ void __launch_bounds__(256) feed_image_kernel(const unsigned char* __restrict__ img, T* __restrict__ out, int n, int h, int w,
                                                          int c, int c_store, const float* __restrict__ sub, const float* __restrict__ mul,
                                                          float post_div, const int* __restrict__ flip) {
    const long long hw = (long long)h * w, total = (long long)n * hw, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const long long im = i / hw;
        const unsigned char* s = img + src_pixel(im, y, x, h, w, flip) * c;
        if (NCHW) {
            for (int k = 0; k < c; ++k)
                out[(im * c + k) * hw + (long long)y * w + x] = from_f<T>(feed_value(s[k], sub[k], mul[k], post_div));
        } else {
            T* d = out + i * c_store;
            for (int k = 0; k < c_store; ++k) d[k] = from_f<T>(k < c ? feed_value(s[k], sub[k], mul[k], post_div) : 0.f);
        }
    }
}
// the common case of the tensor-core path: 3 (or 4) bands stored as 8 bf16 channels -> one 16-byte store per pixel
__global__ void __launch_bounds__(256) feed_image_nhwc8_kernel(const unsigned char* __restrict__ img, bf16* __restrict__ out, int n, int h,
                                                                int w, int c, const float* __restrict__ sub, const float* __restrict__ mul,
                                                                float post_div, const int* __restrict__ flip) {
    const long long hw = (long long)h * w, total = (long long)n * hw, stride = (long long)gridDim.x * blockDim.x;
    float sb[8], ml[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sb[k] = k < c ? sub[k] : 0.f; ml[k] = k < c ? mul[k] : 0.f; }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const unsigned char* s = img + src_pixel(i / hw, y, x, h, w, flip) * c;
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = k < c ? feed_value(s[k], sb[k], ml[k], post_div) : 0.f;
        Vec<bf16> v; v.set(f); v.store(out + i * 8);
    }
}

// dataset.py:128-131: mask = (uint8)(float32(png) / 255.0) -> 1 only where the PNG holds 255; then float32, CHW.
__global__ void __launch_bounds__(256) feed_mask_kernel(const unsigned char* __restrict__ mask, float* __restrict__ out, int n, int h, int w,
                                                         int k, const int* __restrict__ flip) {
    const long long hw = (long long)h * w, total = (long long)n * hw, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const long long im = i / hw;
        const unsigned char* s = mask + src_pixel(im, y, x, h, w, flip) * k;
        for (int j = 0; j < k; ++j) out[(im * k + j) * hw + (long long)y * w + x] = (float)(unsigned char)((float)s[j] / 255.0f);
    }
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_feed_image_u8(const unsigned char* img, void* out, int dtype, int nchw, int n, int h, int w, int c, int c_store, const float* sub,
                      const float* mul, float post_div, const int* flip_codes, ssg_stream_t s) {
    SSG_CHECK_ARG(img && out && sub && mul && n > 0 && h > 0 && w > 0 && c > 0 && c <= 8 && c_store >= c, "feed_image_u8: bad arguments");
    SSG_CHECK_ARG(!nchw || c_store == c, "feed_image_u8: NCHW output is never channel-padded");
    const unsigned g = grid_for((long long)n * h * w, 256 * 2);
    if (!nchw && dtype == SSG_BF16 && c_store == 8) {
        feed_image_nhwc8_kernel<<<g, 256, 0, (cudaStream_t)s>>>(img, (bf16*)out, n, h, w, c, sub, mul, post_div, flip_codes);
    } else {
        SSG_DISPATCH_DTYPE(dtype, {
            if (nchw) feed_image_kernel<T, true><<<g, 256, 0, (cudaStream_t)s>>>(img, (T*)out, n, h, w, c, c_store, sub, mul, post_div, flip_codes);
            else feed_image_kernel<T, false><<<g, 256, 0, (cudaStream_t)s>>>(img, (T*)out, n, h, w, c, c_store, sub, mul, post_div, flip_codes);
        });
    }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_feed_mask_u8(const unsigned char* mask, float* out_nchw, int n, int h, int w, int classes, const int* flip_codes, ssg_stream_t s) {
    SSG_CHECK_ARG(mask && out_nchw && n > 0 && h > 0 && w > 0 && classes > 0, "feed_mask_u8: bad arguments");
    feed_mask_kernel<<<grid_for((long long)n * h * w, 256 * 2), 256, 0, (cudaStream_t)s>>>(mask, out_nchw, n, h, w, classes, flip_codes);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
