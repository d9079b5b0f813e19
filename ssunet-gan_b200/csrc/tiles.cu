// Tiled-inference merge (aerial_image_segmentation_api.py:129-217, the loop body of patch_merge): every patch votes
// 0 / 1 per class and pixel -- (uint8)(p * 255) > 127, which is what post_process_resized_mask leaves of the patch when
// no resize is involved -- into integer counters of the full raster; the final mask is (uint8)(votes / patches * 255)
// pushed through the same 127 threshold.  The reference accumulates {0.0, 1.0} in float64, so integer counting is exact,
// and the final division / scaling is done in fp64 exactly as numpy does it: the uint8 masks are bit-identical.
#include "common.cuh"
#include "cvresize.h"

namespace ssg {

// cv2.resize(uint8 NHWC, INTER_LINEAR): one thread per output element (get_patched_input, aerial_image_segmentation_api.py:361).
__global__ void __launch_bounds__(256) resize_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int n, int h,
                                                         int w, int c, int oh, int ow, const ssg_lin_tap* __restrict__ xt,
                                                         const ssg_lin_tap* __restrict__ yt, int area2x) {
    const long long total = (long long)n * oh * ow * c, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ch = (int)(i % c);
        long long p = i / c;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const long long im = p / oh;
        dst[i] = ssg_cv_resize_px(src + im * h * w * c, w, c, ch, oy, ox, xt, yt, area2x);
    }
}

// mask_vote_kernel with the reference's resize bridge (:151-152): the S x S map is quantised to uint8, upsampled to the
// patch size with cv2.resize's arithmetic, thresholded at 127 and voted -- per OUTPUT pixel, nothing is materialised.
__global__ void __launch_bounds__(256) mask_vote_resized_kernel(const float* __restrict__ v, const int* __restrict__ win, int P, int C, int S,
                                                                 int PS, int H, int W, int apply_sigmoid, const ssg_lin_tap* __restrict__ xt,
                                                                 const ssg_lin_tap* __restrict__ yt, int area2x, int* __restrict__ pos,
                                                                 int* __restrict__ cnt) {
    const long long total = (long long)P * PS * PS, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % PS), y = (int)((i / PS) % PS), p = (int)(i / ((long long)PS * PS));
        const int gy = win[2 * p] + y, gx = win[2 * p + 1] + x;
        if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
        atomicAdd(cnt + (long long)gy * W + gx, 1);
        int y0, y1, x0, x1, a0 = 0, a1 = 0, b0 = 0, b1 = 0;
        if (area2x) { y0 = 2 * y; y1 = 2 * y + 1; x0 = 2 * x; x1 = 2 * x + 1; }
        else {
            const ssg_lin_tap tx = xt[x], ty = yt[y];
            y0 = ty.i0; y1 = ty.i1; x0 = tx.i0; x1 = tx.i1; a0 = tx.c0; a1 = tx.c1; b0 = ty.c0; b1 = ty.c1;
        }
        for (int c = 0; c < C; ++c) {
            const float* m = v + ((long long)p * C + c) * S * S;
            int q[4];
            const int yy[4] = {y0, y0, y1, y1}, xx[4] = {x0, x1, x0, x1};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float pr = m[(long long)yy[k] * S + xx[k]];
                if (apply_sigmoid) pr = 1.0f / (1.0f + expf(-pr));                 // torch.sigmoid in fp32 (:384)
                const float scaled = pr * 255.0f;                                    // (mask * 255).astype('uint8') (:150)
                q[k] = scaled >= 255.f ? 255 : (scaled <= 0.f ? 0 : (int)scaled);
            }
            const int u8 = area2x ? ssg_cv_area2_u8(q[0], q[1], q[2], q[3]) : ssg_cv_lin_u8(q[0], q[1], q[2], q[3], a0, a1, b0, b1);
            if (u8 > 127) atomicAdd(pos + ((long long)c * H + gy) * W + gx, 1);      // post_process: > 127 -> 255 -> / 255.0 = 1.0
        }
    }
}

// probs_or_logits: [P][C][S][S] fp32 (NCHW, as the model returns them); win: [P][2] = (h1, w1) of every patch.
__global__ void __launch_bounds__(256) mask_vote_kernel(const float* __restrict__ v, const int* __restrict__ win, int P, int C, int S,
                                                         int H, int W, int apply_sigmoid, int* __restrict__ pos, int* __restrict__ cnt) {
    const long long total = (long long)P * S * S, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % S), y = (int)((i / S) % S), p = (int)(i / ((long long)S * S));
        const int gy = win[2 * p] + y, gx = win[2 * p + 1] + x;
        if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
        atomicAdd(cnt + (long long)gy * W + gx, 1);
        for (int c = 0; c < C; ++c) {
            float pr = v[(((long long)p * C + c) * S + y) * S + x];
            if (apply_sigmoid) pr = 1.0f / (1.0f + expf(-pr));                 // torch.sigmoid in fp32 (:384)
            const float scaled = pr * 255.0f;                                    // (mask * 255).astype('uint8') (:151)
            const int u8 = scaled >= 255.f ? 255 : (scaled <= 0.f ? 0 : (int)scaled);
            if (u8 > 127) atomicAdd(pos + ((long long)c * H + gy) * W + gx, 1);  // post_process: > 127 -> 255 -> / 255.0 = 1.0
        }
    }
}

__global__ void __launch_bounds__(256) mask_finalize_kernel(const int* __restrict__ pos, const int* __restrict__ cnt, int C, long long HW,
                                                             unsigned char* __restrict__ out) {
    const long long total = (long long)C * HW, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long px = i % HW;
        const int n = cnt[px];
        const double full = (double)pos[i] / (double)(n == 0 ? 1 : n);          // np.divide(merged_mask, mask_merge_div) (:211-214)
        const int u8 = (int)(full * 255.0);                                      // (full_mask * 255).astype('uint8') (:215)
        out[i] = u8 > 127 ? 255 : 0;                                             // post_process_resized_mask (:30-42): {0, 255}
    }
}

}  // namespace ssg
using namespace ssg;

extern "C" {

int ssg_mask_vote(const float* values, const int* windows, int patches, int classes, int patch_size, int h, int w, int apply_sigmoid,
                  int* pos_votes, int* patch_count, ssg_stream_t s) {
    SSG_CHECK_ARG(values && windows && pos_votes && patch_count && patches > 0 && classes > 0 && patch_size > 0 && h > 0 && w > 0,
                  "mask_vote: bad arguments");
    mask_vote_kernel<<<grid_for((long long)patches * patch_size * patch_size, 256 * 4), 256, 0, (cudaStream_t)s>>>(
        values, windows, patches, classes, patch_size, h, w, apply_sigmoid, pos_votes, patch_count);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_resize_u8_linear(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* xtab,
                         const int* ytab, ssg_stream_t s) {
    SSG_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "resize_u8_linear: bad arguments");
    const int area2x = (h == 2 * oh && w == 2 * ow) ? 1 : 0;
    SSG_CHECK_ARG(area2x || (xtab && ytab), "resize_u8_linear: coefficient tables missing");
    resize_u8_kernel<<<grid_for((long long)n * oh * ow * c, 256 * 4), 256, 0, (cudaStream_t)s>>>(
        src, dst, n, h, w, c, oh, ow, (const ssg_lin_tap*)xtab, (const ssg_lin_tap*)ytab, area2x);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_mask_vote_resized(const float* values, const int* windows, int patches, int classes, int map_size, int patch_size, int h, int w,
                          int apply_sigmoid, const int* xtab, const int* ytab, int* pos_votes, int* patch_count, ssg_stream_t s) {
    SSG_CHECK_ARG(values && windows && pos_votes && patch_count && patches > 0 && classes > 0 && map_size > 0 && patch_size > 0 && h > 0 && w > 0,
                  "mask_vote_resized: bad arguments");
    const int area2x = (map_size == 2 * patch_size) ? 1 : 0;
    SSG_CHECK_ARG(area2x || (xtab && ytab), "mask_vote_resized: coefficient tables missing");
    mask_vote_resized_kernel<<<grid_for((long long)patches * patch_size * patch_size, 256 * 4), 256, 0, (cudaStream_t)s>>>(
        values, windows, patches, classes, map_size, patch_size, h, w, apply_sigmoid, (const ssg_lin_tap*)xtab, (const ssg_lin_tap*)ytab,
        area2x, pos_votes, patch_count);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int ssg_mask_finalize(const int* pos_votes, const int* patch_count, int classes, int h, int w, unsigned char* masks, ssg_stream_t s) {
    SSG_CHECK_ARG(pos_votes && patch_count && masks && classes > 0 && h > 0 && w > 0, "mask_finalize: bad arguments");
    mask_finalize_kernel<<<grid_for((long long)classes * h * w, 256 * 4), 256, 0, (cudaStream_t)s>>>(pos_votes, patch_count, classes,
                                                                                                   (long long)h * w, masks);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // extern "C"
