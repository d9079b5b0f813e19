// cv2.resize(uint8, INTER_LINEAR) arithmetic, shared by the device kernels (tiles.cu) and a host harness (tests/): OpenCV's
// resizeGeneric_ for 8U -- 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS), horizontal pass in int32, vertical pass
// ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2 (VResizeLinear<uchar, int, short, FixedPtCast<int, uchar, 22>>).
// Tables (one entry per output coordinate, built on the host: aerial_image_segmentation_api._linear_table) hold the two source
// indices, already clamped to the image, and the two coefficients.  Exact 2x shrinking is rerouted by cv::resize to the
// INTER_AREA fast path: (a + b + c + d + 2) >> 2.
#pragma once

#if defined(__CUDACC__)
#define SSG_HD __host__ __device__ __forceinline__
#else
#define SSG_HD inline
#endif

struct ssg_lin_tap {      // one output coordinate
    int i0, i1;           // source indices of the two taps (clamped)
    int c0, c1;           // 11-bit coefficients, c0 + c1 == 2048
};

SSG_HD int ssg_cv_lin_u8(int s00, int s01, int s10, int s11, int a0, int a1, int b0, int b1) {
    const int r0 = s00 * a0 + s01 * a1;       // <= 255 * 2048
    const int r1 = s10 * a0 + s11 * a1;
    return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

SSG_HD int ssg_cv_area2_u8(int s00, int s01, int s10, int s11) { return (s00 + s01 + s10 + s11 + 2) >> 2; }

// One output element of the uint8 resize of an NHWC raster: (oy, ox, ch) of image `img` (h x w x c).
SSG_HD unsigned char ssg_cv_resize_px(const unsigned char* img, int w, int c, int ch, int oy, int ox, const ssg_lin_tap* xt,
                                      const ssg_lin_tap* yt, int area2x) {
    if (area2x) {
        const unsigned char* p = img + ((long long)(2 * oy) * w + 2 * ox) * c + ch;
        return (unsigned char)ssg_cv_area2_u8(p[0], p[c], p[(long long)w * c], p[(long long)w * c + c]);
    }
    const ssg_lin_tap tx = xt[ox], ty = yt[oy];
    const unsigned char* r0 = img + (long long)ty.i0 * w * c + ch;
    const unsigned char* r1 = img + (long long)ty.i1 * w * c + ch;
    const int v = ssg_cv_lin_u8(r0[(long long)tx.i0 * c], r0[(long long)tx.i1 * c], r1[(long long)tx.i0 * c], r1[(long long)tx.i1 * c],
                                tx.c0, tx.c1, ty.c0, ty.c1);
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
