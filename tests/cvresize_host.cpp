// Host harness for tests/test_cpu_host.py: runs the SAME per-pixel function the CUDA kernels call (csrc/cvresize.h) over a
// raster on the CPU, so the fixed-point arithmetic and the table format are checked against cv2 without a GPU.
#include "../ssunet-gan_b200/csrc/cvresize.h"

extern "C" void host_resize_u8(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, int oh, int ow,
                               const int* xtab, const int* ytab) {
    const int area2x = (h == 2 * oh && w == 2 * ow) ? 1 : 0;
    for (int im = 0; im < n; ++im)
        for (int oy = 0; oy < oh; ++oy)
            for (int ox = 0; ox < ow; ++ox)
                for (int ch = 0; ch < c; ++ch)
                    dst[(((long long)im * oh + oy) * ow + ox) * c + ch] =
                        ssg_cv_resize_px(src + (long long)im * h * w * c, w, c, ch, oy, ox, (const ssg_lin_tap*)xtab,
                                         (const ssg_lin_tap*)ytab, area2x);
}
