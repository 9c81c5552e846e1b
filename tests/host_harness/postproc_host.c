/* Host harness for tests/test_postproc.py: runs the per-pixel function of xmem2_b200/csrc/postproc_math.h (the one the CUDA
 * kernel calls) over a whole image on the CPU.  Test infrastructure only; never linked into the product library. */
#include "../../xmem2_b200/csrc/postproc_math.h"

void resize_argmax_host(const float* prob, int channels, int in_h, int in_w, long long stride_c, long long stride_h, int out_h, int out_w,
                        const unsigned char* lut, unsigned char* out) {
    for (int oy = 0; oy < out_h; ++oy)
        for (int ox = 0; ox < out_w; ++ox)
            out[(long long)oy * out_w + ox] = xm_resize_argmax_pixel(prob, channels, in_h, in_w, stride_c, stride_h, out_h, out_w, oy, ox, lut);
}
