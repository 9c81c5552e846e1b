// postproc.cu — fused post-processing of a frame's probabilities (driver side):
// bilinear resize to the original frame size + argmax over the objects + optional label remap, one byte per pixel out.
// Replaces `_post_process` of the reference driver (inference/run_on_video.py:165-173: F.interpolate -> argmax ->
// uint8), which materialises the resized fp32 probabilities of every object first.  Memory-bound elementwise work on
// CUDA cores: one thread per output pixel, 4 taps per channel (L1/L2 resident), 1 byte written.
// The per-pixel math lives in postproc_math.h and is verified on the CPU against torch (tests/test_postproc.py).
#include "common.h"
#include "tc5.cuh"
#include "postproc_math.h"

namespace {

__global__ void resize_argmax_kernel(const float* __restrict__ prob, int channels, int in_h, int in_w, int64_t stride_c, int64_t stride_h,
                                     int out_h, int out_w, const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
    tc5::pdl_wait();
    tc5::pdl_launch_dependents();
    const int64_t total = (int64_t)out_h * out_w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(i / out_w), ox = (int)(i - (int64_t)oy * out_w);
        out[i] = xm_resize_argmax_pixel(prob, channels, in_h, in_w, stride_c, stride_h, out_h, out_w, oy, ox, lut);
    }
}

}  // namespace

extern "C" int xm_resize_argmax(const float* prob, int32_t channels, int32_t in_h, int32_t in_w, int64_t stride_c, int64_t stride_h,
                                int32_t out_h, int32_t out_w, const uint8_t* lut, uint8_t* out, void* stream) {
    XM_REQUIRE(prob && out, "xm_resize_argmax: null pointer");
    XM_REQUIRE(channels >= 1 && channels <= 256, "xm_resize_argmax: 1..256 channels (labels are bytes)");
    XM_REQUIRE(in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1 && stride_h >= in_w && stride_c >= (int64_t)(in_h - 1) * stride_h + in_w,
               "xm_resize_argmax: bad shape/strides");
    const int64_t total = (int64_t)out_h * out_w;
    int blocks = (int)((total + 255) / 256);
    const int cap = xm_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    XM_CHECK_CUDA(tc5_launch(resize_argmax_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, prob, (int)channels, (int)in_h, (int)in_w,
                             stride_c, stride_h, (int)out_h, (int)out_w, lut, out));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}
