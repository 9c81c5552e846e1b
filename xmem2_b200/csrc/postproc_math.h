// postproc_math.h — per-pixel math of the fused "bilinear resize + argmax (+ label remap)" post-processing kernel
// (reference inference/run_on_video.py:165-173 `_post_process`: F.interpolate(prob, shape, mode='bilinear',
// align_corners=False) -> torch.argmax(dim=0) -> uint8; the saver then remaps labels, util/image_saver.py + MaskMapper).
// Plain C so that the SAME function is compiled into the CUDA kernel (postproc.cu) and into a host test harness
// (tests/host_harness/postproc_host.c, built with gcc by tests/test_postproc.py) — the arithmetic is checked on the CPU
// against torch even when no GPU is around.
#ifndef XM_POSTPROC_MATH_H
#define XM_POSTPROC_MATH_H
#include <stdint.h>

#ifdef __CUDACC__
#define XM_HD __host__ __device__ __forceinline__
#else
#define XM_HD static inline
#endif

// PyTorch's source index for align_corners=False (ATen UpSample.h, area_pixel_compute_source_index): the centre of
// destination pixel `dst` mapped back, clamped at 0; i0 = floor, i1 = i0 + (i0 < in-1), lambda = fractional part.
XM_HD void xm_bilinear_tap(int dst, int in_size, int out_size, int* i0, int* i1, float* lambda1) {
    // exact IEEE division: the library is compiled with --use_fast_math, whose approximate division (2 ulp) moved the
    // source coordinate of 480 -> 1080 resizes by ~1e-4 pixels (round-1 failure of tests/test_gpu_zz_postproc.py on the
    // B200).  ATen's CUDA kernel evaluates `scale * (dst + 0.5) - 0.5` with nvcc's default FMA contraction.
#ifdef __CUDA_ARCH__
    const float scale = __fdiv_rn((float)in_size, (float)out_size);
    float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
#else
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
#endif
    if (src < 0.f) src = 0.f;
    int i = (int)src;
    if (i > in_size - 1) i = in_size - 1;
    *i0 = i;
    *i1 = i + ((i < in_size - 1) ? 1 : 0);
    *lambda1 = src - (float)i;
}

// label of output pixel (oy, ox): argmax over channels of the bilinearly resized probabilities (first maximum wins, as
// torch.argmax), optionally mapped through a 256-entry table (MaskMapper.remap_index_mask).
XM_HD uint8_t xm_resize_argmax_pixel(const float* prob, int channels, int in_h, int in_w, int64_t stride_c, int64_t stride_h,
                                     int out_h, int out_w, int oy, int ox, const uint8_t* lut) {
    int y0, y1, x0, x1;
    float ly, lx;
    xm_bilinear_tap(oy, in_h, out_h, &y0, &y1, &ly);
    xm_bilinear_tap(ox, in_w, out_w, &x0, &x1, &lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    int best = 0;
    float best_v = 0.f;
    for (int c = 0; c < channels; ++c) {
        const float* p = prob + (int64_t)c * stride_c;
        const float* r0 = p + (int64_t)y0 * stride_h;
        const float* r1 = p + (int64_t)y1 * stride_h;
        // same association as ATen's upsample_bilinear2d kernel
        const float v = hy * (hx * r0[x0] + lx * r0[x1]) + ly * (hx * r1[x0] + lx * r1[x1]);
        if (c == 0 || v > best_v) { best_v = v; best = c; }
    }
    return lut ? lut[best & 255] : (uint8_t)best;
}
#endif
