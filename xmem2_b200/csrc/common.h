// common.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/xmem2_b200.h"

void xm_set_error(const char* fmt, ...);

#define XM_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            xm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return XM_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define XM_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            xm_set_error(__VA_ARGS__);   \
            return XM_ERR_ARG;           \
        }                                \
    } while (0)

// Encode a tiled fp16 tensor map (128-byte swizzle).  dims/box are innermost-first; strides_bytes has
// rank-1 entries (stride of dims 1..rank-1).  Returns 0 on success.
int xm_make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

int xm_num_sms();      // SM count of the CURRENT device (cached per device ordinal)

// Per-device one-time initialisation (cudaFuncSetAttribute, __device__ symbol uploads): function attributes and symbols
// belong to a device's context, so a process that drives several GPUs has to repeat them on each one.
struct XmPerDevice { unsigned long long done_mask; };       // zero-initialised static at the call site; <= 64 devices
bool xm_first_use_on_device(XmPerDevice* token);             // true exactly once per (token, current device); thread-safe

void xm_count_launches(int n);

// CTA-pair (cta_group::2) convolution, conv_igemm_pair.cu; same contract as xm_conv2d_nhwc, cout_pad % 128 == 0 only
int xm_conv2d_pair(const xm_conv_args_t* a, void* stream);
