// stem7x7.cu — the 7x7 / stride-2 stem convolutions of the key and value encoders as ONE kernel each
// (reference: model/resnet.py:120 `conv1` of ResNet-50 in KeyEncoder, model/modules.py:124-136 ValueEncoder.forward:
// torch.cat([image, mask, others]) -> conv1 -> bn1), BatchNorm folded into weight + bias, optional ReLU.
//
// Round 1 materialised the im2col matrix ([pixels][192 or 256] fp16 = 40 MB per 480p frame, 60 MB of DRAM writes because
// its 48-byte rows straddle sectors) and ran a 1x1 GEMM over it: 22 + 11 us per frame.  Here the im2col tile never leaves
// the SM:
//   tile = 16 x 8 output pixels (UMMA M = 128) of one object; 2 persistent CTAs per SM (256 threads) walk over the tiles
//   1. weights [64][KPAD] fp16 -> shared memory in the UMMA K-major SW128 layout (before the programmatic-dependency wait:
//      they are constants)
//   2. the (2*16+5) x (2*8+5) input patch of every channel -> shared memory as fp16 (planar, zero padded); channel 3 is the
//      object's mask, channel 4 the sum of the other objects' masks (modules.py:126-134) -- built here, no torch.cat
//   3. two threads per output pixel gather the pixel's KPAD patch values (compile-time offsets) and store them as the
//      128-byte-swizzled A rows
//   4. one thread issues KPAD/16 tcgen05.mma (M128 x N64 x K16) into 64 TMEM columns
//   5. epilogue: TMEM -> registers -> +bias, ReLU -> fp16 -> staged in the (now free) A buffer -> coalesced NHWC stores
// K layout (must match XMem._ensure_packed, model/network.py put_stem):
//   C = 3: k = kh*24 + kw*3 + c   (kernel rows padded to 24, 8 rows -> 192)
//   C = 5: k = (kh*7 + kw)*5 + c  (245 -> 256)
#include <cuda_fp16.h>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

namespace {

constexpr int TW = 16, TH = 8;                       // output tile
constexpr int PW = 2 * TW + 5, PH = 2 * TH + 5;      // input patch 37 x 21
constexpr int PWP = 48;                              // padded patch row (halves): two rows = 48 words, so the two output rows
                                                     // a warp covers hit disjoint banks
template <int C> struct Cfg { static constexpr int KPAD = (C == 3) ? 192 : 256; static constexpr int KSTEPS = KPAD / 64; };

// patch offset (in halves, relative to the pixel's base) of K index k, or -1 for a zero slot
template <int C>
__host__ __device__ constexpr int k_offset(int k) {
    if (C == 3) {
        const int kh = k / 24, r = k % 24;
        if (kh >= 7 || r >= 21) return -1;
        return (r % 3) * (PH * PWP) + kh * PWP + r / 3;
    } else {
        if (k >= 49 * C) return -1;
        const int tap = k / C, c = k % C;
        return c * (PH * PWP) + (tap / 7) * PWP + tap % 7;
    }
}

template <int C>
struct Smem {
    alignas(1024) uint8_t a[Cfg<C>::KSTEPS][128 * 128];     // A: [k-step][128 pixel rows x 128 B], SW128; later the output stage
    alignas(1024) uint8_t b[Cfg<C>::KSTEPS][64 * 128];      // B: [k-step][64 cout rows x 128 B], SW128
    __half patch[C][PH][PWP];
    float bias[64];
    uint64_t done;
    uint32_t tmem_base;
};

template <int C, int HF>
__device__ __forceinline__ void build_rows(Smem<C>& sm, int r, const __half* pix) {
    // chunks j = HF, HF + 2, ... of this pixel's row: 8 consecutive k each -> one 16-byte store
#pragma unroll
    for (int jj = 0; jj < Cfg<C>::KPAD / 16; ++jj) {
        const int j = 2 * jj + HF;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o0 = k_offset<C>(8 * j + 2 * e), o1 = k_offset<C>(8 * j + 2 * e + 1);
            const uint32_t lo = (o0 >= 0) ? (uint32_t)__half_as_ushort(pix[o0 >= 0 ? o0 : 0]) : 0u;
            const uint32_t hi = (o1 >= 0) ? (uint32_t)__half_as_ushort(pix[o1 >= 0 ? o1 : 0]) : 0u;
            w[e] = lo | (hi << 16);
        }
        const int ks = j >> 3, jc = j & 7;
        *reinterpret_cast<uint4*>(&sm.a[ks][r * 128 + ((jc ^ (r & 7)) << 4)]) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int C>
__global__ void __launch_bounds__(256, 2)
stem7x7_kernel(const float* __restrict__ image, const float* __restrict__ masks, int n, int H, int W,
               const __half* __restrict__ weight, const float* __restrict__ bias, int relu, __half* __restrict__ out) {
    constexpr int KPAD = Cfg<C>::KPAD, KSTEPS = Cfg<C>::KSTEPS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    Smem<C>& sm = *reinterpret_cast<Smem<C>*>(base);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int Ho = H / 2, Wo = W / 2;
    const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
    const int n_tiles = tiles_x * tiles_y * n;

    // ---- constants: weights, bias, barrier, TMEM (all before the dependency wait; once per CTA, the CTA walks over tiles)
    {
        constexpr int NW = 64 * (KPAD / 8) / 256;
        uint4 wv[NW];
#pragma unroll
        for (int m = 0; m < NW; ++m) {
            const int i = t + 256 * m, nrow = i / (KPAD / 8), j = i % (KPAD / 8);
            wv[m] = __ldg(reinterpret_cast<const uint4*>(weight + (size_t)nrow * KPAD) + j);
        }
#pragma unroll
        for (int m = 0; m < NW; ++m) {
            const int i = t + 256 * m, nrow = i / (KPAD / 8), j = i % (KPAD / 8);
            *reinterpret_cast<uint4*>(&sm.b[j >> 3][nrow * 128 + (((j & 7) ^ (nrow & 7)) << 4)]) = wv[m];
        }
    }
    if (t < 64) sm.bias[t] = __ldg(bias + t);
    if (t == 0) { mbar_init(&sm.done, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&sm.tmem_base, 64); tmem_relinquish(); }
    pdl_wait();
    pdl_launch_dependents();

    const size_t plane = (size_t)H * W;
    uint32_t tmem = 0;
    int iter = 0;
    constexpr int NL = (C * PH * PW + 255) / 256;
    float pv[NL];                                             // this thread's share of a tile's input patch
    auto fetch = [&](int tile) {                              // global -> registers: every load in flight at once
        const int b = tile / (tiles_x * tiles_y), tl = tile % (tiles_x * tiles_y);
        const int gy0 = 2 * (tl / tiles_x) * TH - 3, gx0 = 2 * (tl % tiles_x) * TW - 3;
#pragma unroll
        for (int m = 0; m < NL; ++m) {
            const int i = t + 256 * m;
            const int x = i % PW, y = (i / PW) % PH, c = i / (PW * PH);
            const int gy = gy0 + y, gx = gx0 + x;
            pv[m] = 0.f;
            if (i < C * PH * PW && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const size_t off = (size_t)gy * W + gx;
                if (c < 3) pv[m] = __ldg(image + c * plane + off);
                else if (c == 3) pv[m] = __ldg(masks + b * plane + off);
                else {
                    float s = 0.f;
                    for (int j = 0; j < n; ++j) if (j != b) s += __ldg(masks + j * plane + off);
                    pv[m] = s;
                }
            }
        }
    };
    auto stash = [&]() {                                      // registers -> fp16 planar [c][y][x]
#pragma unroll
        for (int m = 0; m < NL; ++m) {
            const int i = t + 256 * m;
            if (i < C * PH * PW) sm.patch[i / (PW * PH)][(i / PW) % PH][i % PW] = __float2half_rn(pv[m]);
        }
    };
    if ((int)blockIdx.x < n_tiles) fetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        const int b = tile / (tiles_x * tiles_y), tl = tile % (tiles_x * tiles_y);
        const int tx0 = (tl % tiles_x) * TW, ty0 = (tl / tiles_x) * TH;
        stash();
        __syncthreads();

        // ---- im2col rows straight into the swizzled A operand
        {
            const int r = t & 127, px = r & 15, py = r >> 4;
            const __half* pix = &sm.patch[0][2 * py][2 * px];
            if (t < 128) build_rows<C, 0>(sm, r, pix); else build_rows<C, 1>(sm, r, pix);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem = sm.tmem_base;
        if (t == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, 64);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    mma_f16_ss(tmem, make_desc_sw128(smem_u32(sm.a[ks]) + j * 32), make_desc_sw128(smem_u32(sm.b[ks]) + j * 32), idesc,
                               (ks | j) ? 1u : 0u);
            mma_commit(&sm.done);
        }
        if (tile + (int)gridDim.x < n_tiles) fetch(tile + gridDim.x);     // next tile's loads fly during the MMAs and the epilogue
        mbar_wait(&sm.done, iter & 1, 60);
        tc_fence_after();

        // ---- epilogue: warp w reads TMEM lanes (w & 3) * 32.. (rows) and columns (w >> 2) * 32.. (output channels)
        {
            const int quad = warp & 3, ch = warp >> 2;
            const int r = quad * 32 + lane;
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + ch * 32, v);
            tmem_ld_wait();
            uint8_t* stage = sm.a[0];                         // the MMAs are done: A is free
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float f0 = __uint_as_float(v[q * 8 + 2 * e]) + sm.bias[ch * 32 + q * 8 + 2 * e];
                    float f1 = __uint_as_float(v[q * 8 + 2 * e + 1]) + sm.bias[ch * 32 + q * 8 + 2 * e + 1];
                    if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                    w[e] = pack_half2(f0, f1);
                }
                const int c16 = ch * 4 + q;
                *reinterpret_cast<uint4*>(&stage[r * 128 + ((c16 ^ (r & 7)) << 4)]) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        tc_fence_before();
        __syncthreads();
        {
            const uint8_t* stage = sm.a[0];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int i = t + 256 * m, r = i >> 3, c16 = i & 7;
                const int y = ty0 + (r >> 4), x = tx0 + (r & 15);
                if (y < Ho && x < Wo) {
                    const uint4 val = *reinterpret_cast<const uint4*>(&stage[r * 128 + ((c16 ^ (r & 7)) << 4)]);
                    *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + y) * Wo + x) * 64 + c16 * 8) = val;
                }
            }
        }
        // the next iteration's stash() is followed by a __syncthreads before anything overwrites the stage / A buffers;
        // the patch itself was last read before the barrier that preceded the MMAs
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(sm.tmem_base, 64); }
}

template <int C>
int launch_stem(const float* image, const float* masks, int n, int H, int W, const void* weight, const float* bias, int relu,
                void* out, cudaStream_t stream) {
    const size_t smem = sizeof(Smem<C>) + 1024;
    static XmPerDevice attr_token = {0};
    if (xm_first_use_on_device(&attr_token))
        XM_CHECK_CUDA(cudaFuncSetAttribute(stem7x7_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int Ho = H / 2, Wo = W / 2;
    const int n_tiles = ((Wo + TW - 1) / TW) * ((Ho + TH - 1) / TH) * n;
    const int resident = 2 * xm_num_sms();                      // two CTAs per SM (shared memory); each walks over its tiles
    const dim3 grid(n_tiles < resident ? n_tiles : resident);
    XM_CHECK_CUDA(tc5_launch(stem7x7_kernel<C>, grid, dim3(256), smem, stream, image, masks, n, H, W, (const __half*)weight, bias, relu,
                             (__half*)out));
    xm_count_launches(1);
    return XM_OK;
}

}  // namespace

// image fp32 [3][H][W]; masks fp32 [n][H][W] or NULL (key encoder: 3 channels, kpad 192; value encoder: 5 channels, kpad 256);
// weight fp16 [64][kpad] in the K layout above, bias fp32 [64]; out fp16 NHWC [n][H/2][W/2][64].
extern "C" int xm_stem7x7(const float* image, const float* masks, int32_t n, int32_t H, int32_t W, const void* weight, const float* bias,
                          int32_t kpad, int32_t relu, void* out, void* stream) {
    XM_REQUIRE(image && weight && bias && out && n >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "xm_stem7x7: bad arguments");
    XM_REQUIRE(masks ? kpad == 256 : (kpad == 192 && n == 1), "xm_stem7x7: kpad must be 192 (image only, n = 1) or 256 (image + masks)");
    if (masks) return launch_stem<5>(image, masks, n, H, W, weight, bias, relu, out, (cudaStream_t)stream);
    return launch_stem<3>(image, nullptr, n, H, W, weight, bias, relu, out, (cudaStream_t)stream);
}
