// EXPERIMENTAL (round-2 head start, NOT part of libxmem2_b200.so, never run on a GPU yet):
// conv_igemm_halo.cu — 3x3 stride-1 convolutions WITHOUT re-loading the input tile for every filter tap.
// The production kernel (../conv_igemm.cu) issues one 16 KB TMA box of input pixels per (tap, 64-channel block): the nine
// taps of a 3x3 filter read nine shifted copies of (almost) the same pixels, and the kernel is bound by exactly those
// L2->SM bytes (ROUND1_NOTES.md / tests/conv_model.py).  Here a CTA loads, per 64-channel block, ONE haloed tile
//     rows y0-1 .. y0+16 (18), columns x0-1 .. x0+14 (16, of which 10 are used), 64 channels  = 36 KB
// and the nine taps are nine VIEWS of it: the output tile is 8 wide x 16 tall, so the 128 GEMM rows are 16 groups of
// 8 consecutive pixels, group g = output row y0+g; for tap (kh, kw) group g starts at halo row (g + kh), column kw, i.e. at
//     byte offset ((g + kh) * 16 + kw) * 128  ->  start = (kh*16 + kw) * 128,  stride between groups SBO = 16 * 128 = 2048 B,
// a legal K-major SWIZZLE_128B operand whose start is shifted by kw rows inside the 1024-byte swizzle atom: the UMMA
// descriptor's base_offset field (bits 49-51) = kw tells the tensor core the phase TMA used when it wrote the rows.
// L2->SM bytes per 64-channel block: 36 KB + 9 * BN * 128 B   instead of   9 * (16 KB + BN * 128 B):
//     BN = 64 : 108 KB vs 216 KB (2.0x less)      BN = 128 : 180 KB vs 288 KB (1.6x less)
// Fallback layout (XMEM_HALO_3BOX=1, template THREE_BOX): three 8-pixel-wide boxes per channel block, one per kw (x0-1+kw ..), 18 KB each
// = 54 KB; every tap view then starts on a 1024-byte boundary (kh * 1024) with the canonical SBO of 1024 B and base_offset 0 — only
// descriptor forms the production kernels already use — at 1.5x the input bytes of the single haloed tile (still 2.7x fewer than today).
// Pipeline: the halo tile is double buffered per channel block (afull/aempty), the nine weight tiles of a block stream
// through their own ring (bfull/bempty).  Everything else (concatenated sources, epilogue, TMA store) is the production
// code with the tile rectangle fixed to 8 x 16.  Only ksize 3 / stride 1; no split-K.
// Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -c conv_igemm_halo.cu
//
// (header of the production file follows)
// conv_igemm.cu — implicit-GEMM convolution on tcgen05 for NHWC fp16 activations.
//
// One kernel serves every 1x1 / 3x3 (stride 1 or 2, padding k/2) convolution on the XMem++ path
// (reference: nn.Conv2d call sites in model/resnet.py:46-114, model/modules.py:22-41,178-211,229-250,
// model/group_modules.py:29-54), with BatchNorm folded into weight+bias on the host and bias /
// residual-add / ReLU fused into the epilogue.
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[cout, tap, cin]
//   M tile = 128 output pixels arranged as a TW x TH rectangle (TW*TH = 128)
//   N tile = BN output channels (64 or 128), K step = 64 input channels of one filter tap.
//   A operand: ONE TMA box [64 ch, TW, TH, 1] of the NHWC input, shifted by the tap offset; TMA's
//     out-of-bounds zero fill implements the convolution padding, and the box lands in shared memory
//     as 128 rows (pixels) x 128 B with the 128-byte swizzle == the UMMA K-major SW128 layout.
//     Stride-2 convs view the input as parity planes [C, 2, W/2, 2, H/2] (rank-5 map), which turns the
//     strided gather into a plain box again.
//   B operand: TMA box [64, BN] of the [cout_pad][taps*cin] weight matrix.
//   Channel-concatenated inputs (torch.cat along C in the reference) are read from up to three source
//   tensors without materialising the concat; a source may be broadcast over the batch.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
#include <cstdio>
#include <cstdlib>
#include "../common.h"
#include "../tc5.cuh"

using namespace tc5;

namespace {

constexpr int HALO_W = 16;                       // pixel pitch of the haloed tile (x0-1 .. x0+14)
constexpr int HALO_H = 18;                       // rows y0-1 .. y0+16
constexpr int HALO_BYTES = HALO_H * HALO_W * 128;   // 36 KB per 64-channel block (a multiple of 1024)
constexpr int BOX3_BYTES = HALO_H * 8 * 128;        // 18 KB: one of the three 8-pixel-wide boxes of the fallback layout
constexpr int A_BYTES = 3 * BOX3_BYTES;             // stage size that fits both layouts (54 KB)
constexpr int A_STAGES = 2;

// K-major SWIZZLE_128B operand whose 8-row groups are `sbo_bytes` apart and whose first row sits `row_phase` rows into
// a 1024-byte swizzle atom (base_offset, bits 49-51)
__device__ __forceinline__ uint64_t make_desc_sw128_view(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t row_phase) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(row_phase & 7u) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

struct alignas(64) ConvMaps {
    CUtensorMap a[3];
    CUtensorMap w;
    CUtensorMap o;      // output   [out_stride, Wo, Ho, B]   box [64, TW, TH, 1]   (TMA-store epilogue)
    CUtensorMap r;      // residual [cout, Wo, Ho, B|1]       box [64, TW, TH, 1]
};

struct ConvP {
    int n_src;
    int cblocks[3];      // channels / 64 per source
    int choff[3];        // channel offset of the source inside the concatenated input
    int bcast[3];
    int cin_total;
    int ksize, stride, pad;
    int tw, th, tiles_x, tiles_y;
    int Ho, Wo, batch;
    int cout;
    int relu;
    const float* bias;
    const __half* residual;
    int residual_bcast, residual_stride;
    __half* out;
    __half* out_relu;
    int out_stride, out_offset;
    int tma_epilogue;                 // 1: stage the tile in swizzled smem, residual in / output out through TMA
};

template <int BN, int CONV_STAGES>
struct ConvSmem {
    alignas(1024) uint8_t a[A_STAGES][A_BYTES];         // haloed input tile (36 KB) or three kw boxes (54 KB) of one 64-channel block
    alignas(1024) uint8_t b[CONV_STAGES][BN * 128];     // weight tiles (one per tap), ring
    alignas(8) uint64_t afull[A_STAGES];
    uint64_t aempty[A_STAGES];
    uint64_t full[CONV_STAGES];
    uint64_t empty[CONV_STAGES];
    uint64_t done;
    uint64_t resbar;
    uint32_t tmem_base;
    float bias[BN];
};

template <int BN, int CONV_STAGES, bool THREE_BOX>
__global__ void __launch_bounds__(192)
conv_igemm_halo_kernel(const __grid_constant__ ConvMaps maps, const ConvP p) {
    extern __shared__ uint8_t smem_raw[];
    using Smem = ConvSmem<BN, CONV_STAGES>;
    Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int tile = blockIdx.x;
    const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
    const int b = tile;
    const int x0 = tx_i * p.tw, y0 = ty_i * p.th;
    const int n0 = blockIdx.y * BN;

    int cb_total = 0;
    for (int s = 0; s < p.n_src; ++s) cb_total += p.cblocks[s];

    if (threadIdx.x == 0) {
        for (int i = 0; i < CONV_STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < A_STAGES; ++i) { mbar_init(&sm.afull[i], 1); mbar_init(&sm.aempty[i], 1); }
        mbar_init(&sm.done, 1);
        mbar_init(&sm.resbar, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, BN); tmem_relinquish(); }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.w);
        for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&maps.a[s]);
    }
    if (warp >= 2) {      // stage this CTA's bias slice (weights: independent of the preceding kernel)
        const int t0 = threadIdx.x - 64;
        for (int t = t0; t < BN; t += 128) sm.bias[t] = (n0 + t < p.cout) ? __ldg(p.bias + n0 + t) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_wait();                    // inputs (and the split-K workspace) come from preceding kernels
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&maps.a[s]);
            int it = 0;
            for (int cbg = 0; cbg < cb_total; ++cbg) {
                int cb = cbg, s = 0;
                if (p.n_src > 1 && cb >= p.cblocks[0]) { cb -= p.cblocks[0]; s = 1; }
                if (p.n_src > 2 && s == 1 && cb >= p.cblocks[1]) { cb -= p.cblocks[1]; s = 2; }
                const int bb = p.bcast[s] ? 0 : b;
                const int ast = cbg % A_STAGES, aph = (cbg / A_STAGES) & 1;
                mbar_wait(&sm.aempty[ast], aph ^ 1, 25);
                // rows y0-1.., columns x0-1..: out-of-range pixels are zero-filled = the convolution padding
                if (THREE_BOX) {
                    mbar_expect_tx(&sm.afull[ast], 3 * BOX3_BYTES);
                    for (int kw = 0; kw < 3; ++kw)
                        tma_load_4d(sm.a[ast] + kw * BOX3_BYTES, &maps.a[s], &sm.afull[ast], cb * 64, x0 - 1 + kw, y0 - 1, bb);
                } else {
                    mbar_expect_tx(&sm.afull[ast], HALO_BYTES);
                    tma_load_4d(sm.a[ast], &maps.a[s], &sm.afull[ast], cb * 64, x0 - 1, y0 - 1, bb);
                }
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int st = it % CONV_STAGES, ph = (it / CONV_STAGES) & 1;
                    mbar_wait(&sm.empty[st], ph ^ 1, 21);
                    mbar_expect_tx(&sm.full[st], BN * 128);
                    tma_load_2d(sm.b[st], &maps.w, &sm.full[st], tap * p.cin_total + p.choff[s] + cb * 64, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, BN);
            int it = 0;
            for (int cbg = 0; cbg < cb_total; ++cbg) {
                const int ast = cbg % A_STAGES, aph = (cbg / A_STAGES) & 1;
                mbar_wait(&sm.afull[ast], aph, 26);
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    const int st = it % CONV_STAGES, ph = (it / CONV_STAGES) & 1;
                    const int kh = tap / 3, kw = tap - 3 * kh;
                    mbar_wait(&sm.full[st], ph, 22);
                    tc_fence_after();
                    const uint32_t a0 = THREE_BOX ? smem_u32(sm.a[ast]) + (uint32_t)(kw * BOX3_BYTES + kh * 1024)
                                                  : smem_u32(sm.a[ast]) + (uint32_t)((kh * HALO_W + kw) * 128);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint64_t ad = THREE_BOX ? make_desc_sw128(a0 + j * 32) : make_desc_sw128_view(a0 + j * 32, HALO_W * 128, (uint32_t)kw);
                        uint64_t bd = make_desc_sw128(smem_u32(sm.b[st]) + j * 32);
                        mma_f16_ss(tmem, ad, bd, idesc, (it | j) ? 1u : 0u);
                    }
                    mma_commit(&sm.empty[st]);
                }
                mma_commit(&sm.aempty[ast]);
            }
            mma_commit(&sm.done);
        }
    } else {
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int yo = y0 + row / p.tw, xo = x0 + row % p.tw;
        const bool pix_ok = (yo < p.Ho) && (xo < p.Wo);
        const size_t pix = ((size_t)b * p.Ho + yo) * p.Wo + xo;
        const size_t rpix = ((size_t)(p.residual_bcast ? 0 : b) * p.Ho + yo) * p.Wo + xo;
        mbar_wait(&sm.done, 0, 23);
        tc_fence_after();
        if (p.tma_epilogue) {
            // Stage buffers are free now (every MMA has completed): a[] holds the output tile, b[] the residual tile,
            // both as 64-channel boxes of 128 pixel rows x 128 B with the 128-byte swizzle (conflict-free 16-B accesses).
            uint8_t* stage_out = &sm.a[0][0];
            uint8_t* stage_res = &sm.b[0][0];
            const int nbox = min(BN / 64, (p.cout - n0) / 64);        // cout is a multiple of 64 on this path
            if (p.residual && threadIdx.x == 64) {
                mbar_expect_tx(&sm.resbar, nbox * 128 * 128);
                for (int k = 0; k < nbox; ++k)
                    tma_load_4d(stage_res + k * 128 * 128, &maps.r, &sm.resbar, n0 + 64 * k, x0, y0, p.residual_bcast ? 0 : b);
            }
#pragma unroll 1
            for (int k = 0; k < nbox; ++k) {
                float v[64];
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int c0 = 64 * k + 32 * hlf;
                    {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[32 * hlf + j] = __uint_as_float(r[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[32 * hlf + j] += sm.bias[c0 + j];
                }
                if (p.residual) {
                    if (k == 0) mbar_wait(&sm.resbar, 0, 24);
                    const uint8_t* rrow = stage_res + k * 128 * 128 + row * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 u = *reinterpret_cast<const uint4*>(rrow + ((c ^ (row & 7)) << 4));
                        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h2[e]);
                            v[8 * c + 2 * e] += f.x; v[8 * c + 2 * e + 1] += f.y;
                        }
                    }
                }
                uint8_t* orow = stage_out + k * 128 * 128 + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 u;
                    if (p.relu) {
                        u.x = pack_half2(fmaxf(v[8 * c], 0.f), fmaxf(v[8 * c + 1], 0.f)); u.y = pack_half2(fmaxf(v[8 * c + 2], 0.f), fmaxf(v[8 * c + 3], 0.f));
                        u.z = pack_half2(fmaxf(v[8 * c + 4], 0.f), fmaxf(v[8 * c + 5], 0.f)); u.w = pack_half2(fmaxf(v[8 * c + 6], 0.f), fmaxf(v[8 * c + 7], 0.f));
                    } else {
                        u.x = pack_half2(v[8 * c], v[8 * c + 1]); u.y = pack_half2(v[8 * c + 2], v[8 * c + 3]);
                        u.z = pack_half2(v[8 * c + 4], v[8 * c + 5]); u.w = pack_half2(v[8 * c + 6], v[8 * c + 7]);
                    }
                    *reinterpret_cast<uint4*>(orow + ((c ^ (row & 7)) << 4)) = u;
                }
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                for (int k = 0; k < nbox; ++k) tma_store_4d(&maps.o, stage_out + k * 128 * 128, p.out_offset + n0 + 64 * k, x0, y0, b);
                tma_store_commit();
                tma_store_wait_read();
            }
            goto teardown;
        }
        const bool vec_ok = (p.out_stride % 8 == 0) && (p.out_offset % 8 == 0) && (p.residual_stride % 8 == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float acc[32];
            {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
            }
            const int n = n0 + c0;
            if (!pix_ok || n >= p.cout) continue;
            const bool full = (n + 32 <= p.cout) && vec_ok;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = acc[j] + sm.bias[c0 + j];
            if (p.residual) {
                const __half* rp = p.residual + rpix * p.residual_stride + n;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 u = *reinterpret_cast<const uint4*>(rp + j);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 f = __half22float2(h2[e]);
                            v[j + 2 * e] += f.x; v[j + 2 * e + 1] += f.y;
                        }
                    }
                } else {
                    for (int j = 0; j < 32 && n + j < p.cout; ++j) v[j] += __half2float(rp[j]);
                }
            }
            if (p.out) {
                __half* op = p.out + pix * p.out_stride + p.out_offset + n;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 u;
                        u.x = pack_half2(p.relu ? fmaxf(v[j], 0.f) : v[j], p.relu ? fmaxf(v[j + 1], 0.f) : v[j + 1]);
                        u.y = pack_half2(p.relu ? fmaxf(v[j + 2], 0.f) : v[j + 2], p.relu ? fmaxf(v[j + 3], 0.f) : v[j + 3]);
                        u.z = pack_half2(p.relu ? fmaxf(v[j + 4], 0.f) : v[j + 4], p.relu ? fmaxf(v[j + 5], 0.f) : v[j + 5]);
                        u.w = pack_half2(p.relu ? fmaxf(v[j + 6], 0.f) : v[j + 6], p.relu ? fmaxf(v[j + 7], 0.f) : v[j + 7]);
                        *reinterpret_cast<uint4*>(op + j) = u;
                    }
                } else {
                    for (int j = 0; j < 32 && n + j < p.cout; ++j) op[j] = __float2half_rn(p.relu ? fmaxf(v[j], 0.f) : v[j]);
                }
            }
            if (p.out_relu) {
                __half* op = p.out_relu + pix * p.out_stride + p.out_offset + n;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 u;
                        u.x = pack_half2(fmaxf(v[j], 0.f), fmaxf(v[j + 1], 0.f));
                        u.y = pack_half2(fmaxf(v[j + 2], 0.f), fmaxf(v[j + 3], 0.f));
                        u.z = pack_half2(fmaxf(v[j + 4], 0.f), fmaxf(v[j + 5], 0.f));
                        u.w = pack_half2(fmaxf(v[j + 6], 0.f), fmaxf(v[j + 7], 0.f));
                        *reinterpret_cast<uint4*>(op + j) = u;
                    }
                } else {
                    for (int j = 0; j < 32 && n + j < p.cout; ++j) op[j] = __float2half_rn(fmaxf(v[j], 0.f));
                }
            }
        }
    }
teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, BN);
}

template <int BN, int STAGES, bool THREE_BOX>
int launch_conv(const ConvMaps& maps, const ConvP& p, int cout_pad, cudaStream_t stream) {
    tc5_debug_init();
    static bool attr_done = false;          // one per template instantiation
    const int smem = (int)sizeof(ConvSmem<BN, STAGES>) + 1024;
    if (!attr_done) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_halo_kernel<BN, STAGES, THREE_BOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    // epilogue staging: BN/64 output boxes at the start of the halo buffers, BN/64 residual boxes at the start of the weight ring
    static_assert(A_STAGES * A_BYTES >= (BN / 64) * 128 * 128 && STAGES * BN * 128 >= (BN / 64) * 128 * 128, "rings too small for the epilogue boxes");
    dim3 grid(p.tiles_x * p.tiles_y * p.batch, cout_pad / BN, 1);
    XM_CHECK_CUDA(tc5_launch(conv_igemm_halo_kernel<BN, STAGES, THREE_BOX>, grid, dim3(192), smem, stream, maps, p));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

}  // namespace

extern "C" int xm_conv2d_nhwc_halo(const xm_conv_args_t* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    XM_REQUIRE(a, "xm_conv2d_nhwc: null args");
    XM_REQUIRE(a->n_src >= 1 && a->n_src <= 3, "xm_conv2d_nhwc: n_src must be 1..3");
    XM_REQUIRE(a->ksize == 3 && a->stride == 1, "xm_conv2d_nhwc_halo: only 3x3 stride-1 convolutions");
    XM_REQUIRE(a->batch >= 1 && a->H > 0 && a->W > 0, "xm_conv2d_nhwc: bad shape");
    XM_REQUIRE(a->cout >= 1 && a->cout_pad >= a->cout && a->cout_pad % 64 == 0, "xm_conv2d_nhwc: cout_pad must be a multiple of 64 >= cout");
    XM_REQUIRE(a->weight && a->bias && (a->out || a->out_relu), "xm_conv2d_nhwc: null weight/bias/out");
    XM_REQUIRE(a->out_stride >= a->out_offset + a->cout, "xm_conv2d_nhwc: out_stride too small");
    if (a->stride == 2) {
        XM_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "xm_conv2d_nhwc: stride-2 needs even H, W");
        XM_REQUIRE(a->batch == 1, "xm_conv2d_nhwc: stride-2 convolutions are launched one image at a time");
    }
    ConvP p;
    p.n_src = a->n_src;
    p.cin_total = 0;
    for (int s = 0; s < 3; ++s) { p.cblocks[s] = 0; p.choff[s] = 0; p.bcast[s] = 0; }
    for (int s = 0; s < a->n_src; ++s) {
        XM_REQUIRE(a->src[s].ptr && a->src[s].channels > 0 && a->src[s].channels % 64 == 0,
                   "xm_conv2d_nhwc: source %d channels must be a positive multiple of 64", s);
        p.cblocks[s] = a->src[s].channels / 64;
        p.choff[s] = p.cin_total;
        p.bcast[s] = a->src[s].broadcast;
        p.cin_total += a->src[s].channels;
    }
    p.ksize = a->ksize; p.stride = a->stride; p.pad = a->ksize / 2;
    p.Ho = a->H / a->stride; p.Wo = a->W / a->stride; p.batch = a->batch;
    p.tw = 8; p.th = 16;          // 16 groups of 8 consecutive pixels: what the tap views of the haloed tile need
    p.tiles_x = (p.Wo + p.tw - 1) / p.tw; p.tiles_y = (p.Ho + p.th - 1) / p.th;
    p.cout = a->cout; p.relu = a->relu; p.bias = a->bias;
    p.residual = (const __half*)a->residual; p.residual_bcast = a->residual_broadcast; p.residual_stride = a->cout;
    p.out = (__half*)a->out; p.out_relu = (__half*)a->out_relu; p.out_stride = a->out_stride; p.out_offset = a->out_offset;

    static int three_box = -1;
    if (three_box < 0) { const char* e = getenv("XMEM_HALO_3BOX"); three_box = (e && e[0] == '1') ? 1 : 0; }
    ConvMaps maps;
    for (int s = 0; s < 3; ++s) {
        const int ss = s < a->n_src ? s : 0;
        const uint64_t C = a->src[ss].channels;
        const uint64_t nb = a->src[ss].broadcast ? 1 : a->batch;
        uint64_t d[4] = {C, (uint64_t)a->W, (uint64_t)a->H, nb};
        uint64_t st[3] = {C * 2, (uint64_t)a->W * C * 2, (uint64_t)a->H * a->W * C * 2};
        uint32_t bx[4] = {64, (uint32_t)(three_box ? 8 : HALO_W), (uint32_t)HALO_H, 1};          // haloed tile (36 KB) or one kw box (18 KB)
        if (xm_make_tmap_f16(&maps.a[s], a->src[ss].ptr, 4, d, st, bx)) return XM_ERR_CUDA;
    }
    // debug/tuning override: XMEM_CONV_FORCE="bn,splits,stages" (0 = keep the heuristic)
    static int f_bn = -1, f_split = 0, f_depth = 0;
    if (f_bn < 0) {
        f_bn = 0;
        if (const char* e = getenv("XMEM_CONV_FORCE")) sscanf(e, "%d,%d,%d", &f_bn, &f_split, &f_depth);
    }
    // N tile: 128 couts when that already gives at least half a wave of CTAs, else 64 (more CTAs beat split-K: the
    // split-K fix-up costs ~10 us, measured in profiles/r1_conv_config_sweep.txt)
    const int sms_ = xm_num_sms();
    int BN = (a->cout_pad % 128 == 0) ? 128 : 64;
    int cb_all = 0;
    for (int s = 0; s < p.n_src; ++s) cb_all += p.cblocks[s];
    const int ksteps_all = a->ksize * a->ksize * cb_all;
    if (BN == 128 && ksteps_all <= 128 && p.tiles_x * p.tiles_y * p.batch * (a->cout_pad / 128) * 2 <= sms_) BN = 64;
    if (f_bn == 64 || (f_bn == 128 && a->cout_pad % 128 == 0)) BN = f_bn;
    {
        const uint64_t K = (uint64_t)a->ksize * a->ksize * p.cin_total;
        uint64_t d[2] = {K, (uint64_t)a->cout_pad};
        uint64_t st[1] = {K * 2};
        uint32_t bx[2] = {64, (uint32_t)BN};
        if (xm_make_tmap_f16(&maps.w, a->weight, 2, d, st, bx)) return XM_ERR_CUDA;
    }
    // TMA epilogue: whole 64-channel boxes, a single plain output (no second ReLU'd copy), 16-byte aligned channel offsets
    p.tma_epilogue = (a->cout % 64 == 0 && a->out && !a->out_relu && a->out_offset % 8 == 0 && a->out_stride % 8 == 0) ? 1 : 0;
    {
        const void* obase = p.tma_epilogue ? a->out : a->src[0].ptr;
        const uint64_t OC = p.tma_epilogue ? (uint64_t)a->out_stride : (uint64_t)a->src[0].channels;
        const uint64_t OW = p.tma_epilogue ? (uint64_t)p.Wo : (uint64_t)a->W, OH = p.tma_epilogue ? (uint64_t)p.Ho : (uint64_t)a->H;
        uint64_t d[4] = {OC, OW, OH, (uint64_t)(p.tma_epilogue ? a->batch : 1)};
        uint64_t st[3] = {OC * 2, OW * OC * 2, OH * OW * OC * 2};
        uint32_t bx[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
        if (xm_make_tmap_f16(&maps.o, obase, 4, d, st, bx)) return XM_ERR_CUDA;
        if (p.tma_epilogue && a->residual) {
            uint64_t dr[4] = {(uint64_t)a->cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)(a->residual_broadcast ? 1 : a->batch)};
            uint64_t sr[3] = {(uint64_t)a->cout * 2, (uint64_t)p.Wo * a->cout * 2, (uint64_t)p.Ho * p.Wo * a->cout * 2};
            if (xm_make_tmap_f16(&maps.r, a->residual, 4, dr, sr, bx)) return XM_ERR_CUDA;
        } else {
            maps.r = maps.o;
        }
    }
    // weight ring: 8 KB (BN=64) / 16 KB (BN=128) tiles; 8 / 6 deep keeps ~64-96 KB of weights in flight next to the two halo buffers
    (void)f_split; (void)f_depth;
    if (three_box) {
        if (BN == 128) return launch_conv<128, 6, true>(maps, p, a->cout_pad, stream);
        return launch_conv<64, 8, true>(maps, p, a->cout_pad, stream);
    }
    if (BN == 128) return launch_conv<128, 6, false>(maps, p, a->cout_pad, stream);
    return launch_conv<64, 8, false>(maps, p, a->cout_pad, stream);
}
