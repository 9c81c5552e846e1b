// conv_igemm_pair.cu — the production implicit-GEMM convolution (../conv_igemm.cu) on CTA PAIRS (tcgen05 cta_group::2):
// two CTAs of a cluster (2,1,1) own two neighbouring 128-pixel tiles and the SAME BN output channels; the leader issues
// one M=256 x N=BN MMA per 16 channels that reads A (128 pixel rows) from each CTA's own shared memory and B from BOTH
// (each CTA holds BN/2 weight rows), and writes each CTA's 128 x BN accumulator into that CTA's TMEM.  Per K-step an SM
// ingests 16 KB (A) + BN*64 B (half of B) instead of 16 KB + BN*128 B, and BN may be 256:
//     today   128x128 tile : 32 KB per 2.1 MFLOP per SM
//     here    BN=128       : 24 KB per 2.1 MFLOP per SM   (1.33x)
//             BN=256       : 32 KB per 4.2 MFLOP per SM   (2x)
// which is what the layers that fill the machine need: they are bound by the ~65-90 GB/s an SM can pull from L2
// (NOTES.md).  Protocol (the one CUTLASS' sm100 2SM pipeline uses):
//   * TMEM is allocated/freed with cta_group::2 by warp 1 of BOTH CTAs; cluster barrier after the barrier init and
//     before the free;
//   * both producers wait on their OWN empty[s] and issue `cp.async.bulk.tensor ... .cta_group::2` loads into their own
//     stage buffers whose complete_tx goes to the LEADER's full[s] (mbarrier address with the peer bit 24 cleared); the
//     leader's producer arms full[s] with the bytes of both CTAs;
//   * the leader's MMA thread waits full[s], issues 4 x tcgen05.mma.cta_group::2 and commits with
//     `.multicast::cluster` mask 0b11 to empty[s] (and finally `done`) of both CTAs;
//   * each CTA runs the normal epilogue on its own TMEM lanes (bias / residual / ReLU / TMA store).
// No split-K here.  cout_pad must be a multiple of BN (128 or 256).  An odd number of pixel tiles is padded with a CTA
// whose loads are zero-filled (batch coordinate out of range) and whose stores are clipped.
// Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -c conv_igemm_pair.cu
//
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
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

namespace {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// pair-wide TMEM management: executed by the same warp index in both CTAs
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// loads into THIS CTA's shared memory; the transaction bytes are credited to the LEADER's mbarrier at the same offset
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// D[tmem of both CTAs] (+)= A[both CTAs' smem, 128 rows each] * B[both CTAs' smem, N/2 rows each]^T; leader thread only
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the same-offset mbarrier of both CTAs once all previously issued pair MMAs have completed
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
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
    alignas(1024) uint8_t a[CONV_STAGES][128 * 128];
    alignas(1024) uint8_t b[CONV_STAGES][(BN / 2) * 128];      // this CTA's half of the weight rows
    alignas(8) uint64_t full[CONV_STAGES];
    uint64_t empty[CONV_STAGES];
    uint64_t done;
    uint64_t resbar;
    uint32_t tmem_base;
    float bias[BN];
};

template <int BN, int CONV_STAGES>
__global__ void __launch_bounds__(192)
conv_igemm_2cta_kernel(const __grid_constant__ ConvMaps maps, const ConvP p) {
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

    const int taps = p.ksize * p.ksize;
    int cb_total = 0;
    for (int s = 0; s < p.n_src; ++s) cb_total += p.cblocks[s];
    const int ksteps_all = taps * cb_total;
    const int k_begin = 0;
    const int ksteps = ksteps_all;

    if (threadIdx.x == 0) {
        for (int i = 0; i < CONV_STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        mbar_init(&sm.done, 1);
        mbar_init(&sm.resbar, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc_2sm(&sm.tmem_base, BN); tmem_relinquish_2sm(); }
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
    cluster_sync_all();            // the leader's mbarriers must be initialised before the peer's loads/commits reach them
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t crank = cluster_ctarank();      // 0 = leader (issues the MMAs), 1 = peer
    pdl_wait();                    // inputs (and the split-K workspace) come from preceding kernels
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&maps.a[s]);
            for (int it = 0; it < ksteps; ++it) {
                const int kk = k_begin + it;
                const int tap = kk / cb_total;
                int cb = kk - tap * cb_total;
                int s = 0;
                if (p.n_src > 1 && cb >= p.cblocks[0]) { cb -= p.cblocks[0]; s = 1; }
                if (p.n_src > 2 && s == 1 && cb >= p.cblocks[1]) { cb -= p.cblocks[1]; s = 2; }
                const int kh = tap / p.ksize, kw = tap % p.ksize;
                const int bb = p.bcast[s] ? 0 : b;
                const int st = it % CONV_STAGES, ph = (it / CONV_STAGES) & 1;
                mbar_wait(&sm.empty[st], ph ^ 1, 21);
                if (crank == 0) mbar_expect_tx(&sm.full[st], 2 * (128 * 128 + (BN / 2) * 128));     // both CTAs' A + B halves
                if (p.stride == 1) {
                    tma_load_4d_2sm(sm.a[st], &maps.a[s], &sm.full[st], cb * 64, x0 + kw - p.pad, y0 + kh - p.pad, bb);
                } else {
                    // input pixel (2*yo + kh - pad, 2*xo + kw - pad) -> parity plane + half coordinate
                    const int dy = kh - p.pad, dx = kw - p.pad;
                    const int py = dy & 1, px = dx & 1;
                    const int hy = (dy - py) / 2, hx = (dx - px) / 2;
                    tma_load_5d_2sm(sm.a[st], &maps.a[s], &sm.full[st], cb * 64, px, x0 + hx, py, y0 + hy);
                }
                tma_load_2d_2sm(sm.b[st], &maps.w, &sm.full[st], tap * p.cin_total + p.choff[s] + cb * 64, n0 + (int)crank * (BN / 2));
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(256, BN);
            for (int it = 0; it < ksteps; ++it) {
                const int st = it % CONV_STAGES, ph = (it / CONV_STAGES) & 1;
                mbar_wait(&sm.full[st], ph, 22);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t a = make_desc_sw128(smem_u32(sm.a[st]) + j * 32);
                    uint64_t bd = make_desc_sw128(smem_u32(sm.b[st]) + j * 32);
                    mma_f16_ss_2sm(tmem, a, bd, idesc, (it | j) ? 1u : 0u);
                }
                mma_commit_2sm(&sm.empty[st]);
            }
            mma_commit_2sm(&sm.done);
        }
    } else {
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int yo = y0 + row / p.tw, xo = x0 + row % p.tw;
        const bool pix_ok = (yo < p.Ho) && (xo < p.Wo) && (b < p.batch);      // b == batch: the padding CTA of an odd grid
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
    cluster_sync_all();            // the pair's MMAs/commits/loads touch both CTAs: leave (and free TMEM) together
    if (warp == 1) tmem_dealloc_2sm(tmem, BN);
}

template <int BN, int STAGES>
int launch_conv(const ConvMaps& maps, const ConvP& p, int cout_pad, cudaStream_t stream) {
    tc5_debug_init();
    static XmPerDevice attr_token = {0};
    const int smem = (int)sizeof(ConvSmem<BN, STAGES>) + 1024;
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    // epilogue staging: BN/64 output boxes in the A ring, BN/64 residual boxes in the B ring (16 KB each)
    static_assert(STAGES * 128 * 128 >= (BN / 64) * 128 * 128 && STAGES * (BN / 2) * 128 >= (BN / 64) * 128 * 128, "stage rings too small for the epilogue boxes");
    dim3 grid((p.tiles_x * p.tiles_y * p.batch + 1) / 2 * 2, cout_pad / BN, 1);     // whole pairs along the pixel-tile axis
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = xm_pdl_enabled() ? 2 : 1;
    XM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_2cta_kernel<BN, STAGES>, maps, p));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

}  // namespace

int xm_conv2d_pair(const xm_conv_args_t* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    XM_REQUIRE(a, "xm_conv2d_nhwc: null args");
    XM_REQUIRE(a->n_src >= 1 && a->n_src <= 3, "xm_conv2d_nhwc: n_src must be 1..3");
    XM_REQUIRE(a->ksize == 1 || a->ksize == 3, "xm_conv2d_nhwc: ksize must be 1 or 3");
    XM_REQUIRE(a->stride == 1 || a->stride == 2, "xm_conv2d_nhwc: stride must be 1 or 2");
    XM_REQUIRE(a->batch >= 1 && a->H > 0 && a->W > 0, "xm_conv2d_nhwc: bad shape");
    XM_REQUIRE(a->cout >= 1 && a->cout_pad >= a->cout && a->cout_pad % 128 == 0, "xm_conv2d_nhwc_2cta: cout_pad must be a multiple of 128 >= cout");
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
    // tile rectangle: minimise padded area
    int best_tw = 16; long best = -1;
    for (int tw = 8; tw <= 32; tw *= 2) {
        const int th = 128 / tw;
        long area = (long)((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
        if (best < 0 || area < best) { best = area; best_tw = tw; }
    }
    p.tw = best_tw; p.th = 128 / best_tw;
    p.tiles_x = (p.Wo + p.tw - 1) / p.tw; p.tiles_y = (p.Ho + p.th - 1) / p.th;
    p.cout = a->cout; p.relu = a->relu; p.bias = a->bias;
    p.residual = (const __half*)a->residual; p.residual_bcast = a->residual_broadcast; p.residual_stride = a->cout;
    p.out = (__half*)a->out; p.out_relu = (__half*)a->out_relu; p.out_stride = a->out_stride; p.out_offset = a->out_offset;

    ConvMaps maps;
    for (int s = 0; s < 3; ++s) {
        const int ss = s < a->n_src ? s : 0;
        const uint64_t C = a->src[ss].channels;
        const uint64_t nb = a->src[ss].broadcast ? 1 : a->batch;
        if (a->stride == 1) {
            uint64_t d[4] = {C, (uint64_t)a->W, (uint64_t)a->H, nb};
            uint64_t st[3] = {C * 2, (uint64_t)a->W * C * 2, (uint64_t)a->H * a->W * C * 2};
            uint32_t bx[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
            if (xm_make_tmap_f16(&maps.a[s], a->src[ss].ptr, 4, d, st, bx)) return XM_ERR_CUDA;
        } else {
            uint64_t d[5] = {C, 2, (uint64_t)a->W / 2, 2, (uint64_t)a->H / 2};
            uint64_t st[4] = {C * 2, 2 * C * 2, (uint64_t)a->W * C * 2, 2 * (uint64_t)a->W * C * 2};
            uint32_t bx[5] = {64, 1, (uint32_t)p.tw, 1, (uint32_t)p.th};
            if (xm_make_tmap_f16(&maps.a[s], a->src[ss].ptr, 5, d, st, bx)) return XM_ERR_CUDA;
        }
    }
    // debug/tuning override: XMEM_CONV_FORCE="bn,splits,stages" (0 = keep the heuristic)
    static int f_bn = -1, f_split = 0, f_depth = 0;
    if (f_bn < 0) {
        f_bn = 0;
        if (const char* e = getenv("XMEM_CONV_FORCE")) sscanf(e, "%d,%d,%d", &f_bn, &f_split, &f_depth);
    }
    // N tile of the PAIR: 256 output channels when the layer has them (2x the FLOPs per ingested byte), else 128
    int BN = (a->cout_pad % 256 == 0) ? 256 : 128;
    if (f_bn == 128 || (f_bn == 256 && a->cout_pad % 256 == 0)) BN = f_bn;
    {
        const uint64_t K = (uint64_t)a->ksize * a->ksize * p.cin_total;
        uint64_t d[2] = {K, (uint64_t)a->cout_pad};
        uint64_t st[1] = {K * 2};
        uint32_t bx[2] = {64, (uint32_t)BN / 2};          // each CTA of the pair fetches half of the rows
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
    // pipeline depth: one CTA per SM either way (>= 96 KB of stages); 6 stages when the ring fits, 4 for short K loops
    int cb_total = 0;
    for (int s = 0; s < p.n_src; ++s) cb_total += p.cblocks[s];
    const int ksteps = a->ksize * a->ksize * cb_total;
    int depth = (ksteps <= 8) ? 4 : 6;
    if (f_depth == 4 || f_depth == 6) depth = f_depth;
    (void)f_split;
    if (BN == 256) {
        if (depth == 4) return launch_conv<256, 4>(maps, p, a->cout_pad, stream);
        return launch_conv<256, 6>(maps, p, a->cout_pad, stream);
    }
    if (depth == 4) return launch_conv<128, 4>(maps, p, a->cout_pad, stream);
    return launch_conv<128, 6>(maps, p, a->cout_pad, stream);
}
