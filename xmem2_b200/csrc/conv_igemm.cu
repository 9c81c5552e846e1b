// conv_igemm.cu — implicit-GEMM convolution on tcgen05 for NHWC fp16 activations, with CLUSTER split-K.
//
// Small grids (the 30x54 layers cover 28-112 of the 148 SMs) are bound by what each SM can pull through TMA (~57 GB/s per
// SM measured, tests/microbench/tma_ingest.cu), so the only way to go faster is to put more SMs on the layer: the S <= 3
// CTAs that share an output tile (blockIdx.z = 0..S-1) form a thread-block cluster (1,1,S); every CTA accumulates its K
// range in its own TMEM, the peers (rank > 0) push their fp32 accumulators into the leader's shared memory through DSMEM
// (st.shared::cluster), and the leader adds them in rank order and runs the usual epilogue.  No global workspace, no atomics.
// Protocol after the main loop (all 192 threads of every CTA execute both cluster barriers):
//   epilogue warps wait `done` (their CTA's MMAs finished)  ->  cluster barrier #1: every CTA's stage buffers are free
//   peers: TMEM -> registers -> leader's smem partial[rank-1] (layout [BN/4][128 rows] float4: conflict-free)
//   cluster barrier #2 (release/acquire): partials visible to the leader; peers leave
//   leader: own TMEM + partial[0] (+ partial[1]) -> bias / residual / ReLU -> TMA store (or the plain store path)
// Layers that fill more than one wave of 128x128 tiles but fit one wave of CTA pairs go to conv_igemm_pair.cu
// (cta_group::2, M = 256, each CTA loads half of the weight tile: half the bytes per FLOP through each SM's TMA).
// Measured on B200 (round 2, python -m xmem2_b200.util.conv_bench): 905 -> 853 us per 480p frame against the round-1 kernel with its global
// split-K; the pair kernel takes the 60x108 512->512 decoder convolution from 57 to 32 us.
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

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local_smem_addr` in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32x4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct alignas(64) ConvMaps {
    CUtensorMap a[3];
    CUtensorMap w;
    CUtensorMap o;      // output   [out_stride, Wo, Ho, B]   box [64, TW, TH, 1]   (TMA-store epilogue)
    CUtensorMap r;      // residual [cout, Wo, Ho, B|1]       box [64, TW, TH, 1]
    CUtensorMap o2;     // second, ReLU'd output (same geometry as o) when relu_copy_tma
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
    int splits, ksteps_per_split;     // cluster split-K: blockIdx.z (= cluster rank) owns k-steps [z*kps, min((z+1)*kps, ksteps))
    int tma_epilogue;                 // 1: stage the tile in swizzled smem, residual in / output out through TMA
    int relu_copy_tma;                // 1: the TMA epilogue also writes max(x, 0) to out_relu (splits == 1, >= 3 stages at BN = 128)
};

template <int BN, int CONV_STAGES>
struct ConvSmem {
    alignas(1024) uint8_t a[CONV_STAGES][128 * 128];
    alignas(1024) uint8_t b[CONV_STAGES][BN * 128];
    alignas(8) uint64_t full[CONV_STAGES];
    uint64_t empty[CONV_STAGES];
    uint64_t done;
    uint64_t resbar;
    uint32_t tmem_base;
    float bias[BN];
};

template <int BN, int CONV_STAGES>
__global__ void __launch_bounds__(192)
conv_igemm_csk_kernel(const __grid_constant__ ConvMaps maps, const ConvP p) {
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
    const int k_begin = blockIdx.z * p.ksteps_per_split;
    const int k_end = min(ksteps_all, k_begin + p.ksteps_per_split);
    const int ksteps = k_end - k_begin;

    if (threadIdx.x == 0) {
        for (int i = 0; i < CONV_STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
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
                mbar_expect_tx(&sm.full[st], 128 * 128 + BN * 128);
                if (p.stride == 1) {
                    tma_load_4d(sm.a[st], &maps.a[s], &sm.full[st], cb * 64, x0 + kw - p.pad, y0 + kh - p.pad, bb);
                } else {
                    // input pixel (2*yo + kh - pad, 2*xo + kw - pad) -> parity plane + half coordinate
                    const int dy = kh - p.pad, dx = kw - p.pad;
                    const int py = dy & 1, px = dx & 1;
                    const int hy = (dy - py) / 2, hx = (dx - px) / 2;
                    tma_load_5d(sm.a[st], &maps.a[s], &sm.full[st], cb * 64, px, x0 + hx, py, y0 + hy);
                }
                tma_load_2d(sm.b[st], &maps.w, &sm.full[st], tap * p.cin_total + p.choff[s] + cb * 64, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, BN);
            for (int it = 0; it < ksteps; ++it) {
                const int st = it % CONV_STAGES, ph = (it / CONV_STAGES) & 1;
                mbar_wait(&sm.full[st], ph, 22);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t a = make_desc_sw128(smem_u32(sm.a[st]) + j * 32);
                    uint64_t bd = make_desc_sw128(smem_u32(sm.b[st]) + j * 32);
                    mma_f16_ss(tmem, a, bd, idesc, (it | j) ? 1u : 0u);
                }
                mma_commit(&sm.empty[st]);
            }
            mma_commit(&sm.done);
        }
    } else {
        mbar_wait(&sm.done, 0, 23);        // this CTA's accumulator is complete, its stage buffers are idle
        tc_fence_after();
    }
    __syncwarp();                          // warps 0/1 re-converge before the aligned cluster barriers
    // ---------------------------------------------------------------- cluster split-K fix-up
    // partial[peer][BN/4][128] float4 in the LEADER's A ring, behind the (BN/64) TMA-store boxes
    const uint32_t crank = (p.splits > 1) ? cluster_ctarank() : 0u;
    uint8_t* const partial0 = &sm.a[0][0] + (BN / 64) * 128 * 128;
    if (p.splits > 1) {
        cluster_sync_all();                // #1: all MMAs of the cluster retired -> the leader's ring may be overwritten
        if (warp >= 2 && crank > 0) {
            const int lane_base = (warp & 3) * 32;
            const int row = lane_base + lane;
            const uint32_t dst0 = map_to_cta(smem_u32(partial0 + (size_t)(crank - 1) * 128 * BN * 4), 0);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    st_cluster_f32x4(dst0 + (uint32_t)(((c0 + j) / 4 * 128 + row) * 16), r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
        cluster_sync_all();                // #2: partials are visible in the leader's shared memory
    }
    if (warp >= 2 && crank == 0) {
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int yo = y0 + row / p.tw, xo = x0 + row % p.tw;
        const bool pix_ok = (yo < p.Ho) && (xo < p.Wo);
        const size_t pix = ((size_t)b * p.Ho + yo) * p.Wo + xo;
        const size_t rpix = ((size_t)(p.residual_bcast ? 0 : b) * p.Ho + yo) * p.Wo + xo;
        const int n_peers = p.splits - 1;
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
                    for (int z = 0; z < n_peers; ++z) {          // fixed order: deterministic sums
                        const uint8_t* src = partial0 + (size_t)z * 128 * BN * 4;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 f = *reinterpret_cast<const float4*>(src + ((c0 + j) / 4 * 128 + row) * 16);
                            v[32 * hlf + j] += f.x; v[32 * hlf + j + 1] += f.y; v[32 * hlf + j + 2] += f.z; v[32 * hlf + j + 3] += f.w;
                        }
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
                if (p.relu_copy_tma) {
                    // second output max(x, 0) (GroupResBlock hands both g and relu(g) on): box k lives in a pipeline stage that
                    // neither the output nor the residual staging uses (BN = 64: a[1]; BN = 128: a[2], b[2])
                    uint8_t* rbase = (BN == 64) ? (&sm.a[0][0] + 128 * 128) : (k == 0 ? (&sm.a[0][0] + 2 * 128 * 128) : (&sm.b[0][0] + 2 * BN * 128));
                    uint8_t* rrow2 = rbase + row * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4 u;
                        u.x = pack_half2(fmaxf(v[8 * c], 0.f), fmaxf(v[8 * c + 1], 0.f)); u.y = pack_half2(fmaxf(v[8 * c + 2], 0.f), fmaxf(v[8 * c + 3], 0.f));
                        u.z = pack_half2(fmaxf(v[8 * c + 4], 0.f), fmaxf(v[8 * c + 5], 0.f)); u.w = pack_half2(fmaxf(v[8 * c + 6], 0.f), fmaxf(v[8 * c + 7], 0.f));
                        *reinterpret_cast<uint4*>(rrow2 + ((c ^ (row & 7)) << 4)) = u;
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
                if (p.relu_copy_tma) {
                    for (int k = 0; k < nbox; ++k) {
                        const uint8_t* rbase = (BN == 64) ? (&sm.a[0][0] + 128 * 128) : (k == 0 ? (&sm.a[0][0] + 2 * 128 * 128) : (&sm.b[0][0] + 2 * BN * 128));
                        tma_store_4d(&maps.o2, rbase, p.out_offset + n0 + 64 * k, x0, y0, b);
                    }
                }
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
            for (int z = 0; z < n_peers; ++z) {
                const uint8_t* src = partial0 + (size_t)z * 128 * BN * 4;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 f = *reinterpret_cast<const float4*>(src + ((c0 + j) / 4 * 128 + row) * 16);
                    acc[j] += f.x; acc[j + 1] += f.y; acc[j + 2] += f.z; acc[j + 3] += f.w;
                }
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

template <int BN, int STAGES>
int launch_conv(const ConvMaps& maps, const ConvP& p, int cout_pad, cudaStream_t stream) {
    tc5_debug_init();
    static XmPerDevice attr_token = {0};
    const int smem = (int)sizeof(ConvSmem<BN, STAGES>) + 1024;
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_csk_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    dim3 grid(p.tiles_x * p.tiles_y * p.batch, cout_pad / BN, p.splits);
    // the splits of one output tile are one cluster; the partial area must fit behind the TMA-store boxes of the A ring
    static_assert(STAGES * 128 * 128 >= (BN / 64) * 128 * 128, "A ring smaller than the store boxes");
    if ((int64_t)(p.splits - 1) * 128 * BN * 4 > (int64_t)(STAGES - BN / 64) * 128 * 128) return XM_ERR_ARG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = p.splits;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = xm_pdl_enabled() ? 2 : 1;
    XM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_csk_kernel<BN, STAGES>, maps, p));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

}  // namespace

extern "C" int xm_conv2d_nhwc(const xm_conv_args_t* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    XM_REQUIRE(a, "xm_conv2d_nhwc: null args");
    XM_REQUIRE(a->n_src >= 1 && a->n_src <= 3, "xm_conv2d_nhwc: n_src must be 1..3");
    XM_REQUIRE(a->ksize == 1 || a->ksize == 3, "xm_conv2d_nhwc: ksize must be 1 or 3");
    XM_REQUIRE(a->stride == 1 || a->stride == 2, "xm_conv2d_nhwc: stride must be 1 or 2");
    XM_REQUIRE(a->batch >= 1 && a->H > 0 && a->W > 0, "xm_conv2d_nhwc: bad shape");
    XM_REQUIRE(a->cout >= 1 && a->cout_pad >= a->cout && a->cout_pad % 64 == 0, "xm_conv2d_nhwc: cout_pad must be a multiple of 64 >= cout");
    XM_REQUIRE(a->weight && a->bias && (a->out || a->out_relu), "xm_conv2d_nhwc: null weight/bias/out");
    XM_REQUIRE(a->out_stride >= a->out_offset + a->cout, "xm_conv2d_nhwc: out_stride too small");
    {
        // CTA pairs (conv_igemm_pair.cu) when the 128x128 tiling needs more than one wave but pairs with BN = 256 fit in one:
        // each SM then ingests half the weight bytes per FLOP (the big layers are bound by per-SM TMA ingest)
        const int Ho = a->H / (a->stride ? a->stride : 1), Wo = a->W / (a->stride ? a->stride : 1);
        int best = -1;
        for (int tw = 8; tw <= 32; tw *= 2) {
            const int th = 128 / tw;
            const int area = ((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
            if (best < 0 || area < best) best = area;
        }
        const int mtiles = best * a->batch;
        const int sms = xm_num_sms();
        if (a->stride == 1 && a->cout_pad % 256 == 0 && a->cout % 64 == 0 && mtiles * (a->cout_pad / 128) > sms &&
            ((mtiles + 1) / 2) * 2 * (a->cout_pad / 256) <= sms)
            return xm_conv2d_pair(a, stream_);
    }
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
    // N tile: 128 couts when that already gives at least half a wave of CTAs, else 64 (more CTAs beat split-K: the
    // split-K fix-up costs ~10 us, measured in profiles/r1_conv_config_sweep.txt)
    const int sms_ = xm_num_sms();
    int BN = (a->cout_pad % 128 == 0) ? 128 : 64;
    int cb_all = 0;
    for (int s = 0; s < p.n_src; ++s) cb_all += p.cblocks[s];
    const int ksteps_all = a->ksize * a->ksize * cb_all;
    // (short K loops only: with >= 49 k-steps a 128-wide tile split in two over a cluster pulls 1.5x fewer bytes per FLOP through
    // each SM than 64-wide tiles and the DSMEM reduction is cheap -- 30x54 512->512: 26.1 -> 18.8 us)
    if (BN == 128 && ksteps_all <= 48 && p.tiles_x * p.tiles_y * p.batch * (a->cout_pad / 128) * 2 <= sms_) BN = 64;
    if (f_bn == 64 || (f_bn == 128 && a->cout_pad % 128 == 0)) BN = f_bn;
    {
        const uint64_t K = (uint64_t)a->ksize * a->ksize * p.cin_total;
        uint64_t d[2] = {K, (uint64_t)a->cout_pad};
        uint64_t st[1] = {K * 2};
        uint32_t bx[2] = {64, (uint32_t)BN};
        if (xm_make_tmap_f16(&maps.w, a->weight, 2, d, st, bx)) return XM_ERR_CUDA;
    }
    // occupancy plan: many CTAs -> 3 stages (2 CTAs/SM overlap prologue/epilogue); few CTAs -> 6 stages (hide L2
    // latency in the k-loop) and split-K over blockIdx.z so that idle SMs share the reduction.
    int cb_total = 0;
    for (int s = 0; s < p.n_src; ++s) cb_total += p.cblocks[s];
    const int ksteps = a->ksize * a->ksize * cb_total;
    const int ctas = p.tiles_x * p.tiles_y * p.batch * (a->cout_pad / BN);
    const int sms = xm_num_sms();
    // cluster split-K: put idle SMs on the layer (each CTA's K-step pulls 16 KB + BN*128 B through ITS SM's L2 port);
    // every split keeps at least 8 k-steps, whole clusters must fit in one wave
    p.splits = 1;
    const int max_split = (BN == 64) ? 3 : 2;            // room in the leader's A ring (6 stages), see the kernel header
    while (p.splits < max_split && ctas * (p.splits + 1) <= sms && ksteps / (p.splits + 1) >= 8) ++p.splits;
    if (f_split > 0) p.splits = f_split > max_split ? max_split : f_split;
    if (p.splits > ksteps) p.splits = ksteps;
    p.ksteps_per_split = (ksteps + p.splits - 1) / p.splits;
    p.splits = (ksteps + p.ksteps_per_split - 1) / p.ksteps_per_split;      // no empty split: every CTA issues >= 1 MMA
    int depth = (p.splits > 1) ? 6 : ((p.ksteps_per_split <= 4) ? 2 : ((ctas < 2 * sms) ? 6 : 3));
    if (p.splits == 1 && (f_depth == 2 || f_depth == 3 || f_depth == 6)) depth = f_depth;
    // TMA epilogue: whole 64-channel boxes, 16-byte aligned channel offsets; a second ReLU'd copy only when a pipeline stage is
    // left over to stage it in (no split-K partials in the A ring, >= 3 stages at BN = 128)
    const bool relu_copy_ok = a->out_relu == nullptr || (p.splits == 1 && depth >= (BN == 64 ? 2 : 3));
    p.tma_epilogue = (a->cout % 64 == 0 && a->out && relu_copy_ok && a->out_offset % 8 == 0 && a->out_stride % 8 == 0) ? 1 : 0;
    p.relu_copy_tma = (p.tma_epilogue && a->out_relu) ? 1 : 0;
    {
        const void* obase = p.tma_epilogue ? a->out : a->src[0].ptr;
        const uint64_t OC = p.tma_epilogue ? (uint64_t)a->out_stride : (uint64_t)a->src[0].channels;
        const uint64_t OW = p.tma_epilogue ? (uint64_t)p.Wo : (uint64_t)a->W, OH = p.tma_epilogue ? (uint64_t)p.Ho : (uint64_t)a->H;
        uint64_t d[4] = {OC, OW, OH, (uint64_t)(p.tma_epilogue ? a->batch : 1)};
        uint64_t st[3] = {OC * 2, OW * OC * 2, OH * OW * OC * 2};
        uint32_t bx[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
        if (xm_make_tmap_f16(&maps.o, obase, 4, d, st, bx)) return XM_ERR_CUDA;
        if (p.relu_copy_tma) {
            if (xm_make_tmap_f16(&maps.o2, a->out_relu, 4, d, st, bx)) return XM_ERR_CUDA;
        } else {
            maps.o2 = maps.o;
        }
        if (p.tma_epilogue && a->residual) {
            uint64_t dr[4] = {(uint64_t)a->cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)(a->residual_broadcast ? 1 : a->batch)};
            uint64_t sr[3] = {(uint64_t)a->cout * 2, (uint64_t)p.Wo * a->cout * 2, (uint64_t)p.Ho * p.Wo * a->cout * 2};
            if (xm_make_tmap_f16(&maps.r, a->residual, 4, dr, sr, bx)) return XM_ERR_CUDA;
        } else {
            maps.r = maps.o;
        }
    }
    if (BN == 128) {
        if (depth == 2) return launch_conv<128, 2>(maps, p, a->cout_pad, stream);
        if (depth == 3) return launch_conv<128, 3>(maps, p, a->cout_pad, stream);
        return launch_conv<128, 6>(maps, p, a->cout_pad, stream);
    }
    if (depth == 2) return launch_conv<64, 2>(maps, p, a->cout_pad, stream);
    if (depth == 3) return launch_conv<64, 3>(maps, p, a->cout_pad, stream);
    return launch_conv<64, 6>(maps, p, a->cout_pad, stream);
}
