// pair_dissim.cu — fused cycle dissimilarity of ordered frame pairs for the annotation-candidate selector
// (reference inference/frame_selection/frame_selection.py:213-221):
//     score(A, B) = mean over [HW, HW] of relu(S_ab[x, y] - S_ba[x, y])
//     S_ab[x, y]  = similarity(memory = pixel x of A, query = pixel y of B weighted by B's selection)   (memory_util.py:7-39)
//     S_ba[x, y]  = similarity(memory = pixel x of B, query = pixel y of A weighted by A's selection)
// With the packed operands of the memory read (k1_affinity.cu: Kp = (k^2, k), Qp = (-e, 2ke), b_sq = sum e k^2):
//     S_ab[x, y] = (Qp_B[y] . Kp_A[x] - bsq_B[y]) * shrinkage_A[x] / 8
//     S_ba[x, y] = (Qp_A[y] . Kp_B[x] - bsq_A[y]) * shrinkage_B[x] / 8
// One CTA owns a 128 (y) x 128 (x) tile of BOTH matrices: 8 TMA boxes (4 operands x 2 K-halves, 128 KB), 16 tcgen05 MMAs into
// two 128-column TMEM accumulators, and an epilogue that never writes a score: each thread folds relu(S_ab - S_ba) of its row
// into one float; per-CTA partial sums go to a [pair][tile] array that a second kernel adds in a fixed order (deterministic).
// Zero padding does the masking: padded query rows have Qp = 0, bsq = 0; padded memory columns have shrinkage = 0.
// The production selector (xmem2_b200/inference/frame_selection/frame_selection.py) currently obtains the two matrices from
// the read kernel's similarity dump (2 x 10.8 MB written and re-read per pair at 480p); this kernel reads 128 KB per tile and
// writes 4 bytes.  Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -c pair_dissim.cu
#include <cstdio>
#include <cstdlib>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

namespace {

constexpr int PT = 128;            // tile edge (pixels)
constexpr int PKP = 128;           // packed operand width (fp16 elements)

struct alignas(64) PairMaps {
    CUtensorMap kp;                // [128, hw_pad, F]  box [64, 128, 1]   memory-side rows (k^2, k)
    CUtensorMap qp;                // [128, hw_pad, F]  box [64, 128, 1]   query-side rows (-e, 2ke)
};

struct PairSmem {
    alignas(1024) uint8_t qb[2][PT * 128];     // Qp of B, rows y        (A operand of D1)
    alignas(1024) uint8_t ka[2][PT * 128];     // Kp of A, rows x        (B operand of D1)
    alignas(1024) uint8_t qa[2][PT * 128];     // Qp of A, rows y        (A operand of D2)
    alignas(1024) uint8_t kb[2][PT * 128];     // Kp of B, rows x        (B operand of D2)
    float ms_a[PT], ms_b[PT];
    float red[4];
    alignas(8) uint64_t full;
    uint64_t done;
    uint32_t tmem_base;
};

__device__ __forceinline__ float score2(uint32_t acc_bits, float bsq8, float ms) {
    return fmaf(__uint_as_float(acc_bits), 0.125f, -bsq8) * ms;       // same expression as k1_affinity.cu
}

__global__ void __launch_bounds__(192, 1)
pair_dissim_kernel(const __grid_constant__ PairMaps maps, const float* __restrict__ bsq_all, const float* __restrict__ ms_all, int hw_pad,
                   const int* __restrict__ pair_a, const int* __restrict__ pair_b, float* __restrict__ partial) {
    extern __shared__ uint8_t smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ytile = blockIdx.x, xtile = blockIdx.y, pair = blockIdx.z;

    if (threadIdx.x == 0) {
        mbar_init(&sm.full, 1);
        mbar_init(&sm.done, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, 256); tmem_relinquish(); }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&maps.kp); tma_prefetch_desc(&maps.qp); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    pdl_wait();
    pdl_launch_dependents();
    const int fa = pair_a[pair], fb = pair_b[pair];

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&sm.full, 8 * PT * 128);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tma_load_3d(sm.qb[h], &maps.qp, &sm.full, 64 * h, ytile * PT, fb);
                tma_load_3d(sm.ka[h], &maps.kp, &sm.full, 64 * h, xtile * PT, fa);
                tma_load_3d(sm.qa[h], &maps.qp, &sm.full, 64 * h, ytile * PT, fa);
                tma_load_3d(sm.kb[h], &maps.kp, &sm.full, 64 * h, xtile * PT, fb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(PT, PT);
            mbar_wait(&sm.full, 0, 41);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_f16_ss(tmem, make_desc_sw128(smem_u32(sm.qb[h]) + j * 32), make_desc_sw128(smem_u32(sm.ka[h]) + j * 32), idesc, (h | j) ? 1u : 0u);
                    mma_f16_ss(tmem + PT, make_desc_sw128(smem_u32(sm.qa[h]) + j * 32), make_desc_sw128(smem_u32(sm.kb[h]) + j * 32), idesc, (h | j) ? 1u : 0u);
                }
            mma_commit(&sm.done);
        }
    } else {
        const int t = threadIdx.x - 64;                    // 0..127
        sm.ms_a[t] = __ldg(ms_all + (size_t)fa * hw_pad + xtile * PT + t);
        sm.ms_b[t] = __ldg(ms_all + (size_t)fb * hw_pad + xtile * PT + t);
        const int lane_base = (warp & 3) * 32;
        const int y = ytile * PT + lane_base + lane;
        const float bsq_b8 = __ldg(bsq_all + (size_t)fb * hw_pad + y) * 0.125f;
        const float bsq_a8 = __ldg(bsq_all + (size_t)fa * hw_pad + y) * 0.125f;
        asm volatile("bar.sync 1, 128;" ::: "memory");     // ms_a / ms_b visible to the four epilogue warps
        mbar_wait(&sm.done, 0, 42);
        tc_fence_after();
        float sum = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < PT; c0 += 32) {
            uint32_t r1[32], r2[32];
            tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + c0, r1);
            tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + PT + c0, r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float s_ab = score2(r1[j], bsq_b8, sm.ms_a[c0 + j]);
                const float s_ba = score2(r2[j], bsq_a8, sm.ms_b[c0 + j]);
                sum += fmaxf(s_ab - s_ba, 0.f);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sm.red[warp & 3] = sum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64)
            partial[((size_t)pair * gridDim.x + ytile) * gridDim.y + xtile] = (sm.red[0] + sm.red[1]) + (sm.red[2] + sm.red[3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// one warp per pair: fixed-order sum of the pair's tile partials, scaled to the mean over [hw, hw]
__global__ void pair_reduce_kernel(const float* __restrict__ partial, int tiles, int n_pairs, float inv_count, float* __restrict__ scores) {
    pdl_wait();
    pdl_launch_dependents();
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= n_pairs) return;
    float s = 0.f;
    for (int i = lane; i < tiles; i += 32) s += partial[(size_t)pair * tiles + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) scores[pair] = s * inv_count;
}

}  // namespace

// kp_all / qp_all: fp16 [n_frames][hw_pad][128]; bsq_all / ms_all: fp32 [n_frames][hw_pad] (zero in the padding);
// pair_a / pair_b: device int32 [n_pairs] frame indices (A = already chosen frame, B = candidate);
// partial: fp32 workspace [n_pairs][(hw_pad/128)^2]; scores: fp32 [n_pairs].
extern "C" int xm_pair_dissimilarity(const void* kp_all, const void* qp_all, const float* bsq_all, const float* ms_all, int32_t n_frames,
                                     int32_t hw, int32_t hw_pad, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs,
                                     float* partial, float* scores, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    XM_REQUIRE(kp_all && qp_all && bsq_all && ms_all && pair_a && pair_b && partial && scores, "xm_pair_dissimilarity: null pointer");
    XM_REQUIRE(n_frames >= 1 && hw >= 1 && hw_pad >= hw && hw_pad % PT == 0, "xm_pair_dissimilarity: hw_pad must be hw rounded up to 128");
    XM_REQUIRE(n_pairs >= 1 && n_pairs <= 65535, "xm_pair_dissimilarity: 1..65535 pairs per call");
    tc5_debug_init();
    PairMaps maps;
    {
        uint64_t d[3] = {PKP, (uint64_t)hw_pad, (uint64_t)n_frames};
        uint64_t st[2] = {PKP * 2, (uint64_t)hw_pad * PKP * 2};
        uint32_t b[3] = {64, PT, 1};
        if (xm_make_tmap_f16(&maps.kp, kp_all, 3, d, st, b)) return XM_ERR_CUDA;
        if (xm_make_tmap_f16(&maps.qp, qp_all, 3, d, st, b)) return XM_ERR_CUDA;
    }
    static XmPerDevice attr_token = {0};
    const int smem = (int)sizeof(PairSmem) + 1024;
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(pair_dissim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int tiles = hw_pad / PT;
    XM_CHECK_CUDA(tc5_launch(pair_dissim_kernel, dim3(tiles, tiles, n_pairs), dim3(192), smem, stream, maps, bsq_all, ms_all, (int)hw_pad,
                             (const int*)pair_a, (const int*)pair_b, partial));
    XM_CHECK_CUDA(tc5_launch(pair_reduce_kernel, dim3((n_pairs + 7) / 8), dim3(256), 0, stream, (const float*)partial, tiles * tiles, (int)n_pairs,
                             1.0f / ((float)hw * (float)hw), scores));
    xm_count_launches(2);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}
