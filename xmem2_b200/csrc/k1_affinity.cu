// k1_affinity.cu — fused memory read for XMem++:  similarity -> top-k softmax -> value readout (+usage)
//
// Replaces the reference sequence get_similarity (model/memory_util.py:7-39) -> do_softmax top-k branch
// (:41-54, no max-subtraction) -> usage = affinity.sum (:62-63) -> v @ affinity (memory_manager.py:57-59)
// over the concatenated long-term | working | permanent banks (memory_manager.py:82-128,143-182),
// without ever materialising the N x HW similarity / affinity matrices.
//
// Math.  With packed operands  Kp[n] = (k_n^2 , k_n)  and  Qp[q] = (-e_q , 2 k_q e_q)  (fp16, 128 wide)
//     S'[q,n] = Qp[q] . Kp[n] = -sum_c e k_n^2 + 2 sum_c k_n k_q e          (tcgen05.mma, fp32 accumulate)
//     S[q,n]  = (S'[q,n] - bsq[q]) * shrinkage[n] / sqrt(64)                 (bsq = sum_c e k_q^2, fp32)
// Pass 1 (k1_topk_pass1): every CTA owns 128 queries x one slice of memory columns, streams 64-column
//     key tiles through TMA -> smem -> tcgen05 -> TMEM, and each of 128 threads keeps the running top-k of
//     ITS query (one TMEM lane = one query).  k1_topk_merge combines the slices: tau[q] = k-th largest
//     S, inv_den[q] = 1 / sum_topk exp(S).
// Pass 2 (k1_readout_pass2): recomputes the same S tiles (bit-identical: same instruction stream), forms
//     P = (S >= tau) ? exp(S) * inv_den : 0 as fp16 in shared memory, and accumulates
//     O^T[c, q] += V[c, n-tile] . P[q, n-tile]^T  in TMEM (dense tensor-core contraction, as the
//     reference's dense v @ affinity).  Split over column slices; k1_finish sums the slices.
#include <cfloat>
#include <cmath>
#include <cstring>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

namespace {

constexpr int TQ = 128;          // queries per CTA (UMMA M / TMEM lanes)
constexpr int TN = 64;           // memory columns per tile
constexpr int KP = 128;          // packed key width
constexpr int LISTK = XM_MAX_TOPK;
constexpr int P1_STAGES = 4;
constexpr int P1_SBUF = 4;
constexpr int P2_KSTAGES = 3;
constexpr int P2_VSTAGES = 3;
constexpr int P2_SBUF = 2;
constexpr int P2_PBUF = 2;
constexpr int CHALF = 256;       // value channels per pass-2 CTA
constexpr float LOG2E = 1.4426950408889634f;

struct alignas(64) K1Maps {
    CUtensorMap q;       // [128, hw_pad]            box [64,128]
    CUtensorMap k[3];    // [128, cap_b]             box [64, 64]
    CUtensorMap v[3];    // [cap_b, 512, n_obj_cap]  box [64,128,1]
};

struct K1Seg {
    int nseg;
    int bank[3];
    int begin[3];
    int end[3];
    int origin[3];       // begin rounded down to 8 columns: TMA needs 16-byte aligned starts on the contiguous (column) axis of V
    int tile0[4];        // prefix sum of tiles per segment
    int col0[3];         // first column of the segment inside the group's concatenated column index
    const float* shr[3];
    float* usage[3];
};

// tile t -> segment s, first column `col` of the 64-wide tile, valid lanes [lo, hi) inside the tile
__device__ __forceinline__ void locate_tile(const K1Seg& sg, int t, int& s, int& col, int& lo, int& hi) {
    s = 0;
    if (sg.nseg > 1 && t >= sg.tile0[1]) s = 1;
    if (sg.nseg > 2 && t >= sg.tile0[2]) s = 2;
    col = sg.origin[s] + (t - sg.tile0[s]) * TN;
    lo = max(0, sg.begin[s] - col);
    hi = min(TN, sg.end[s] - col);
}

__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * LOG2E));
    return y;
}

// ---------------------------------------------------------------------------------------------
// operand packing
// ---------------------------------------------------------------------------------------------
__global__ void query_pack_kernel(const __half* __restrict__ key, const __half* __restrict__ sel, int hw, int hw_pad,
                                  __half* __restrict__ qp, float* __restrict__ bsq) {
    pdl_wait();
    pdl_launch_dependents();
    // one warp per query row; lane handles channels lane, lane+32
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= hw_pad) return;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int c = lane + 32 * h;
        __half k = __float2half(0.f), e = __float2half(0.f);
        if (row < hw) {
            k = key[(size_t)row * XM_CK + c];
            e = sel[(size_t)row * XM_CK + c];
        }
        __half ke = __hmul(k, e);                        // fp16 product, as qk*qe under autocast
        qp[(size_t)row * KP + c] = __hneg(e);
        qp[(size_t)row * KP + XM_CK + c] = __hadd(ke, ke);
        float kf = __half2float(k);
        acc += __half2float(e) * (kf * kf);              // fp32, memory_util.py:26
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) bsq[row] = acc;
}

__global__ void key_pack_kernel(const __half* __restrict__ key, int n, __half* __restrict__ dst) {
    pdl_wait();
    pdl_launch_dependents();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * XM_CK) return;
    int row = i / XM_CK, c = i % XM_CK;
    __half k = key[i];
    float kf = __half2float(k);
    dst[(size_t)row * KP + c] = __float2half_rn(kf * kf);    // mk.pow(2) in fp32 then fp16 for the GEMM
    dst[(size_t)row * KP + XM_CK + c] = k;
}

// ---------------------------------------------------------------------------------------------
// scan kernel (two modes) : streams S tiles and keeps, per query, a 32-entry summary of each column slice
//   MODE_SLOTMAX : slot[j] = max over the slice's tiles of S[q, 64*t + j]  -> 32 running maxima of disjoint
//                  column subsets, branch-free (one FMNMX per score).  The k-th largest of all slot maxima is a
//                  LOWER bound tau_lo of the true k-th largest score (they are distinct memory columns) and in
//                  practice within a few ranks of it.
//   MODE_COLLECT : keeps the 32 largest scores > pred(tau_lo) of the slice (append until full, then
//                  replace-min) -> exact, and 99.7 % of the scores fail the first compare.
// Threads: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = 8 selection warps; the warp pair
// (w, w+4) shares a TMEM lane quadrant and splits the 64 tile columns in halves.
// ---------------------------------------------------------------------------------------------
constexpr int MODE_SLOTMAX = 0;
constexpr int MODE_COLLECT = 1;
constexpr int SCAN_THREADS = 64 + 256;

struct ScanSmem {
    alignas(1024) uint8_t q[2][TQ * 128];                 // 2 K-halves x (128 rows x 128 B)
    alignas(1024) uint8_t k[P1_STAGES][2][TN * 128];      // per stage 2 K-halves x (64 rows x 128 B)
    float list[2][2 * LISTK][TQ];                         // MODE_COLLECT: per (column half, query) append list (2x capacity)
    alignas(8) uint64_t qfull;
    uint64_t kfull[P1_STAGES], kempty[P1_STAGES];
    uint64_t sfull[P1_SBUF], sempty[P1_SBUF];
    uint32_t tmem_base;
};

__device__ __forceinline__ float score2(uint32_t acc_bits, float bsq8, float ms) {
    // ((S' - b_sq) * shrinkage) / 8 == (S'/8 - b_sq/8) * shrinkage exactly (power-of-two scaling commutes with rounding)
    return fmaf(__uint_as_float(acc_bits), 0.125f, -bsq8) * ms;
}

template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
k1_scan(const __grid_constant__ K1Maps maps, const K1Seg* __restrict__ sgp, const float* __restrict__ bsq, const float* __restrict__ tau_lo,
        int hw_pad, float* __restrict__ cand, float* __restrict__ dbg_scores) {
    extern __shared__ uint8_t smem_raw[];
    ScanSmem& sm = *reinterpret_cast<ScanSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ K1Seg sg;                       // column-range table lives in device memory (CUDA-graph friendly)
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x < sizeof(K1Seg) / 4) reinterpret_cast<uint32_t*>(&sg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(sgp)[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, split = blockIdx.y;
    const int total_tiles = sg.tile0[sg.nseg];
    const int tiles_per_split = (total_tiles + gridDim.y - 1) / gridDim.y;
    const int t_begin = split * tiles_per_split;
    const int t_end = min(total_tiles, t_begin + tiles_per_split);
    const int nt = max(0, t_end - t_begin);

    if (threadIdx.x == 0) {
        mbar_init(&sm.qfull, 1);
        for (int i = 0; i < P1_STAGES; ++i) { mbar_init(&sm.kfull[i], 1); mbar_init(&sm.kempty[i], 1); }
        for (int i = 0; i < P1_SBUF; ++i) { mbar_init(&sm.sfull[i], 1); mbar_init(&sm.sempty[i], 256); }
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.q);
            mbar_expect_tx(&sm.qfull, 2 * TQ * 128);
            tma_load_2d(sm.q[0], &maps.q, &sm.qfull, 0, qtile * TQ);
            tma_load_2d(sm.q[1], &maps.q, &sm.qfull, 64, qtile * TQ);
            for (int i = 0; i < nt; ++i) {
                int s, col, lo, nv;
                locate_tile(sg, t_begin + i, s, col, lo, nv);
                const int st = i % P1_STAGES, ph = (i / P1_STAGES) & 1;
                mbar_wait(&sm.kempty[st], ph ^ 1, 2);
                mbar_expect_tx(&sm.kfull[st], 2 * TN * 128);
                const CUtensorMap* km = &maps.k[sg.bank[s]];
                tma_load_2d(sm.k[st][0], km, &sm.kfull[st], 0, col);
                tma_load_2d(sm.k[st][1], km, &sm.kfull[st], 64, col);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(TQ, TN);
            mbar_wait(&sm.qfull, 0, 1);
            for (int i = 0; i < nt; ++i) {
                const int st = i % P1_STAGES, ph = (i / P1_STAGES) & 1;
                const int sb = i % P1_SBUF, sph = (i / P1_SBUF) & 1;
                mbar_wait(&sm.kfull[st], ph, 3);
                mbar_wait(&sm.sempty[sb], sph ^ 1, 4);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint64_t a = make_desc_sw128(smem_u32(sm.q[h]) + j * 32);
                        uint64_t b = make_desc_sw128(smem_u32(sm.k[st][h]) + j * 32);
                        mma_f16_ss(tmem + sb * TN, a, b, idesc, (h | j) ? 1u : 0u);
                    }
                mma_commit(&sm.kempty[st]);
                mma_commit(&sm.sfull[sb]);
            }
        }
    } else {
        const int lane_base = (warp & 3) * 32;
        const int half = (warp - 2) >> 2;               // which 32 of the tile's 64 columns
        const int row = lane_base + lane;
        const int q = qtile * TQ + row;
        const float bsq8 = bsq[q] * 0.125f;
        float slot[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) slot[j] = -INFINITY;
        // MODE_COLLECT state: append-only list of capacity 2*LISTK, compacted to the LISTK largest when it could overflow
        float thr = -INFINITY;
        int count = 0;
        float* mylist = &sm.list[half][0][row];          // element u at mylist[u * TQ]
        if (MODE == MODE_COLLECT) {
            const float t = tau_lo[q];
            thr = (t == -INFINITY) ? -INFINITY : ((t == INFINITY) ? FLT_MAX : __uint_as_float(
                      t > 0.f ? __float_as_uint(t) - 1u : (t < 0.f ? __float_as_uint(t) + 1u : 0x80000001u)));   // pred(tau_lo)
        }
        auto compact = [&]() {      // keep the LISTK largest of `count` entries; thr = the smallest kept (rare path)
            for (int keep = 0; keep < LISTK; ++keep) {
                float m = mylist[keep * TQ]; int p = keep;
                for (int u = keep + 1; u < count; ++u) { const float v = mylist[u * TQ]; if (v > m) { m = v; p = u; } }
                const float t0 = mylist[keep * TQ]; mylist[keep * TQ] = m; mylist[p * TQ] = t0;
            }
            count = LISTK;
            thr = mylist[(LISTK - 1) * TQ];
        };
        // software prefetch of the shrinkage values of the next tile (hides the L2 latency behind this tile's work)
        float ms_next[32];
        auto load_ms = [&](int i, float (&ms)[32]) {
            int s, col, lo, nv;
            locate_tile(sg, t_begin + i, s, col, lo, nv);
            const int c0 = col + half * 32;
            const int jlo = lo - half * 32, jhi = nv - half * 32;
            const float* shr = sg.shr[s] + c0;
            if (jlo <= 0 && jhi >= 32) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(shr + j));   // c0 is a multiple of 8
                    ms[j] = f.x; ms[j + 1] = f.y; ms[j + 2] = f.z; ms[j + 3] = f.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) ms[j] = (j >= jlo && j < jhi) ? __ldg(shr + j) : 1.f;
            }
        };
        if (nt > 0) load_ms(0, ms_next);
        for (int i = 0; i < nt; ++i) {
            int s, col, lo, nv;
            locate_tile(sg, t_begin + i, s, col, lo, nv);
            float ms[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) ms[j] = ms_next[j];
            if (i + 1 < nt) load_ms(i + 1, ms_next);
            const int sb = i % P1_SBUF, sph = (i / P1_SBUF) & 1;
            mbar_wait(&sm.sfull[sb], sph, 5);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + sb * TN + half * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sm.sempty[sb]);
            const int c0 = col + half * 32;                  // first column of this thread's 32
            const int jlo = lo - half * 32, jhi = nv - half * 32;   // valid j in [jlo, jhi)
            const bool full_tile = (jlo <= 0) && (jhi >= 32);
            float sc[32];
            if (full_tile) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sc[j] = score2(r[j], bsq8, ms[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) sc[j] = (j >= jlo && j < jhi) ? score2(r[j], bsq8, ms[j]) : -INFINITY;
            }
            if (MODE == MODE_SLOTMAX) {
#pragma unroll
                for (int j = 0; j < 32; ++j) slot[j] = fmaxf(slot[j], sc[j]);
                if (dbg_scores) {        // tests only (warp-uniform)
                    float* dbg = dbg_scores + ((ptrdiff_t)sg.col0[s] + (c0 - sg.begin[s])) * (ptrdiff_t)hw_pad + q;
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (sc[j] != -INFINITY) dbg[(ptrdiff_t)j * hw_pad] = sc[j];
                }
            } else {
                if (count > LISTK) compact();                 // guarantees room for 32 appends below
                float* wp = mylist + count * TQ;
#pragma unroll
                for (int j = 0; j < 32; ++j) {                // branch-free predicated append
                    const bool take = sc[j] > thr;
                    if (take) *wp = sc[j];
                    wp += take ? TQ : 0;
                }
                count = static_cast<int>(wp - mylist) / TQ;
            }
        }
        float* dst = cand + ((size_t)(split * 2 + half) * hw_pad + q) * LISTK;
        if (MODE == MODE_SLOTMAX) {
#pragma unroll
            for (int u = 0; u < LISTK; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(slot[u], slot[u + 1], slot[u + 2], slot[u + 3]);
        } else {
            if (count > LISTK) compact();
            for (int u = 0; u < LISTK; ++u) dst[u] = (u < count) ? mylist[u * TQ] : -INFINITY;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// k-th largest over the per-slice 32-entry lists: one warp per query, lane l owns list l (nlists <= 32).
// Exact selection by bisection on the order-preserving integer image of fp32 (32 branch-free count rounds) instead
// of k rounds of max extraction.  want_den: also 1 / sum_topk exp(S) (do_softmax top-k branch, memory_util.py:48-49:
// no max subtraction); entries tied with the k-th value all count, consistently with pass 2's (S >= tau).
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__global__ void k1_topk_merge(const float* __restrict__ cand, int nlists, int hw, int hw_pad, int top_k, int want_den,
                              float* __restrict__ tau, float* __restrict__ inv_den) {
    pdl_wait();
    pdl_launch_dependents();
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= hw_pad) return;
    if (q >= hw) {                       // padded query rows never select anything
        if (lane == 0) { tau[q] = INFINITY; if (want_den) inv_den[q] = 0.f; }
        return;
    }
    uint32_t v[LISTK];
#pragma unroll
    for (int u = 0; u < LISTK; ++u) v[u] = f2ord(-INFINITY);
    if (lane < nlists) {
        const float4* src = reinterpret_cast<const float4*>(cand + ((size_t)lane * hw_pad + q) * LISTK);
#pragma unroll
        for (int u = 0; u < LISTK / 4; ++u) {
            const float4 f = src[u];
            v[4 * u] = f2ord(f.x); v[4 * u + 1] = f2ord(f.y); v[4 * u + 2] = f2ord(f.z); v[4 * u + 3] = f2ord(f.w);
        }
    }
    // largest t with count(v >= t) >= top_k  ==  the top_k-th largest value (if fewer finite entries: -inf's image)
    uint32_t t = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int u = 0; u < LISTK; ++u) c += (v[u] >= trial) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (c >= top_k) t = trial;
    }
    const float kth = ord2f(t);
    if (want_den) {
        float den = 0.f;
#pragma unroll
        for (int u = 0; u < LISTK; ++u) den += (v[u] >= t) ? fast_exp(ord2f(v[u])) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        if (lane == 0) inv_den[q] = 1.f / den;
    }
    if (lane == 0) tau[q] = kth;
}

// T-sharded mode: compress the per-slice candidate lists of THIS rank to its 32 largest scores per query, the
// record that is all-gathered across ranks (a rank's 32 largest necessarily contain its share of the global top-k).
__global__ void k1_export_top32(const float* __restrict__ cand, int nlists, int hw_pad, float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= hw_pad) return;
    uint32_t v[LISTK];
#pragma unroll
    for (int u = 0; u < LISTK; ++u) v[u] = f2ord(-INFINITY);
    if (lane < nlists) {
        const float4* src = reinterpret_cast<const float4*>(cand + ((size_t)lane * hw_pad + q) * LISTK);
#pragma unroll
        for (int u = 0; u < LISTK / 4; ++u) {
            const float4 f = src[u];
            v[4 * u] = f2ord(f.x); v[4 * u + 1] = f2ord(f.y); v[4 * u + 2] = f2ord(f.z); v[4 * u + 3] = f2ord(f.w);
        }
    }
    uint32_t t = 0u;                                   // image of the 32nd largest local candidate
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int u = 0; u < LISTK; ++u) c += (v[u] >= trial) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (c >= LISTK) t = trial;
    }
    // strictly-greater entries first, then ties with t until the record is full
    float* dst = out + (size_t)q * LISTK;
    int mine_gt = 0, mine_eq = 0;
#pragma unroll
    for (int u = 0; u < LISTK; ++u) { mine_gt += (v[u] > t) ? 1 : 0; mine_eq += (v[u] == t) ? 1 : 0; }
    int pre_gt = mine_gt, pre_eq = mine_eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, pre_gt, o), b2 = __shfl_up_sync(0xffffffffu, pre_eq, o);
        if (lane >= o) { pre_gt += a; pre_eq += b2; }
    }
    const int total_gt = __shfl_sync(0xffffffffu, pre_gt, 31);
    int pos_gt = pre_gt - mine_gt, pos_eq = total_gt + pre_eq - mine_eq;
#pragma unroll
    for (int u = 0; u < LISTK; ++u) {
        if (v[u] > t) { dst[pos_gt++] = ord2f(v[u]); }
        else if (v[u] == t) { if (pos_eq < LISTK) dst[pos_eq] = ord2f(v[u]); ++pos_eq; }
    }
    const int filled = min(LISTK, total_gt + __shfl_sync(0xffffffffu, pre_eq, 31));
    for (int u = filled + lane; u < LISTK; u += 32) dst[u] = -INFINITY;
}

// ---------------------------------------------------------------------------------------------
// pass 2: P = (S >= tau) ? exp(S) / den : 0 ;  O^T[c,q] += V[c,n] P[q,n]
// ---------------------------------------------------------------------------------------------
struct P2Smem {
    alignas(1024) uint8_t q[2][TQ * 128];
    alignas(1024) uint8_t k[P2_KSTAGES][2][TN * 128];
    alignas(1024) uint8_t v[P2_VSTAGES][2][128 * 128];    // 2 M-chunks x (128 channel rows x 64 columns)
    alignas(1024) uint8_t p[P2_PBUF][TQ * 128];           // 128 query rows x 64 columns fp16
    alignas(8) uint64_t qfull;
    uint64_t kfull[P2_KSTAGES], kempty[P2_KSTAGES];
    uint64_t vfull[P2_VSTAGES], vempty[P2_VSTAGES];
    uint64_t sfull[P2_SBUF], sempty[P2_SBUF];
    uint64_t pfull[P2_PBUF], pempty[P2_PBUF];
    uint64_t ofull;
    uint32_t tmem_base;
};

constexpr int P2_THREADS = 64 + 512;     // warp 0 TMA, warp 1 MMA, 16 softmax warps: 4 per TMEM lane quadrant, 16 columns each

__global__ void __launch_bounds__(P2_THREADS, 1)
k1_readout_pass2(const __grid_constant__ K1Maps maps, const K1Seg* __restrict__ sgp, const float* __restrict__ bsq,
                 const float* __restrict__ tau, const float* __restrict__ inv_den, int hw_pad, int obj_begin,
                 int n_obj, int do_usage, float* __restrict__ partial) {
    extern __shared__ uint8_t smem_raw[];
    P2Smem& sm = *reinterpret_cast<P2Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ K1Seg sg;
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x < sizeof(K1Seg) / 4) reinterpret_cast<uint32_t*>(&sg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(sgp)[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x;
    const int obj = blockIdx.y >> 1, chalf = blockIdx.y & 1;       // local object index within the group
    const int split = blockIdx.z;
    const int total_tiles = sg.tile0[sg.nseg];
    const int tiles_per_split = (total_tiles + gridDim.z - 1) / gridDim.z;
    const int t_begin = split * tiles_per_split;
    const int t_end = min(total_tiles, t_begin + tiles_per_split);
    const int nt = max(0, t_end - t_begin);

    if (threadIdx.x == 0) {
        mbar_init(&sm.qfull, 1);
        for (int i = 0; i < P2_KSTAGES; ++i) { mbar_init(&sm.kfull[i], 1); mbar_init(&sm.kempty[i], 1); }
        for (int i = 0; i < P2_VSTAGES; ++i) { mbar_init(&sm.vfull[i], 1); mbar_init(&sm.vempty[i], 1); }
        for (int i = 0; i < P2_SBUF; ++i) { mbar_init(&sm.sfull[i], 1); mbar_init(&sm.sempty[i], 512); }
        for (int i = 0; i < P2_PBUF; ++i) { mbar_init(&sm.pfull[i], 512); mbar_init(&sm.pempty[i], 1); }
        mbar_init(&sm.ofull, 1);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t tmem_s = tmem + 256;        // S buffers after the two 128-column O^T chunks

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&sm.qfull, 2 * TQ * 128);
            tma_load_2d(sm.q[0], &maps.q, &sm.qfull, 0, qtile * TQ);
            tma_load_2d(sm.q[1], &maps.q, &sm.qfull, 64, qtile * TQ);
            for (int i = 0; i < nt; ++i) {
                int s, col, lo, nv;
                locate_tile(sg, t_begin + i, s, col, lo, nv);
                const int b = sg.bank[s];
                {
                    const int st = i % P2_KSTAGES, ph = (i / P2_KSTAGES) & 1;
                    mbar_wait(&sm.kempty[st], ph ^ 1, 2);
                    mbar_expect_tx(&sm.kfull[st], 2 * TN * 128);
                    tma_load_2d(sm.k[st][0], &maps.k[b], &sm.kfull[st], 0, col);
                    tma_load_2d(sm.k[st][1], &maps.k[b], &sm.kfull[st], 64, col);
                }
                {
                    const int st = i % P2_VSTAGES, ph = (i / P2_VSTAGES) & 1;
                    mbar_wait(&sm.vempty[st], ph ^ 1, 6);
                    mbar_expect_tx(&sm.vfull[st], 2 * 128 * 128);
                    tma_load_3d(sm.v[st][0], &maps.v[b], &sm.vfull[st], col, chalf * CHALF, obj_begin + obj);
                    tma_load_3d(sm.v[st][1], &maps.v[b], &sm.vfull[st], col, chalf * CHALF + 128, obj_begin + obj);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(TQ, TN);
            constexpr uint32_t idesc_o = make_idesc_f16(128, TQ);
            mbar_wait(&sm.qfull, 0, 1);
            for (int i = 0; i <= nt; ++i) {
                if (i < nt) {          // S(i) = Qp . Kp^T
                    const int st = i % P2_KSTAGES, ph = (i / P2_KSTAGES) & 1;
                    const int sb = i % P2_SBUF, sph = (i / P2_SBUF) & 1;
                    mbar_wait(&sm.kfull[st], ph, 3);
                    mbar_wait(&sm.sempty[sb], sph ^ 1, 4);
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint64_t a = make_desc_sw128(smem_u32(sm.q[h]) + j * 32);
                            uint64_t b = make_desc_sw128(smem_u32(sm.k[st][h]) + j * 32);
                            mma_f16_ss(tmem_s + sb * TN, a, b, idesc_s, (h | j) ? 1u : 0u);
                        }
                    mma_commit(&sm.kempty[st]);
                    mma_commit(&sm.sfull[sb]);
                }
                if (i > 0) {           // O^T += V(i-1) . P(i-1)^T
                    const int u = i - 1;
                    const int vs = u % P2_VSTAGES, vph = (u / P2_VSTAGES) & 1;
                    const int pb = u % P2_PBUF, pph = (u / P2_PBUF) & 1;
                    mbar_wait(&sm.vfull[vs], vph, 7);
                    mbar_wait(&sm.pfull[pb], pph, 8);
                    tc_fence_after();
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint64_t a = make_desc_sw128(smem_u32(sm.v[vs][m]) + j * 32);
                            uint64_t b = make_desc_sw128(smem_u32(sm.p[pb]) + j * 32);
                            mma_f16_ss(tmem + m * 128, a, b, idesc_o, (u > 0 || j > 0) ? 1u : 0u);
                        }
                    mma_commit(&sm.vempty[vs]);
                    mma_commit(&sm.pempty[pb]);
                }
            }
            mma_commit(&sm.ofull);
        }
    } else {
        const int lane_base = (warp & 3) * 32;
        const int quarter = (warp - 2) >> 2;           // which 16 of the tile's 64 columns
        const int row = lane_base + lane;
        const int q = qtile * TQ + row;
        const float bsq8 = bsq[q] * 0.125f;
        const float my_tau = tau[q];
        const float my_inv = inv_den[q];
        const bool usage_cta = do_usage && blockIdx.y == 0;
        // The affinity tile is ~99.9 % zeros (k of N columns per query).  The two P buffers are zeroed once; per tile a
        // thread only evaluates the threshold test, writes its few non-zero entries and remembers them (bit mask per
        // buffer) so that it can clear them again when the buffer comes round.
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int pb = 0; pb < P2_PBUF; ++pb)
#pragma unroll
                for (int c = 0; c < 2; ++c) *reinterpret_cast<uint4*>(sm.p[pb] + row * 128 + (((quarter * 2 + c) ^ (row & 7)) << 4)) = z;
            fence_proxy_async_smem();
        }
        uint32_t written0 = 0u, written1 = 0u;
        auto p_addr = [&](int pb, int j) -> __half* {          // element (row, column quarter*16 + j) of the swizzled P tile
            const int n = quarter * 16 + j;
            return reinterpret_cast<__half*>(sm.p[pb] + row * 128 + (((n >> 3) ^ (row & 7)) << 4) + (n & 7) * 2);
        };
        auto load_ms = [&](int i, float (&ms)[16]) {
            int s, col, lo, nv;
            locate_tile(sg, t_begin + i, s, col, lo, nv);
            const int c0 = col + quarter * 16;
            const int jlo = lo - quarter * 16, jhi = nv - quarter * 16;
            const float* shr = sg.shr[s] + c0;
            if (jlo <= 0 && jhi >= 16) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(shr + j));
                    ms[j] = f.x; ms[j + 1] = f.y; ms[j + 2] = f.z; ms[j + 3] = f.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) ms[j] = (j >= jlo && j < jhi) ? __ldg(shr + j) : 1.f;
            }
        };
        auto process = [&](int i, const float (&ms)[16]) {
            int s, col, lo, nv;
            locate_tile(sg, t_begin + i, s, col, lo, nv);
            const int sb = i % P2_SBUF, sph = (i / P2_SBUF) & 1;
            const int pb = i % P2_PBUF, pph = (i / P2_PBUF) & 1;
            mbar_wait(&sm.sfull[sb], sph, 5);
            tc_fence_after();
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_s + (static_cast<uint32_t>(lane_base) << 16) + sb * TN + quarter * 16, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sm.sempty[sb]);
            const int c0 = col + quarter * 16;
            const int jlo = lo - quarter * 16, jhi = nv - quarter * 16;
            const bool full_tile = (jlo <= 0) && (jhi >= 16);
            float* usage = (usage_cta && sg.usage[s]) ? sg.usage[s] + c0 : nullptr;
            uint32_t h0 = 0u, h1 = 0u, h2 = 0u, h3 = 0u;          // four independent accumulation chains (ILP)
            if (full_tile) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h0 |= (score2(r[j], bsq8, ms[j]) >= my_tau) ? (1u << j) : 0u;
                    h1 |= (score2(r[4 + j], bsq8, ms[4 + j]) >= my_tau) ? (16u << j) : 0u;
                    h2 |= (score2(r[8 + j], bsq8, ms[8 + j]) >= my_tau) ? (256u << j) : 0u;
                    h3 |= (score2(r[12 + j], bsq8, ms[12 + j]) >= my_tau) ? (4096u << j) : 0u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) h0 |= ((j >= jlo && j < jhi) && score2(r[j], bsq8, ms[j]) >= my_tau) ? (1u << j) : 0u;
            }
            const uint32_t hit = h0 | h1 | h2 | h3;              // bit j: S[q, c0 + j] >= tau  (k of N columns)
            mbar_wait(&sm.pempty[pb], pph ^ 1, 9);
            const uint32_t old = pb ? written1 : written0;
            if (old | hit) {                                      // rare: this thread owns non-zero affinity entries
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if ((old >> j) & 1u) *p_addr(pb, j) = __float2half(0.f);
                    if ((hit >> j) & 1u) {
                        const float pv = fast_exp(score2(r[j], bsq8, ms[j])) * my_inv;
                        *p_addr(pb, j) = __float2half_rn(pv);
                        if (usage) atomicAdd(usage + j, pv);
                    }
                }
            }
            if (pb) written1 = hit; else written0 = hit;
            fence_proxy_async_smem();
            mbar_arrive(&sm.pfull[pb]);
        };
        static_assert(P2_PBUF == 2, "the non-zero bookkeeping assumes two P buffers");
        // two tiles per iteration with ping-pong shrinkage registers: the loads for tile i+1 are in flight while
        // tile i is processed (no register copies)
        float msA[16], msB[16];
        if (nt > 0) load_ms(0, msA);
        for (int i = 0; i < nt; i += 2) {
            if (i + 1 < nt) load_ms(i + 1, msB);
            process(i, msA);
            if (i + 1 < nt) {
                if (i + 2 < nt) load_ms(i + 2, msA);
                process(i + 1, msB);
            }
        }
        // epilogue: the 4 warps of a lane quadrant drain O^T chunk m = quarter>>1, query columns (quarter&1)*64..+63
        mbar_wait(&sm.ofull, 0, 10);
        tc_fence_after();
        const int n_obj_all = gridDim.y >> 1;
        (void)n_obj;
        {
            const int m = quarter >> 1, qh = quarter & 1;
            const int c = chalf * CHALF + m * 128 + row;
            float* dst = partial + (((size_t)split * n_obj_all + obj) * XM_CV + c) * hw_pad + qtile * TQ + qh * 64;
#pragma unroll 1
            for (int qq = 0; qq < 2; ++qq) {
                uint32_t r[32];
                if (nt > 0) {
                    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(lane_base) << 16) + m * 128 + qh * 64 + qq * 32, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(dst + qq * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// sum the column slices, convert to fp16, write CHW and/or HWC
__global__ void k1_finish(const float* __restrict__ partial, int nsplit, int n_obj, int hw, int hw_pad, int obj_begin,
                          __half* __restrict__ out_chw, __half* __restrict__ out_hwc) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float tile[32][33];
    const int o = blockIdx.z;
    const int c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, q = q0 + tx;
        float acc = 0.f;
        if (q < hw)
            for (int s = 0; s < nsplit; ++s) acc += partial[(((size_t)s * n_obj + o) * XM_CV + c) * hw_pad + q];
        tile[r][tx] = acc;
        if (out_chw && q < hw) out_chw[((size_t)(obj_begin + o) * XM_CV + c) * hw + q] = __float2half_rn(acc);
    }
    if (out_hwc) {
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int q = q0 + r, c = c0 + tx;
            if (q < hw) out_hwc[((size_t)(obj_begin + o) * hw + q) * XM_CV + c] = __float2half_rn(tile[tx][r]);
        }
    }
}

// T-sharded mode: sum the local column slices in fp32 [n_obj][512][hw_pad] (all-reduced across ranks afterwards)
__global__ void k1_sum_splits(const float* __restrict__ partial, int nsplit, size_t plane, float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < nsplit; ++s) acc += partial[(size_t)s * plane + i];
        out[i] = acc;
    }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int xm_query_pack(const void* key_hwc, const void* sel_hwc, int32_t hw, int32_t hw_pad, void* qp, float* bsq,
                             void* stream) {
    XM_REQUIRE(key_hwc && sel_hwc && qp && bsq, "xm_query_pack: null pointer");
    XM_REQUIRE(hw > 0 && hw_pad >= hw && hw_pad % TQ == 0, "xm_query_pack: hw_pad must be a multiple of 128 and >= hw");
    const int warps = 8;
    XM_CHECK_CUDA(tc5_launch(query_pack_kernel, dim3((hw_pad + warps - 1) / warps), dim3(warps * 32), 0, (cudaStream_t)stream,
                             (const __half*)key_hwc, (const __half*)sel_hwc, hw, hw_pad, (__half*)qp, bsq));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_key_pack(const void* key_hwc, int32_t n, void* dst_rows, void* stream) {
    XM_REQUIRE(key_hwc && dst_rows && n >= 0, "xm_key_pack: bad arguments");
    if (n == 0) return XM_OK;
    const int total = n * XM_CK;
    XM_CHECK_CUDA(tc5_launch(key_pack_kernel, dim3((total + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const __half*)key_hwc, n, (__half*)dst_rows));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

static const int K1_MAX_SPLIT = 32;
static const int K1_PLAN_BYTES = 4096;          // XM_MAX_GROUPS segment tables at the head of the workspace

extern "C" int64_t xm_affinity_workspace_bytes(int32_t hw, int32_t n_obj_total) {
    const int64_t hw_pad = (hw + TQ - 1) / TQ * TQ;
    int64_t b = K1_PLAN_BYTES;
    b += align_up((size_t)K1_MAX_SPLIT * hw_pad * LISTK * 4, 256);        // candidates
    b += 3 * align_up((size_t)hw_pad * 4, 256);                           // tau_lo, tau, inv_den
    b += align_up((size_t)K1_MAX_SPLIT * n_obj_total * XM_CV * hw_pad * 4, 256);   // partial readouts
    return b;
}

static_assert(sizeof(K1Seg) * XM_MAX_GROUPS <= K1_PLAN_BYTES, "plan area too small");
static_assert(sizeof(K1Seg) % 4 == 0 && sizeof(K1Seg) / 4 <= SCAN_THREADS, "K1Seg is copied by one thread per word");

// Build the per-group column-range tables (host side).  plan_out receives XM_MAX_GROUPS K1Seg records.
static int k1_build_plan(const xm_affinity_args_t* a, K1Seg* plan, bool allow_small = false) {
    XM_REQUIRE(a->n_groups > 0 && a->n_groups <= XM_MAX_GROUPS, "xm_affinity: bad n_groups %d", a->n_groups);
    for (int g = 0; g < a->n_groups; ++g) {
        const xm_group_t& gr = a->groups[g];
        XM_REQUIRE(gr.n_obj > 0 && gr.obj_begin >= 0 && gr.obj_begin + gr.n_obj <= a->n_obj_total, "xm_affinity: bad group %d objects", g);
        K1Seg& sg = plan[g];
        sg.nseg = 0;
        int tiles = 0, cols = 0;
        for (int i = 0; i < 3; ++i) {
            const xm_bank_t& bk = a->banks[i];
            sg.shr[i] = nullptr; sg.usage[i] = nullptr; sg.bank[i] = 0; sg.begin[i] = 0; sg.end[i] = 0; sg.col0[i] = 0; sg.origin[i] = 0;
            if (bk.size <= 0 || !bk.keys) continue;
            XM_REQUIRE(bk.cap % 8 == 0 && bk.size <= bk.cap, "xm_affinity: bank %d cap must be a multiple of 8 and >= size", i);
            XM_REQUIRE(bk.shrinkage && bk.values && bk.n_obj_cap > 0, "xm_affinity: bank %d has null shrinkage/values", i);
            const int begin = gr.begin[i];
            XM_REQUIRE(begin >= 0 && begin <= bk.size, "xm_affinity: group %d bank %d begin %d outside [0,%d]", g, i, begin, bk.size);
            if (begin == bk.size) continue;
            XM_REQUIRE(gr.obj_begin + gr.n_obj <= bk.n_obj_cap, "xm_affinity: bank %d holds fewer value planes than group %d needs", i, g);
            const int s = sg.nseg++;
            sg.bank[s] = i; sg.begin[s] = begin; sg.end[s] = bk.size; sg.tile0[s] = tiles; sg.col0[s] = cols;
            sg.origin[s] = begin & ~7;
            sg.shr[s] = bk.shrinkage; sg.usage[s] = (g == 0) ? bk.usage : nullptr;
            tiles += (bk.size - sg.origin[s] + TN - 1) / TN;
            cols += bk.size - begin;
        }
        for (int s = sg.nseg; s < 4; ++s) sg.tile0[s] = tiles;
        XM_REQUIRE(allow_small || cols >= a->top_k, "xm_affinity: group %d sees %d memory columns < top_k=%d (torch.topk would raise)", g,
                   cols, a->top_k);
    }
    return XM_OK;
}

extern "C" int xm_affinity_plan(const xm_affinity_args_t* a, void* host_plan_out, int64_t bytes) {
    XM_REQUIRE(a && host_plan_out && bytes >= K1_PLAN_BYTES, "xm_affinity_plan: need a %d-byte host buffer", K1_PLAN_BYTES);
    memset(host_plan_out, 0, K1_PLAN_BYTES);
    return k1_build_plan(a, (K1Seg*)host_plan_out);
}

extern "C" int xm_affinity_readout(const xm_affinity_args_t* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    XM_REQUIRE(a, "xm_affinity_readout: null args");
    XM_REQUIRE(a->hw > 0 && a->hw_pad == (a->hw + TQ - 1) / TQ * TQ, "xm_affinity_readout: hw_pad must be hw rounded up to 128");
    XM_REQUIRE(a->top_k > 0 && a->top_k <= XM_MAX_TOPK, "xm_affinity_readout: top_k must be in [1,%d]", XM_MAX_TOPK);
    XM_REQUIRE(a->n_groups > 0 && a->n_groups <= XM_MAX_GROUPS, "xm_affinity_readout: bad n_groups %d", a->n_groups);
    XM_REQUIRE(a->qp && a->bsq && a->workspace, "xm_affinity_readout: null query/workspace");
    XM_REQUIRE(a->readout_chw || a->readout_hwc, "xm_affinity_readout: no output buffer");
    XM_REQUIRE(a->workspace_bytes >= xm_affinity_workspace_bytes(a->hw, a->n_obj_total), "xm_affinity_readout: workspace too small");
    const int hw = a->hw, hw_pad = a->hw_pad;
    const int qtiles = hw_pad / TQ;

    tc5_debug_init();
    static XmPerDevice attr_token = {0};
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_scan<MODE_SLOTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem) + 1024));
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_scan<MODE_COLLECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem) + 1024));
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_readout_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem) + 1024));
    }

    uint8_t* ws = (uint8_t*)a->workspace;
    K1Seg* plan_dev = (K1Seg*)ws;        ws += K1_PLAN_BYTES;
    float* cand = (float*)ws;            ws += align_up((size_t)K1_MAX_SPLIT * hw_pad * LISTK * 4, 256);
    float* tau_lo = (float*)ws;          ws += align_up((size_t)hw_pad * 4, 256);
    float* tau = (float*)ws;             ws += align_up((size_t)hw_pad * 4, 256);
    float* inv_den = (float*)ws;         ws += align_up((size_t)hw_pad * 4, 256);
    float* partial = (float*)ws;

    if (!a->plan_is_resident) {
        // eager convenience path: build the table here and copy it (pageable source: staged before the call returns)
        K1Seg plan[XM_MAX_GROUPS];
        memset(plan, 0, sizeof(plan));
        const int rc = k1_build_plan(a, plan);
        if (rc != XM_OK) return rc;
        XM_CHECK_CUDA(cudaMemcpyAsync(plan_dev, plan, sizeof(K1Seg) * a->n_groups, cudaMemcpyHostToDevice, stream));
    }

    K1Maps maps;
    {
        uint64_t d[2] = {KP, (uint64_t)hw_pad};
        uint64_t s[1] = {KP * 2};
        uint32_t b[2] = {64, TQ};
        if (xm_make_tmap_f16(&maps.q, a->qp, 2, d, s, b)) return XM_ERR_CUDA;
    }
    for (int i = 0; i < 3; ++i) {
        const xm_bank_t& bk = a->banks[i];
        if (!bk.keys || bk.cap <= 0) {      // bank without an arena: alias the query map so the struct is fully initialised
            maps.k[i] = maps.q;
            maps.v[i] = maps.q;
            continue;
        }
        uint64_t d[2] = {KP, (uint64_t)bk.cap};
        uint64_t s[1] = {KP * 2};
        uint32_t b[2] = {64, TN};
        if (xm_make_tmap_f16(&maps.k[i], bk.keys, 2, d, s, b)) return XM_ERR_CUDA;
        uint64_t dv[3] = {(uint64_t)bk.cap, XM_CV, (uint64_t)bk.n_obj_cap};
        uint64_t sv[2] = {(uint64_t)bk.cap * 2, (uint64_t)bk.cap * 2 * XM_CV};
        uint32_t bv[3] = {TN, 128, 1};
        if (xm_make_tmap_f16(&maps.v[i], bk.values, 3, dv, sv, bv)) return XM_ERR_CUDA;
    }

    // launch shapes depend only on (hw, n_obj): one wave of CTAs, each owning a contiguous slice of column tiles
    const int sms = xm_num_sms();
    int nsplit1 = sms / qtiles;
    if (nsplit1 < 1) nsplit1 = 1;
    if (nsplit1 > K1_MAX_SPLIT / 2) nsplit1 = K1_MAX_SPLIT / 2;
    for (int g = 0; g < a->n_groups; ++g) {
        const xm_group_t& gr = a->groups[g];
        XM_REQUIRE(gr.n_obj > 0 && gr.obj_begin >= 0 && gr.obj_begin + gr.n_obj <= a->n_obj_total, "xm_affinity_readout: bad group %d objects", g);
        const K1Seg* sgp = plan_dev + g;
        // scan A (slot maxima) -> tau_lo ; scan B (collect > pred(tau_lo)) -> tau, 1/den
        XM_CHECK_CUDA(tc5_launch(k1_scan<MODE_SLOTMAX>, dim3(qtiles, nsplit1), dim3(SCAN_THREADS), sizeof(ScanSmem) + 1024, stream,
                                 maps, sgp, a->bsq, (const float*)nullptr, hw_pad, cand, g == 0 ? a->debug_scores : (float*)nullptr));
        XM_CHECK_CUDA(cudaGetLastError());
        XM_CHECK_CUDA(tc5_launch(k1_topk_merge, dim3((hw_pad + 3) / 4), dim3(128), 0, stream, (const float*)cand, nsplit1 * 2, hw, hw_pad,
                                 a->top_k, 0, tau_lo, (float*)nullptr));
        XM_CHECK_CUDA(cudaGetLastError());
        XM_CHECK_CUDA(tc5_launch(k1_scan<MODE_COLLECT>, dim3(qtiles, nsplit1), dim3(SCAN_THREADS), sizeof(ScanSmem) + 1024, stream,
                                 maps, sgp, a->bsq, (const float*)tau_lo, hw_pad, cand, (float*)nullptr));
        XM_CHECK_CUDA(cudaGetLastError());
        XM_CHECK_CUDA(tc5_launch(k1_topk_merge, dim3((hw_pad + 3) / 4), dim3(128), 0, stream, (const float*)cand, nsplit1 * 2, hw, hw_pad,
                                 a->top_k, 1, tau, inv_den));
        XM_CHECK_CUDA(cudaGetLastError());

        // pass 2
        const int ctas_per_slice = qtiles * 2 * gr.n_obj;
        int nsplit2 = sms / ctas_per_slice;               // one wave
        nsplit2 = nsplit2 < 1 ? 1 : nsplit2;
        if (nsplit2 > K1_MAX_SPLIT) nsplit2 = K1_MAX_SPLIT;
        XM_CHECK_CUDA(tc5_launch(k1_readout_pass2, dim3(qtiles, 2 * gr.n_obj, nsplit2), dim3(P2_THREADS), sizeof(P2Smem) + 1024, stream,
                                 maps, sgp, a->bsq, (const float*)tau, (const float*)inv_den, hw_pad, gr.obj_begin, gr.n_obj, g == 0 ? 1 : 0, partial));
        XM_CHECK_CUDA(cudaGetLastError());
        xm_count_launches(6);
        XM_CHECK_CUDA(tc5_launch(k1_finish, dim3((hw + 31) / 32, XM_CV / 32, gr.n_obj), dim3(32, 8), 0, stream,
                                 (const float*)partial, nsplit2, gr.n_obj, hw, hw_pad, gr.obj_begin, (__half*)a->readout_chw, (__half*)a->readout_hwc));
        XM_CHECK_CUDA(cudaGetLastError());
    }
    return XM_OK;
}


// ---------------------------------------------------------------------------------------------
// T-sharded memory read: the banks of ONE long video are distributed over R ranks by stored frame.
// Every rank holds the same query; the host interleaves three small NCCL collectives between these stages:
//   stage_a : local slot maxima -> local lower bound                     ... all_reduce(MAX)  tau_lo[hw_pad]
//   stage_b : local scores > pred(tau_lo) -> local 32 largest per query  ... all_gather       top32[R][hw_pad][32]
//   merge   : exact global tau, 1/den from the gathered records (every rank, identical result)
//   stage_c : local P.V with the GLOBAL normalisers -> fp32 partial      ... all_reduce(SUM)  readout_f32
//   cast    : fp32 -> fp16 CHW / NHWC
// Single object group per call (groups[0]); usage stays local to the rank that owns the column.
// ---------------------------------------------------------------------------------------------
struct TshardCtx {
    K1Maps maps; K1Seg* plan_dev; float *cand, *tau_lo, *tau, *inv_den, *partial; int qtiles, nsplit1, nsplit2;
};

static int tshard_setup(const xm_affinity_args_t* a, cudaStream_t stream, TshardCtx& c, bool upload_plan) {
    XM_REQUIRE(a && a->n_groups == 1, "xm_affinity_tshard: exactly one object group per call");
    XM_REQUIRE(a->hw > 0 && a->hw_pad == (a->hw + TQ - 1) / TQ * TQ, "xm_affinity_tshard: hw_pad must be hw rounded up to 128");
    XM_REQUIRE(a->top_k > 0 && a->top_k <= XM_MAX_TOPK && a->qp && a->bsq && a->workspace, "xm_affinity_tshard: bad arguments");
    XM_REQUIRE(a->workspace_bytes >= xm_affinity_workspace_bytes(a->hw, a->n_obj_total), "xm_affinity_tshard: workspace too small");
    tc5_debug_init();
    static XmPerDevice attr_token = {0};
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_scan<MODE_SLOTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem) + 1024));
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_scan<MODE_COLLECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem) + 1024));
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_readout_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem) + 1024));
    }
    const int hw_pad = a->hw_pad;
    uint8_t* ws = (uint8_t*)a->workspace;
    c.plan_dev = (K1Seg*)ws;        ws += K1_PLAN_BYTES;
    c.cand = (float*)ws;            ws += align_up((size_t)K1_MAX_SPLIT * hw_pad * LISTK * 4, 256);
    c.tau_lo = (float*)ws;          ws += align_up((size_t)hw_pad * 4, 256);
    c.tau = (float*)ws;             ws += align_up((size_t)hw_pad * 4, 256);
    c.inv_den = (float*)ws;         ws += align_up((size_t)hw_pad * 4, 256);
    c.partial = (float*)ws;
    if (upload_plan) {
        K1Seg plan[XM_MAX_GROUPS];
        memset(plan, 0, sizeof(plan));
        const int rc = k1_build_plan(a, plan, /*allow_small=*/true);
        if (rc != XM_OK) return rc;
        XM_CHECK_CUDA(cudaMemcpyAsync(c.plan_dev, plan, sizeof(K1Seg), cudaMemcpyHostToDevice, stream));
    }
    {
        uint64_t d[2] = {KP, (uint64_t)hw_pad}; uint64_t st[1] = {KP * 2}; uint32_t b[2] = {64, TQ};
        if (xm_make_tmap_f16(&c.maps.q, a->qp, 2, d, st, b)) return XM_ERR_CUDA;
    }
    for (int i = 0; i < 3; ++i) {
        const xm_bank_t& bk = a->banks[i];
        if (!bk.keys || bk.cap <= 0) { c.maps.k[i] = c.maps.q; c.maps.v[i] = c.maps.q; continue; }
        uint64_t d[2] = {KP, (uint64_t)bk.cap}; uint64_t st[1] = {KP * 2}; uint32_t b[2] = {64, TN};
        if (xm_make_tmap_f16(&c.maps.k[i], bk.keys, 2, d, st, b)) return XM_ERR_CUDA;
        uint64_t dv[3] = {(uint64_t)bk.cap, XM_CV, (uint64_t)bk.n_obj_cap};
        uint64_t sv[2] = {(uint64_t)bk.cap * 2, (uint64_t)bk.cap * 2 * XM_CV};
        uint32_t bv[3] = {TN, 128, 1};
        if (xm_make_tmap_f16(&c.maps.v[i], bk.values, 3, dv, sv, bv)) return XM_ERR_CUDA;
    }
    const int sms = xm_num_sms();
    c.qtiles = hw_pad / TQ;
    c.nsplit1 = sms / c.qtiles; if (c.nsplit1 < 1) c.nsplit1 = 1; if (c.nsplit1 > K1_MAX_SPLIT / 2) c.nsplit1 = K1_MAX_SPLIT / 2;
    const int per = c.qtiles * 2 * a->groups[0].n_obj;
    c.nsplit2 = sms / per; if (c.nsplit2 < 1) c.nsplit2 = 1; if (c.nsplit2 > K1_MAX_SPLIT) c.nsplit2 = K1_MAX_SPLIT;
    return XM_OK;
}

extern "C" int xm_affinity_tshard_stage_a(const xm_affinity_args_t* a, float* tau_lo_local, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TshardCtx c;
    int rc = tshard_setup(a, stream, c, true);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau_lo_local, "xm_affinity_tshard_stage_a: null output");
    XM_CHECK_CUDA(tc5_launch(k1_scan<MODE_SLOTMAX>, dim3(c.qtiles, c.nsplit1), dim3(SCAN_THREADS), sizeof(ScanSmem) + 1024, stream,
                             c.maps, (const K1Seg*)c.plan_dev, a->bsq, (const float*)nullptr, a->hw_pad, c.cand, (float*)nullptr));
    XM_CHECK_CUDA(tc5_launch(k1_topk_merge, dim3((a->hw_pad + 3) / 4), dim3(128), 0, stream, (const float*)c.cand, c.nsplit1 * 2, a->hw,
                             a->hw_pad, a->top_k, 0, tau_lo_local, (float*)nullptr));
    xm_count_launches(2);
    return XM_OK;
}

extern "C" int xm_affinity_tshard_stage_b(const xm_affinity_args_t* a, const float* tau_lo_global, float* top32_local, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TshardCtx c;
    int rc = tshard_setup(a, stream, c, false);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau_lo_global && top32_local, "xm_affinity_tshard_stage_b: null pointer");
    XM_CHECK_CUDA(tc5_launch(k1_scan<MODE_COLLECT>, dim3(c.qtiles, c.nsplit1), dim3(SCAN_THREADS), sizeof(ScanSmem) + 1024, stream,
                             c.maps, (const K1Seg*)c.plan_dev, a->bsq, tau_lo_global, a->hw_pad, c.cand, (float*)nullptr));
    XM_CHECK_CUDA(tc5_launch(k1_export_top32, dim3((a->hw_pad + 3) / 4), dim3(128), 0, stream, (const float*)c.cand, c.nsplit1 * 2, a->hw_pad,
                             top32_local));
    xm_count_launches(2);
    return XM_OK;
}

extern "C" int xm_affinity_tshard_merge(const float* top32_all, int32_t n_ranks, int32_t hw, int32_t hw_pad, int32_t top_k, float* tau,
                                        float* inv_den, void* stream_) {
    XM_REQUIRE(top32_all && tau && inv_den && n_ranks >= 1 && n_ranks <= 32, "xm_affinity_tshard_merge: bad arguments (<= 32 ranks)");
    XM_CHECK_CUDA(tc5_launch(k1_topk_merge, dim3((hw_pad + 3) / 4), dim3(128), 0, (cudaStream_t)stream_, top32_all, n_ranks, hw, hw_pad, top_k, 1,
                             tau, inv_den));
    xm_count_launches(1);
    return XM_OK;
}

extern "C" int xm_affinity_tshard_stage_c(const xm_affinity_args_t* a, const float* tau, const float* inv_den, float* readout_f32,
                                          void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    TshardCtx c;
    int rc = tshard_setup(a, stream, c, false);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau && inv_den && readout_f32, "xm_affinity_tshard_stage_c: null pointer");
    const xm_group_t& gr = a->groups[0];
    XM_CHECK_CUDA(tc5_launch(k1_readout_pass2, dim3(c.qtiles, 2 * gr.n_obj, c.nsplit2), dim3(P2_THREADS), sizeof(P2Smem) + 1024, stream,
                             c.maps, (const K1Seg*)c.plan_dev, a->bsq, tau, inv_den, a->hw_pad, gr.obj_begin, gr.n_obj, 1, c.partial));
    const size_t plane = (size_t)gr.n_obj * XM_CV * a->hw_pad;
    XM_CHECK_CUDA(tc5_launch(k1_sum_splits, dim3(xm_num_sms() * 4), dim3(256), 0, stream, (const float*)c.partial, c.nsplit2, plane, readout_f32));
    xm_count_launches(2);
    return XM_OK;
}

extern "C" int xm_affinity_tshard_cast(const float* readout_f32, int32_t n_obj, int32_t hw, int32_t hw_pad, void* readout_chw,
                                       void* readout_hwc, void* stream_) {
    XM_REQUIRE(readout_f32 && (readout_chw || readout_hwc) && n_obj >= 1, "xm_affinity_tshard_cast: bad arguments");
    XM_CHECK_CUDA(tc5_launch(k1_finish, dim3((hw + 31) / 32, XM_CV / 32, n_obj), dim3(32, 8), 0, (cudaStream_t)stream_, readout_f32, 1, n_obj, hw,
                             hw_pad, 0, (__half*)readout_chw, (__half*)readout_hwc));
    xm_count_launches(1);
    return XM_OK;
}
