// k1_affinity.cu — fused memory read for XMem++ :  similarity -> top-k softmax -> value readout (+usage)
//
// Replaces the reference sequence get_similarity (model/memory_util.py:7-39) -> do_softmax top-k branch
// (:41-54, no max-subtraction) -> usage = affinity.sum (:62-63) -> v @ affinity (memory_manager.py:57-59)
// over the concatenated long-term | working | permanent banks (memory_manager.py:82-128,143-182),
// without ever materialising the N x HW similarity / affinity matrices.
//
// ONE persistent kernel per object group (k1_fused), one CTA per SM, all CTAs co-resident; the phases are separated by
// inter-CTA barriers on global counters:
//   sweep A   S = Qp.Kp^T tiles (tcgen05, 256 queries x 128 memory columns per key tile: two query tiles share every
//             key tile, which halves the per-SM TMA ingest), every scan thread keeps 16 branch-free running maxima of
//             disjoint column subsets ("slot maxima").
//   merge A   tau_lo[q] = k-th largest slot maximum  (a LOWER bound of the true k-th largest score: the maxima belong
//             to distinct memory columns).
//   sweep B   the same S tiles again (bit-identical instruction stream); every score > pred(tau_lo) is appended with its
//             column index to a per-(slice, query) list (exact top-64 of the slice when more than 64 survive).
//   merge B   exact top-k per query on 64-bit keys (score, lowest column first), weights exp(S)/sum exp(S) (no max
//             subtraction, memory_util.py:48-49), usage atomics, final (column, weight) list per query.
//   readout   O[q, c] += P[q, n] V[n, c]  as a dense tcgen05 contraction (the reference's dense v @ affinity):
//             CTA = 256 queries x 256 value channels x a slice of the memory columns, O in TMEM (2 x 256 columns),
//             V tiles by TMA, the P tile is zero except for the listed entries, which are scattered into the swizzled
//             smem operand (and cleared again after use).  No score is recomputed in this phase.
//   reduce    sum of the column-slice partials -> fp16 CHW / NHWC (or fp32 for the T-sharded mode).
// Math.  Packed operands Kp[n] = (k_n^2, k_n), Qp[q] = (-e_q, 2 k_q e_q) (fp16, 128 wide):
//     S'[q,n] = Qp[q] . Kp[n]                                (tcgen05.mma, fp32 accumulate)
//     S[q,n]  = (S'[q,n] - bsq[q]) * shrinkage[n] / sqrt(64)  (bsq = sum_c e k_q^2, fp32)
#include <cfloat>
#include <cmath>
#include <cstring>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

namespace {

constexpr int TQ = 128;                 // queries per q-tile (UMMA M / TMEM lanes)
constexpr int QPAIR = 256;              // queries per CTA: two q-tiles share every key / value tile
constexpr int TN = 128;                 // sweep: memory columns per key tile
constexpr int TK = 64;                  // readout: memory columns per value tile ("k-tile")
constexpr int KP = 128;                 // packed key width
constexpr int LISTK = XM_MAX_TOPK;      // 32
constexpr int NSLOT = 16;               // slot maxima per scan thread
constexpr int LCAP = 64;                // candidate capacity per (column slice, query)
constexpr int MAX_SLICE1 = 32;          // column slices per query pair in the sweeps
constexpr int P1_KSTAGES = 4;
constexpr int P2_VSTAGES = 3;
constexpr int P2_MAXKT = 2048;          // k-tiles per readout batch
constexpr int P2_MAXENT = QPAIR * LISTK;
constexpr int NWORK = 16;               // worker warps (scan / merge / P builders / epilogue / reduce)
constexpr int NTHREADS = 64 + NWORK * 32;
constexpr int SCR_CAP = 512;            // merge scratch entries per warp
constexpr int MAX_ROWS_TABLE = 160;     // readout rows that can get more than one column slice
constexpr int NCTR = 256;               // barrier counters: [0] grid, [1] exit, [8 + pair] query-pair groups
constexpr float LOG2E = 1.4426950408889634f;

constexpr int MODE_SWEEP_A = 1;         // sweep A + merge A  -> tau_lo
constexpr int MODE_SWEEP_B = 2;         // sweep B            -> candidate lists
constexpr int MODE_SELECT = 4;          // merge B            -> final (column, weight) lists + usage
constexpr int MODE_READOUT = 8;         // readout + reduce
constexpr int MODE_EXPORT32 = 16;       // T-shard: the 32 largest local candidate scores per query
constexpr int MODE_EXT_TAU = 32;        // T-shard: selection by the GLOBAL tau / 1/den (all local scores >= tau)
constexpr int MODE_FULL = MODE_SWEEP_A | MODE_SWEEP_B | MODE_SELECT | MODE_READOUT;

struct alignas(64) K1Maps {
    CUtensorMap q;       // [128, hw_pad]            box [64,128]
    CUtensorMap k[3];    // [128, cap_b]             box [64,128]
    CUtensorMap v[3];    // [cap_b, 512, n_obj_cap]  box [64,256,1]
};

// Column ranges of one object group (device-resident "plan": bank sizes are not kernel arguments, so a recorded CUDA graph
// stays valid while the memory grows).
struct K1Seg {
    int nseg;
    int bank[3];
    int begin[3];
    int end[3];
    int origin[3];       // begin rounded down to 8 columns: TMA needs 16-byte aligned starts on the contiguous (column) axis of V
    int cap[3];
    int t128[4];         // prefix sum of 128-column sweep tiles per segment
    int t64[4];          // prefix sum of 64-column readout tiles per segment
    int col0[3];         // first column of the segment inside the group's concatenated column index (debug dump)
    int pad_;
    const float* shr[3];
    float* usage[3];
};

struct K1Params {
    const K1Seg* seg;
    const float* bsq;
    int hw, hw_pad, top_k;
    int obj_begin, n_obj;
    int do_usage;
    int mode;
    int qtiles, qpairs, nslice1;
    int n_rows, n_items;
    unsigned* ctr;
    float* candA;            // [qpairs*256][nslice1*32]
    float* tau_lo;           // [qpairs*256]
    uint2* lists;            // [qpairs*256][nslice1][LCAP]  (score bits, linear column)
    int* lcnt;               // [qpairs*256][nslice1]
    uint2* fin;              // [qpairs*256][32]             (linear column or -1, weight bits)
    float* partial;          // [n_items][256][256]
    const float* tau_ext;    // MODE_EXT_TAU
    const float* inv_ext;
    float* top32_out;        // MODE_EXPORT32
    float* out_f32;          // fp32 [n_obj][512][hw_pad] (T-shard) or null
    __half* out_chw;
    __half* out_hwc;
    int out_obj_total;
    float* dbg;
    uint8_t row_slices[MAX_ROWS_TABLE];
};

__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * LOG2E));
    return y;
}
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__device__ __forceinline__ float score2(uint32_t acc_bits, float bsq8, float ms) {
    // ((S' - b_sq) * shrinkage) / 8 == (S'/8 - b_sq/8) * shrinkage exactly (power-of-two scaling commutes with rounding)
    return fmaf(__uint_as_float(acc_bits), 0.125f, -bsq8) * ms;
}

__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
// global -> shared bulk copy (1-D TMA) completing on an mbarrier; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

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
// shared memory: the sweeps and the readout alias one dynamic buffer
// ---------------------------------------------------------------------------------------------
struct KStage {                                              // one key tile: released by the MMA commit AND the scan warps
    alignas(1024) uint8_t k[2][TN * 128];                    // [K half] 128 column rows x 128 B
    alignas(16) float ms[TN];                                // shrinkage of the tile's columns (bulk copy on the same barrier)
};
struct SweepSmem {
    alignas(1024) uint8_t q[2][2][TQ * 128];                 // [q-tile of the pair][K half] 128 rows x 128 B
    KStage st[P1_KSTAGES];                                   // merge scratch aliases this
    int cnt[QPAIR];                                          // sweep B: entries in this CTA's list of each query
    int lock[QPAIR];
    float lmin[QPAIR];                                       // smallest listed score once a list is full (else -inf)
};
struct ReadSmem {
    alignas(1024) uint8_t v[P2_VSTAGES][256 * 128];          // [stage] 256 channel rows x (64 columns = 128 B); reduce scratch aliases this
    alignas(1024) uint8_t p[2][2][TQ * 128];                 // [buffer][q-tile] 128 query rows x (64 columns = 128 B)
    uint32_t ent[P2_MAXENT];                                 // (q-tile << 13 | row << 6 | column) | fp16 weight << 16, sorted by k-tile
    uint32_t cur[P2_MAXKT + 1];
    uint16_t off[P2_MAXKT + 2];
};
struct CommonSmem {
    alignas(8) uint64_t qfull;
    uint64_t kfull[P1_KSTAGES], kempty[P1_KSTAGES];
    uint64_t sfull[2][2], sempty[2][2];                      // [q-tile][tile parity]
    uint64_t vfull[P2_VSTAGES], vempty[P2_VSTAGES];
    uint64_t pfull[2], pempty[2];
    uint64_t ofull, oempty;
    uint32_t tmem_base;
    int row_item0[MAX_ROWS_TABLE + 1];
    K1Seg sg;
};
constexpr size_t SMEM_MAIN = sizeof(SweepSmem) > sizeof(ReadSmem) ? sizeof(SweepSmem) : sizeof(ReadSmem);
constexpr size_t SMEM_TOTAL = SMEM_MAIN + sizeof(CommonSmem) + 1024;
static_assert(NWORK * SCR_CAP * 8 <= sizeof(SweepSmem::st), "merge scratch must fit in the key stages");
static_assert(NWORK * 32 * 33 * 4 <= sizeof(ReadSmem::v), "reduce scratch must fit in the value stages");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Barrier among the CTAs that increment `ctr` (all co-resident: one CTA per SM, grid <= #SMs).  Bounded: a protocol bug or a
// CTA that never became resident traps instead of hanging the GPU.
__device__ void cta_group_barrier(unsigned* ctr, unsigned target, int tag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned spins = 0;
        while (ld_acquire_u32(ctr) < target) {
            __nanosleep(40);
            if (++spins > (1u << 25)) mbar_timeout(tag, target);
        }
        __threadfence();
    }
    __syncthreads();
}

// sweep tile t (128 columns) -> segment s, first bank column `col`, valid lanes [lo, hi) of the tile
__device__ __forceinline__ void locate_tile128(const K1Seg& sg, int t, int& s, int& col, int& lo, int& hi) {
    s = 0;
    if (sg.nseg > 1 && t >= sg.t128[1]) s = 1;
    if (sg.nseg > 2 && t >= sg.t128[2]) s = 2;
    col = sg.origin[s] + (t - sg.t128[s]) * TN;
    lo = max(0, sg.begin[s] - col);
    hi = min(TN, sg.end[s] - col);
}
// linear column index (what the candidate lists store): 64 * (k-tile) + offset inside the k-tile
__device__ __forceinline__ int linear_col(const K1Seg& sg, int s, int col) { return sg.t64[s] * TK + (col - sg.origin[s]); }
__device__ __forceinline__ void unlinear_col(const K1Seg& sg, int lin, int& s, int& col) {
    const int kt = lin >> 6;
    s = 0;
    if (sg.nseg > 1 && kt >= sg.t64[1]) s = 1;
    if (sg.nseg > 2 && kt >= sg.t64[2]) s = 2;
    col = sg.origin[s] + (lin - sg.t64[s] * TK);
}

// Append (score, column) to this CTA's list of query `ql`; when the list is full keep the LCAP largest (exact top-LCAP of the
// slice).  The four scan threads of a query (different warps) serialise on a shared-memory lock; appends are rare (~0.2 % of
// the scores), replace-min only in the pathological case of > 64 survivors in one slice.
__device__ __noinline__ void list_append(float sc, int lin, int ql, SweepSmem& sm, uint2* glist, float& thr) {
    while (atomicCAS(&sm.lock[ql], 0, 1) != 0) { }
    __threadfence_block();
    volatile int* cntp = &sm.cnt[ql];
    volatile uint2* lst = glist;
    const int c = *cntp;
    if (c < LCAP) {
        lst[c].x = __float_as_uint(sc); lst[c].y = (uint32_t)lin;
        *cntp = c + 1;
        if (c + 1 == LCAP) {
            float m = sc;
            for (int u = 0; u < LCAP - 1; ++u) m = fminf(m, __uint_as_float(lst[u].x));
            *reinterpret_cast<volatile float*>(&sm.lmin[ql]) = m;
            thr = fmaxf(thr, m);
        }
    } else {
        float m1 = INFINITY, m2 = INFINITY; int p1 = 0;
        for (int u = 0; u < LCAP; ++u) {
            const float v = __uint_as_float(lst[u].x);
            if (v < m1) { m2 = m1; m1 = v; p1 = u; } else if (v < m2) { m2 = v; }
        }
        if (sc > m1) {
            lst[p1].x = __float_as_uint(sc); lst[p1].y = (uint32_t)lin;
            m1 = fminf(m2, sc);
        }
        *reinterpret_cast<volatile float*>(&sm.lmin[ql]) = m1;
        thr = fmaxf(thr, m1);
    }
    __threadfence_block();
    atomicExch(&sm.lock[ql], 0);
}

// ---------------------------------------------------------------------------------------------
// warp-level selection on 64-bit keys (score image << 32 | ~column): keys are distinct, so "the m largest" is well defined
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_key(uint32_t score_bits, uint32_t lin) {
    return (static_cast<uint64_t>(f2ord(__uint_as_float(score_bits))) << 32) | static_cast<uint64_t>(0xffffffffu - lin);
}
// m-th largest key among scr[0..n) (n > m >= 1); all lanes return the same value
__device__ uint64_t warp_kth_key(const uint64_t* scr, int n, int m, int lane) {
    uint64_t t = 0;
#pragma unroll 1
    for (int bit = 63; bit >= 0; --bit) {
        const uint64_t trial = t | (1ull << bit);
        int c = 0;
        for (int e = lane; e < n; e += 32) c += (scr[e] >= trial) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= m) t = trial;
    }
    return t;
}
// keep the keys >= thr at the front of scr (in place, order preserved); returns the number kept
__device__ int warp_compact_ge(uint64_t* scr, int n, uint64_t thr, int lane) {
    int w = 0;
    for (int base = 0; base < n; base += 32) {
        const int e = base + lane;
        const uint64_t key = (e < n) ? scr[e] : 0ull;
        const bool keep = (e < n) && key >= thr;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) scr[w + __popc(b & ((1u << lane) - 1u))] = key;
        w += __popc(b);
        __syncwarp();
    }
    return w;
}
// gather the per-slice lists of query q into scr, never holding more than SCR_CAP entries: when the scratch would overflow it
// is first reduced to its `keep` largest keys (exact).  min_score_ord: entries below it are dropped (T-shard selection).
__device__ int warp_gather_lists(const K1Params& p, int q, uint64_t* scr, int keep, uint32_t min_score_ord, int lane) {
    int n = 0;
    for (int s = 0; s < p.nslice1; ++s) {
        const int c = min(LCAP, __ldcg(p.lcnt + (size_t)q * p.nslice1 + s));
        if (c == 0) continue;
        if (n + c > SCR_CAP) {
            if (n > keep) {
                const uint64_t kth = warp_kth_key(scr, n, keep, lane);
                __syncwarp();
                n = warp_compact_ge(scr, n, kth, lane);
            }
        }
        const uint2* src = p.lists + ((size_t)q * p.nslice1 + s) * LCAP;
        for (int base = 0; base < c; base += 32) {
            const int e = base + lane;
            uint2 ent = make_uint2(0u, 0u);
            bool ok = e < c;
            if (ok) { ent = __ldcg(src + e); ok = f2ord(__uint_as_float(ent.x)) >= min_score_ord; }
            const unsigned b = __ballot_sync(0xffffffffu, ok);
            if (ok) scr[n + __popc(b & ((1u << lane) - 1u))] = make_key(ent.x, ent.y);
            n += __popc(b);
        }
        __syncwarp();
    }
    return n;
}
// descending bitonic sort of one key per lane (32 keys)
__device__ __forceinline__ uint64_t warp_sort_desc(uint64_t key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = ((lane & k) == 0);            // this block sorts descending when up
            const bool lower = ((lane & j) == 0);
            const bool take_max = (up == lower);
            key = take_max ? (key > other ? key : other) : (key < other ? key : other);
        }
    }
    return key;
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
k1_fused(const __grid_constant__ K1Maps maps, const __grid_constant__ K1Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    SweepSmem& sw = *reinterpret_cast<SweepSmem*>(base);
    ReadSmem& rd = *reinterpret_cast<ReadSmem*>(base);
    CommonSmem& cm = *reinterpret_cast<CommonSmem*>(base + SMEM_MAIN);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ww = warp - 2;                                   // worker warp index (valid when warp >= 2)
    const int G = gridDim.x;
    const int cta = blockIdx.x;

    // ------------------------------------------------------------------ prologue (overlaps the previous kernel's tail)
    // sweep geometry of this CTA (needed for the barrier counts)
    const int S1 = p.nslice1;
    const bool in_sweep = cta < p.qpairs * S1;
    const int pair = in_sweep ? cta / S1 : 0, slice = in_sweep ? cta % S1 : 0;
    const int nqh = (pair * 2 + 1 < p.qtiles) ? 2 : 1;       // q-tiles of this pair that exist
    if (threadIdx.x == 0) {
        mbar_init(&cm.qfull, 1);
        for (int i = 0; i < P1_KSTAGES; ++i) { mbar_init(&cm.kfull[i], 1); mbar_init(&cm.kempty[i], 1 + 4 * nqh); }
        for (int b = 0; b < 2; ++b) {
            for (int h = 0; h < 2; ++h) { mbar_init(&cm.sfull[h][b], 1); mbar_init(&cm.sempty[h][b], 4); }
            mbar_init(&cm.pfull[b], 1); mbar_init(&cm.pempty[b], 1);
        }
        for (int i = 0; i < P2_VSTAGES; ++i) { mbar_init(&cm.vfull[i], 1); mbar_init(&cm.vempty[i], 1); }
        mbar_init(&cm.ofull, 1); mbar_init(&cm.oempty, NWORK);
        fence_mbar_init();
        tma_prefetch_desc(&maps.q);
        // readout rows -> first work item
        int acc = 0;
        if (p.n_rows <= MAX_ROWS_TABLE) {
            for (int r = 0; r < p.n_rows; ++r) { cm.row_item0[r] = acc; acc += p.row_slices[r]; }
            cm.row_item0[p.n_rows] = acc;
        }
    }
    if (warp == 1) { tmem_alloc(&cm.tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = cm.tmem_base;
    pdl_wait();                                               // everything below reads what preceding kernels wrote
    if (threadIdx.x < sizeof(K1Seg) / 4) reinterpret_cast<uint32_t*>(&cm.sg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(p.seg)[threadIdx.x];
    __syncthreads();
    const K1Seg& sg = cm.sg;

    unsigned pair_uses = 0, grid_uses = 0;

    // ------------------------------------------------------------------ sweeps (A: slot maxima, B: collect)
    const int total128 = sg.t128[sg.nseg];
    const int t_begin = (int)(((long long)total128 * slice) / S1);
    const int t_end = (int)(((long long)total128 * (slice + 1)) / S1);
    const int nt = in_sweep ? (t_end - t_begin) : 0;
    unsigned* pair_ctr = p.ctr + 8 + pair;
    int sweeps_done = 0;
    bool q_loaded = false;

    for (int sweep = 0; sweep < 2; ++sweep) {
        const int this_mode = sweep == 0 ? MODE_SWEEP_A : MODE_SWEEP_B;
        if (!(p.mode & this_mode)) continue;
        if (in_sweep) {
            const int it0 = sweeps_done * nt;                 // mbarrier phases continue across the two sweeps
            // uses of S buffer b before this sweep: tiles i < nt with (i & 1) == b, per completed sweep
            const int uses0[2] = {sweeps_done * ((nt + 1) >> 1), sweeps_done * (nt >> 1)};
            if (sweep == 1 && threadIdx.x < QPAIR) { sw.cnt[threadIdx.x] = 0; sw.lock[threadIdx.x] = 0; sw.lmin[threadIdx.x] = -INFINITY; }
            __syncthreads();
            if (warp == 0) {
                if (lane == 0) {
                    if (!q_loaded) {
                        mbar_expect_tx(&cm.qfull, nqh * 2 * TQ * 128);
                        for (int h = 0; h < nqh; ++h) {
                            tma_load_2d(sw.q[h][0], &maps.q, &cm.qfull, 0, (pair * 2 + h) * TQ);
                            tma_load_2d(sw.q[h][1], &maps.q, &cm.qfull, 64, (pair * 2 + h) * TQ);
                        }
                    }
                    for (int i = 0; i < nt; ++i) {
                        int s, col, lo, hi;
                        locate_tile128(sg, t_begin + i, s, col, lo, hi);
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const uint32_t ms_bytes = (uint32_t)min(TN, sg.cap[s] - col) * 4u;
                        mbar_wait(&cm.kempty[st], ph ^ 1, 2);
                        mbar_expect_tx(&cm.kfull[st], 2 * TN * 128 + ms_bytes);
                        const CUtensorMap* km = &maps.k[sg.bank[s]];
                        tma_load_2d(sw.st[st].k[0], km, &cm.kfull[st], 0, col);
                        tma_load_2d(sw.st[st].k[1], km, &cm.kfull[st], 64, col);
                        bulk_g2s(sw.st[st].ms, sg.shr[s] + col, ms_bytes, &cm.kfull[st]);
                    }
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    constexpr uint32_t idesc = make_idesc_f16(TQ, TN);
                    if (!q_loaded) mbar_wait(&cm.qfull, 0, 1);
                    for (int i = 0; i < nt; ++i) {
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const int b = i & 1, use = uses0[b] + (i >> 1);
                        mbar_wait(&cm.kfull[st], ph, 3);
                        for (int h = 0; h < nqh; ++h) {
                            mbar_wait(&cm.sempty[h][b], (use & 1) ^ 1, 4);
                            tc_fence_after();
#pragma unroll
                            for (int kh = 0; kh < 2; ++kh)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint64_t a = make_desc_sw128(smem_u32(sw.q[h][kh]) + j * 32);
                                    const uint64_t bd = make_desc_sw128(smem_u32(sw.st[st].k[kh]) + j * 32);
                                    mma_f16_ss(tmem + (h * 2 + b) * TN, a, bd, idesc, (kh | j) ? 1u : 0u);
                                }
                            mma_commit(&cm.sfull[h][b]);
                        }
                        mma_commit(&cm.kempty[st]);
                    }
                }
            } else {
                // scan warps: warp group wg = (q-tile h, tile parity b); a thread owns one query (TMEM lane) and sees
                // every column of the tiles with its parity
                const int wg = ww >> 2, h = wg >> 1, b = wg & 1;
                const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
                const int row = quad * 32 + lane;
                const int ql = h * TQ + row;                   // query inside the pair
                const int q = pair * QPAIR + ql;
                if (h < nqh) {
                    const float bsq8 = p.bsq[q] * 0.125f;
                    float slot[NSLOT];
#pragma unroll
                    for (int j = 0; j < NSLOT; ++j) slot[j] = -INFINITY;
                    float thr = -INFINITY;
                    if (sweep == 1) {
                        const float t = (p.mode & MODE_EXT_TAU) ? p.tau_ext[q] : __ldcg(p.tau_lo + q);
                        thr = (t == -INFINITY) ? -INFINITY : ((t == INFINITY) ? FLT_MAX : __uint_as_float(
                                  t > 0.f ? __float_as_uint(t) - 1u : (t < 0.f ? __float_as_uint(t) + 1u : 0x80000001u)));   // pred(tau_lo)
                    }
                    uint2* glist = p.lists + ((size_t)q * S1 + slice) * LCAP;      // this (query, slice)'s candidate list
                    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16) + (h * 2 + b) * TN;
                    for (int i = b; i < nt; i += 2) {
                        int s, col, lo, hi;
                        locate_tile128(sg, t_begin + i, s, col, lo, hi);
                        const int use = uses0[b] + (i >> 1);
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const float* msp = sw.st[st].ms;
                        if (sweep == 1) thr = fmaxf(thr, *reinterpret_cast<volatile float*>(&sw.lmin[ql]));
                        mbar_wait(&cm.kfull[st], ph, 6);              // the shrinkage slice arrived with the key tile
                        mbar_wait(&cm.sfull[h][b], use & 1, 5);
                        tc_fence_after();
                        const bool full_tile = (lo <= 0) && (hi >= TN);
                        const int lin0 = linear_col(sg, s, col);
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(trow + c * 32, r);
                            float ms[32];
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 f = *reinterpret_cast<const float4*>(msp + c * 32 + j);
                                ms[j] = f.x; ms[j + 1] = f.y; ms[j + 2] = f.z; ms[j + 3] = f.w;
                            }
                            tmem_ld_wait();
                            float sc[32];
                            if (full_tile) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) sc[j] = score2(r[j], bsq8, ms[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int n = c * 32 + j;
                                    sc[j] = (n >= lo && n < hi) ? score2(r[j], bsq8, ms[j]) : -INFINITY;
                                }
                            }
                            if (sweep == 0) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) slot[j & (NSLOT - 1)] = fmaxf(slot[j & (NSLOT - 1)], sc[j]);
                                if (p.dbg) {                   // tests / selector only (warp-uniform)
                                    float* dbg = p.dbg + ((ptrdiff_t)sg.col0[s] + (col + c * 32 - sg.begin[s])) * (ptrdiff_t)p.hw_pad + q;
#pragma unroll
                                    for (int j = 0; j < 32; ++j) if (sc[j] != -INFINITY) dbg[(ptrdiff_t)j * p.hw_pad] = sc[j];
                                }
                            } else {
                                uint32_t h0 = 0u, h1 = 0u, h2 = 0u, h3 = 0u;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    h0 |= (sc[j] > thr) ? (1u << j) : 0u;
                                    h1 |= (sc[8 + j] > thr) ? (256u << j) : 0u;
                                    h2 |= (sc[16 + j] > thr) ? (65536u << j) : 0u;
                                    h3 |= (sc[24 + j] > thr) ? (16777216u << j) : 0u;
                                }
                                const uint32_t hit = h0 | h1 | h2 | h3;
                                uint32_t any = __reduce_or_sync(0xffffffffu, hit);
                                while (any) {                  // ~2 columns per 32x32 block: re-read that column for the warp
                                    const int j = __ffs(any) - 1;
                                    any &= any - 1u;
                                    uint32_t rv;
                                    tmem_ld_32x32b_x1(trow + c * 32 + j, rv);
                                    tmem_ld_wait();
                                    if ((hit >> j) & 1u) {
                                        const float v = score2(rv, bsq8, msp[c * 32 + j]);
                                        if (v > thr) list_append(v, lin0 + c * 32 + j, ql, sw, glist, thr);
                                    }
                                    __syncwarp();
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&cm.sempty[h][b]); mbar_arrive(&cm.kempty[st]); }
                    }
                    if (sweep == 0) {
                        // 32 slot maxima per (slice, query): 16 from each tile parity
                        float* dst = p.candA + ((size_t)q * S1 + slice) * (2 * NSLOT) + b * NSLOT;
#pragma unroll
                        for (int u = 0; u < NSLOT; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(slot[u], slot[u + 1], slot[u + 2], slot[u + 3]);
                    }
                }
            }
            __syncthreads();
            if (sweep == 1 && threadIdx.x < QPAIR) p.lcnt[((size_t)(pair * QPAIR + threadIdx.x)) * S1 + slice] = sw.cnt[threadIdx.x];
            q_loaded = true;
            ++sweeps_done;

            // ---------------------------------------------------------------- merge A: tau_lo = k-th largest slot maximum
            if (sweep == 0) {
                cta_group_barrier(pair_ctr, ++pair_uses * S1, 40);
                if (warp >= 2) {
                    for (int ql = slice + ww * S1; ql < nqh * TQ; ql += NWORK * S1) {
                        const int q = pair * QPAIR + ql;
                        float* out = p.tau_lo;
                        if (q >= p.hw) { if (lane == 0) out[q] = INFINITY; continue; }      // padded queries never select anything
                        uint32_t v[MAX_SLICE1];
                        const float* src = p.candA + (size_t)q * S1 * 32;
#pragma unroll
                        for (int u = 0; u < MAX_SLICE1; ++u) v[u] = (u < S1) ? f2ord(__ldcg(src + u * 32 + lane)) : 0u;
                        uint32_t t = 0u;
#pragma unroll 1
                        for (int bit = 31; bit >= 0; --bit) {
                            const uint32_t trial = t | (1u << bit);
                            int c = 0;
#pragma unroll
                            for (int u = 0; u < MAX_SLICE1; ++u) c += (v[u] >= trial) ? 1 : 0;
                            c = __reduce_add_sync(0xffffffffu, c);
                            if (c >= p.top_k) t = trial;
                        }
                        // fewer than top_k finite maxima: t stays below the image of -inf -> collect everything
                        const float kth = (t <= f2ord(-INFINITY)) ? -INFINITY : ord2f(t);
                        if (lane == 0) out[q] = kth;
                    }
                }
                if (p.mode & MODE_SWEEP_B) cta_group_barrier(pair_ctr, ++pair_uses * S1, 41);
            } else {
                cta_group_barrier(pair_ctr, ++pair_uses * S1, 42);
            }
        }
    }

    // ------------------------------------------------------------------ merge B: exact top-k, weights, usage, final lists
    if ((p.mode & (MODE_SELECT | MODE_EXPORT32)) && in_sweep && warp >= 2) {
        uint64_t* scr = reinterpret_cast<uint64_t*>(&sw.st[0]) + (size_t)ww * SCR_CAP;
        for (int ql = slice + ww * S1; ql < QPAIR; ql += NWORK * S1) {
            const int q = pair * QPAIR + ql;
            if (q >= p.hw || ql >= nqh * TQ) {
                if (p.mode & MODE_SELECT) p.fin[(size_t)q * LISTK + lane] = make_uint2(0xffffffffu, 0u);
                if ((p.mode & MODE_EXPORT32) && q < p.hw_pad) p.top32_out[(size_t)q * LISTK + lane] = -INFINITY;
                continue;
            }
            const bool ext = (p.mode & MODE_EXT_TAU) != 0;
            const uint32_t min_ord = (ext && (p.mode & MODE_SELECT)) ? f2ord(p.tau_ext[q]) : 0u;
            int n = warp_gather_lists(p, q, scr, LISTK, min_ord, lane);
            __syncwarp();
            const int want = (p.mode & MODE_EXPORT32) ? LISTK : (ext ? LISTK : p.top_k);
            if (n > want) {
                const uint64_t kth = warp_kth_key(scr, n, want, lane);
                __syncwarp();
                n = warp_compact_ge(scr, n, kth, lane);
            }
            __syncwarp();
            uint64_t key = (lane < n) ? scr[lane] : 0ull;
            key = warp_sort_desc(key, lane);
            const bool have = key != 0ull;
            const float s = have ? ord2f(static_cast<uint32_t>(key >> 32)) : -INFINITY;
            const uint32_t lin = 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
            if (p.mode & MODE_EXPORT32) p.top32_out[(size_t)q * LISTK + lane] = s;
            if (p.mode & MODE_SELECT) {
                const float ex = have ? fast_exp(s) : 0.f;
                float inv;
                if (ext) {
                    inv = p.inv_ext[q];
                } else {
                    float den = ex;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
                    inv = 1.f / den;
                }
                const float pv = ex * inv;
                p.fin[(size_t)q * LISTK + lane] = have ? make_uint2(lin, __float_as_uint(pv)) : make_uint2(0xffffffffu, 0u);
                if (have && p.do_usage) {
                    int sgi, col;
                    unlinear_col(sg, (int)lin, sgi, col);
                    if (sg.usage[sgi]) atomicAdd(sg.usage[sgi] + col, pv);
                }
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ readout: O[q, c] += P[q, n] V[n, c]
    if (p.mode & MODE_READOUT) {
        fence_proxy_async_smem();                             // the merge scratch (generic proxy) aliases the TMA / UMMA buffers below
        cta_group_barrier(p.ctr, ++grid_uses * G, 43);
        pdl_launch_dependents();
        const int KT = sg.t64[sg.nseg];
        // zero both P buffers once; afterwards only the listed entries are written and cleared again
        if (warp >= 2) {
            uint4* pz = reinterpret_cast<uint4*>(&rd.p[0][0][0]);
            for (int i = threadIdx.x - 64; i < (int)(sizeof(rd.p) / 16); i += NWORK * 32) pz[i] = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async_smem();
        }
        __syncthreads();
        int g0 = 0;                                           // running k-tile count of this CTA (mbarrier phases)
        int item_iter = 0;
        for (int w = cta; w < p.n_items; w += G, ++item_iter) {
            // work item -> (row, slice); row -> (query pair, object, channel half)
            int row, sl, nsl;
            if (p.n_rows <= MAX_ROWS_TABLE) {
                row = 0;
                while (cm.row_item0[row + 1] <= w) ++row;
                sl = w - cm.row_item0[row]; nsl = p.row_slices[row];
            } else { row = w; sl = 0; nsl = 1; }
            const int chalf = row & 1;
            const int obj = (row >> 1) % p.n_obj;
            const int rpair = (row >> 1) / p.n_obj;
            const int rnqh = (rpair * 2 + 1 < p.qtiles) ? 2 : 1;
            const int k0 = (int)(((long long)KT * sl) / nsl), k1 = (int)(((long long)KT * (sl + 1)) / nsl);
            const int nkt = k1 - k0;

            if (warp == 0) {
                if (lane == 0) {
                    for (int n = 0; n < nkt; ++n) {
                        const int kt = k0 + n;
                        int s = 0;
                        if (sg.nseg > 1 && kt >= sg.t64[1]) s = 1;
                        if (sg.nseg > 2 && kt >= sg.t64[2]) s = 2;
                        const int col = sg.origin[s] + (kt - sg.t64[s]) * TK;
                        const int g = g0 + n, st = g % P2_VSTAGES, ph = (g / P2_VSTAGES) & 1;
                        mbar_wait(&cm.vempty[st], ph ^ 1, 7);
                        mbar_expect_tx(&cm.vfull[st], 256 * 128);
                        tma_load_3d(rd.v[st], &maps.v[sg.bank[s]], &cm.vfull[st], col, chalf * 256, p.obj_begin + obj);
                    }
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    constexpr uint32_t idesc_o = make_idesc_f16(TQ, 256);
                    if (item_iter > 0) { mbar_wait(&cm.oempty, (item_iter - 1) & 1, 8); tc_fence_after(); }
                    for (int n = 0; n < nkt; ++n) {
                        const int g = g0 + n, st = g % P2_VSTAGES, ph = (g / P2_VSTAGES) & 1;
                        const int pb = g & 1, pph = (g >> 1) & 1;
                        mbar_wait(&cm.vfull[st], ph, 9);
                        mbar_wait(&cm.pfull[pb], pph, 10);
                        tc_fence_after();
                        for (int h = 0; h < rnqh; ++h)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint64_t a = make_desc_sw128(smem_u32(rd.p[pb][h]) + j * 32);
                                const uint64_t bd = make_desc_sw128(smem_u32(rd.v[st]) + j * 32);
                                mma_f16_ss(tmem + h * 256, a, bd, idesc_o, (n > 0 || j > 0) ? 1u : 0u);
                            }
                        mma_commit(&cm.vempty[st]);
                        mma_commit(&cm.pempty[pb]);
                    }
                    mma_commit(&cm.ofull);
                }
            } else {
                // workers: bucket this CTA's (column, weight) entries by k-tile, then warps 0/1 build the P tiles
                const int wt = threadIdx.x - 64;                // 0..511
                int prev_b = 0, prev_e = 0;                     // builder: entries currently set in its P buffer
                for (int kb = k0; kb < k1 || kb == k0; kb += P2_MAXKT) {
                    const int ke = min(k1, kb + P2_MAXKT), nb = ke - kb;
                    if (nb <= 0) break;
                    asm volatile("bar.sync 1, 512;" ::: "memory");     // previous batch's tables are no longer read
                    for (int i = wt; i <= nb; i += NWORK * 32) rd.cur[i] = 0u;
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    const uint2* fin = p.fin + (size_t)(rpair * QPAIR) * LISTK;
                    for (int i = wt; i < rnqh * TQ * LISTK; i += NWORK * 32) {
                        const uint2 e = __ldcg(fin + i);
                        const int kt = (e.x == 0xffffffffu) ? -1 : (int)(e.x >> 6);
                        if (kt >= kb && kt < ke) atomicAdd(&rd.cur[kt - kb + 1], 1u);
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (ww == 0) {                               // exclusive scan of the counts (one warp)
                        uint32_t carry = 0u;
                        for (int b0 = 0; b0 <= nb; b0 += 32) {
                            const int i = b0 + lane;
                            uint32_t x = (i <= nb) ? rd.cur[i] : 0u;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                            if (i <= nb) { rd.cur[i] = x + carry; rd.off[i] = (uint16_t)(x + carry); }
                            carry += __shfl_sync(0xffffffffu, x, 31);
                        }
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    // cur[i] (i <= nb) now = number of entries in k-tiles < i ... shifted: counts were stored at index kt+1,
                    // so cur[kt] = first entry of k-tile kt and cur[kt + 1] = one past its last
                    for (int i = wt; i < rnqh * TQ * LISTK; i += NWORK * 32) {
                        const uint2 e = __ldcg(fin + i);
                        const int kt = (e.x == 0xffffffffu) ? -1 : (int)(e.x >> 6);
                        if (kt >= kb && kt < ke) {
                            const uint32_t pos = atomicAdd(&rd.cur[kt - kb], 1u);
                            const int ql = i / LISTK;            // query inside the pair
                            const uint32_t where = ((uint32_t)(ql >> 7) << 13) | ((uint32_t)(ql & 127) << 6) | (e.x & 63u);
                            const __half wv = __float2half_rn(__uint_as_float(e.y));
                            rd.ent[pos] = where | ((uint32_t)__half_as_ushort(wv) << 16);
                        }
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (ww < 2) {
                        const int pb = ww;
                        for (int n = kb - k0; n < ke - k0; ++n) {
                            const int g = g0 + n;
                            if ((g & 1) != pb) continue;
                            const int use = g >> 1;
                            mbar_wait(&cm.pempty[pb], (use & 1) ^ 1, 11);
                            for (int e = prev_b + lane; e < prev_e; e += 32) {      // clear what the previous use of this buffer set
                                const uint32_t x = rd.ent[e] & 0xffffu;
                                const uint32_t r = (x >> 6) & 127u, c = x & 63u;
                                *reinterpret_cast<__half*>(&rd.p[pb][x >> 13][r * 128 + (((c >> 3) ^ (r & 7u)) << 4) + (c & 7u) * 2]) = __float2half(0.f);
                            }
                            const int li = (k0 + n) - kb;
                            const int eb = rd.off[li], ee = rd.off[li + 1];
                            for (int e = eb + lane; e < ee; e += 32) {
                                const uint32_t x = rd.ent[e];
                                const uint32_t r = (x >> 6) & 127u, c = x & 63u;
                                *reinterpret_cast<uint16_t*>(&rd.p[pb][(x >> 13) & 1u][r * 128 + (((c >> 3) ^ (r & 7u)) << 4) + (c & 7u) * 2]) = (uint16_t)(x >> 16);
                            }
                            prev_b = eb; prev_e = ee;
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&cm.pfull[pb]);
                        }
                        // end of the batch: the entry table is about to be rebuilt -> clear this buffer's last entries now
                        if (prev_e > prev_b) {
                            // the last use of buffer pb in this batch
                            int last_g = -1;
                            for (int n = ke - k0 - 1; n >= kb - k0; --n) if (((g0 + n) & 1) == pb) { last_g = g0 + n; break; }
                            if (last_g >= 0) {
                                mbar_wait(&cm.pempty[pb], (last_g >> 1) & 1, 12);
                                for (int e = prev_b + lane; e < prev_e; e += 32) {
                                    const uint32_t x = rd.ent[e] & 0xffffu;
                                    const uint32_t r = (x >> 6) & 127u, c = x & 63u;
                                    *reinterpret_cast<__half*>(&rd.p[pb][x >> 13][r * 128 + (((c >> 3) ^ (r & 7u)) << 4) + (c & 7u) * 2]) = __float2half(0.f);
                                }
                                fence_proxy_async_smem();
                            }
                            prev_b = prev_e = 0;
                        }
                    }
                }
                // epilogue: O (2 x 128 lanes x 256 columns fp32) -> this item's partial tile
                mbar_wait(&cm.ofull, item_iter & 1, 13);
                tc_fence_after();
                {
                    const int quad = warp & 3, part = ww >> 2;    // part: q-tile (part >> 1), 128-column half (part & 1)
                    const int h = part >> 1, chh = part & 1;
                    if (h < rnqh) {
                        float* dst = p.partial + (size_t)w * (QPAIR * 256) + (size_t)(h * TQ + quad * 32 + lane) * 256 + chh * 128;
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            uint32_t r[32];
                            if (nkt > 0) {
                                tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + h * 256 + chh * 128 + c * 32, r);
                                tmem_ld_wait();
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) r[j] = 0u;
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                __stcg(reinterpret_cast<uint4*>(dst + c * 32 + j), make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cm.oempty);
                }
            }
            g0 += nkt;
        }

        // ---------------------------------------------------------------- reduce the column-slice partials
        cta_group_barrier(p.ctr, ++grid_uses * G, 44);
        if (warp >= 2) {
            float* tr = reinterpret_cast<float*>(&rd.v[0][0]) + (size_t)ww * 32 * 33;
            const int qblocks = p.hw_pad / 32;
            const int tiles = p.n_obj * qblocks * 16;
            for (int t = cta * NWORK + ww; t < tiles; t += G * NWORK) {
                const int obj = t / (qblocks * 16);
                const int rem = t - obj * qblocks * 16;
                const int qb = rem >> 4, cb = rem & 15;
                const int q0 = qb * 32, c0 = cb * 32;
                const int rpair = q0 / QPAIR, ql0 = q0 - rpair * QPAIR;
                const int chalf = cb >> 3;
                const int row = ((rpair * p.n_obj + obj) << 1) | chalf;
                int item0, nsl;
                if (p.n_rows <= MAX_ROWS_TABLE) { item0 = cm.row_item0[row]; nsl = p.row_slices[row]; } else { item0 = row; nsl = 1; }
                const float* src = p.partial + (size_t)item0 * (QPAIR * 256) + (size_t)ql0 * 256 + (c0 & 255) + lane;
                const int oo = p.obj_begin + obj;
#pragma unroll 4
                for (int qq = 0; qq < 32; ++qq) {
                    float acc = 0.f;
                    for (int s = 0; s < nsl; ++s) acc += __ldcg(src + (size_t)s * (QPAIR * 256) + qq * 256);
                    const int q = q0 + qq;
                    if (p.out_hwc && q < p.hw) p.out_hwc[((size_t)oo * p.hw + q) * XM_CV + c0 + lane] = __float2half_rn(acc);
                    tr[qq * 33 + lane] = acc;
                }
                __syncwarp();
                if (p.out_chw || p.out_f32) {
                    const int q = q0 + lane;
#pragma unroll 4
                    for (int cc = 0; cc < 32; ++cc) {
                        const float v = tr[lane * 33 + cc];
                        if (p.out_chw && q < p.hw) p.out_chw[((size_t)oo * XM_CV + c0 + cc) * p.hw + q] = __float2half_rn(v);
                        if (p.out_f32) p.out_f32[((size_t)obj * XM_CV + c0 + cc) * p.hw_pad + q] = v;
                    }
                }
                __syncwarp();
            }
        }
    } else {
        pdl_launch_dependents();
    }

    // ------------------------------------------------------------------ teardown: last CTA out re-arms the counters
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned old = atomicAdd(p.ctr + 1, 1u);
        if (old == (unsigned)G - 1u) {
            for (int i = 0; i < NCTR; ++i) p.ctr[i] = 0u;
            __threadfence();
        }
    }
}

// k-th largest over per-rank 32-entry records (T-shard merge): one warp per query, lane l owns record l (n_ranks <= 32).
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
    uint32_t t = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int u = 0; u < LISTK; ++u) c += (v[u] >= trial) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
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

// fp32 [n_obj][512][hw_pad] -> fp16 CHW / NHWC (T-shard, after the all-reduce)
__global__ void k1_cast(const float* __restrict__ src, int n_obj, int hw, int hw_pad, __half* __restrict__ out_chw, __half* __restrict__ out_hwc) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float tile[32][33];
    const int o = blockIdx.z;
    const int c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, q = q0 + tx;
        const float acc = (q < hw) ? src[((size_t)o * XM_CV + c) * hw_pad + q] : 0.f;
        tile[r][tx] = acc;
        if (out_chw && q < hw) out_chw[((size_t)o * XM_CV + c) * hw + q] = __float2half_rn(acc);
    }
    if (out_hwc) {
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int q = q0 + r, c = c0 + tx;
            if (q < hw) out_hwc[((size_t)o * hw + q) * XM_CV + c] = __float2half_rn(tile[tx][r]);
        }
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

static const int K1_PLAN_BYTES = 4096;          // XM_MAX_GROUPS column-range tables at the head of the workspace
static_assert(sizeof(K1Seg) * XM_MAX_GROUPS <= K1_PLAN_BYTES, "plan area too small");
static_assert(sizeof(K1Seg) % 4 == 0 && sizeof(K1Seg) / 4 <= NTHREADS, "K1Seg is copied by one thread per word");

namespace {
// launch geometry: depends only on (hw, #objects of the group, #SMs) -> identical for every frame of a recorded CUDA graph
struct K1Geom {
    int qtiles, qpairs, qrows;       // qrows = qpairs * 256 (list arrays are padded to whole pairs)
    int grid, nslice1;
};
K1Geom k1_geom(int hw) {
    K1Geom g;
    const int hw_pad = (hw + TQ - 1) / TQ * TQ;
    g.qtiles = hw_pad / TQ;
    g.qpairs = (g.qtiles + 1) / 2;
    g.qrows = g.qpairs * QPAIR;
    g.grid = xm_num_sms();
    g.nslice1 = g.grid / g.qpairs;
    if (g.nslice1 < 1) g.nslice1 = 1;
    if (g.nslice1 > MAX_SLICE1) g.nslice1 = MAX_SLICE1;
    return g;
}
struct K1Ws {
    K1Seg* plan; unsigned* ctr; float* candA; float* tau_lo; uint2* lists; int* lcnt; uint2* fin; float* partial;
    size_t total;
};
K1Ws k1_carve(void* workspace, int hw, int n_obj_total) {
    const K1Geom g = k1_geom(hw);
    uint8_t* ws = (uint8_t*)workspace;
    size_t off = 0;
    K1Ws w;
    auto take = [&](size_t bytes) { uint8_t* ptr = ws ? ws + off : nullptr; off += align_up(bytes, 256); return ptr; };
    w.plan = (K1Seg*)take(K1_PLAN_BYTES);
    w.ctr = (unsigned*)take(NCTR * 4);
    w.candA = (float*)take((size_t)g.qrows * g.nslice1 * 32 * 4);
    w.tau_lo = (float*)take((size_t)g.qrows * 4);
    w.lists = (uint2*)take((size_t)g.qrows * g.nslice1 * LCAP * 8);
    w.lcnt = (int*)take((size_t)g.qrows * g.nslice1 * 4);
    w.fin = (uint2*)take((size_t)g.qrows * LISTK * 8);
    const size_t rows_max = (size_t)g.qpairs * (n_obj_total > 0 ? n_obj_total : 1) * 2;
    const size_t items_max = rows_max > (size_t)g.grid ? rows_max : (size_t)g.grid;
    w.partial = (float*)take(items_max * QPAIR * 256 * 4);
    w.total = off;
    return w;
}
}  // namespace

extern "C" int64_t xm_affinity_workspace_bytes(int32_t hw, int32_t n_obj_total) {
    return (int64_t)k1_carve(nullptr, hw, n_obj_total).total;
}

// Build the per-group column-range tables (host side).  plan_out receives XM_MAX_GROUPS K1Seg records.
static int k1_build_plan(const xm_affinity_args_t* a, K1Seg* plan, bool allow_small = false) {
    XM_REQUIRE(a->n_groups > 0 && a->n_groups <= XM_MAX_GROUPS, "xm_affinity: bad n_groups %d", a->n_groups);
    for (int g = 0; g < a->n_groups; ++g) {
        const xm_group_t& gr = a->groups[g];
        XM_REQUIRE(gr.n_obj > 0 && gr.obj_begin >= 0 && gr.obj_begin + gr.n_obj <= a->n_obj_total, "xm_affinity: bad group %d objects", g);
        K1Seg& sg = plan[g];
        memset(&sg, 0, sizeof(sg));
        int tiles128 = 0, tiles64 = 0, cols = 0;
        for (int i = 0; i < 3; ++i) {
            const xm_bank_t& bk = a->banks[i];
            if (bk.size <= 0 || !bk.keys) continue;
            XM_REQUIRE(bk.cap % 8 == 0 && bk.size <= bk.cap, "xm_affinity: bank %d cap must be a multiple of 8 and >= size", i);
            XM_REQUIRE(bk.shrinkage && bk.values && bk.n_obj_cap > 0, "xm_affinity: bank %d has null shrinkage/values", i);
            const int begin = gr.begin[i];
            XM_REQUIRE(begin >= 0 && begin <= bk.size, "xm_affinity: group %d bank %d begin %d outside [0,%d]", g, i, begin, bk.size);
            if (begin == bk.size) continue;
            XM_REQUIRE(gr.obj_begin + gr.n_obj <= bk.n_obj_cap, "xm_affinity: bank %d holds fewer value planes than group %d needs", i, g);
            const int s = sg.nseg++;
            sg.bank[s] = i; sg.begin[s] = begin; sg.end[s] = bk.size; sg.cap[s] = (int)bk.cap;
            sg.t128[s] = tiles128; sg.t64[s] = tiles64; sg.col0[s] = cols;
            sg.origin[s] = begin & ~7;
            sg.shr[s] = bk.shrinkage; sg.usage[s] = (g == 0) ? bk.usage : nullptr;
            tiles128 += (bk.size - sg.origin[s] + TN - 1) / TN;
            tiles64 += (bk.size - sg.origin[s] + TK - 1) / TK;
            cols += bk.size - begin;
        }
        for (int s = sg.nseg; s < 4; ++s) { sg.t128[s] = tiles128; sg.t64[s] = tiles64; }
        XM_REQUIRE(allow_small || cols >= a->top_k, "xm_affinity: group %d sees %d memory columns < top_k=%d (torch.topk would raise)", g,
                   cols, a->top_k);
        XM_REQUIRE((long long)tiles64 * TK < (1ll << 31), "xm_affinity: too many memory columns");
    }
    return XM_OK;
}

extern "C" int xm_affinity_plan(const xm_affinity_args_t* a, void* host_plan_out, int64_t bytes) {
    XM_REQUIRE(a && host_plan_out && bytes >= K1_PLAN_BYTES, "xm_affinity_plan: need a %d-byte host buffer", K1_PLAN_BYTES);
    memset(host_plan_out, 0, K1_PLAN_BYTES);
    return k1_build_plan(a, (K1Seg*)host_plan_out);
}

namespace {
int k1_make_maps(const xm_affinity_args_t* a, K1Maps& maps) {
    {
        uint64_t d[2] = {KP, (uint64_t)a->hw_pad};
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
        uint32_t bv[3] = {TK, 256, 1};
        if (xm_make_tmap_f16(&maps.v[i], bk.values, 3, dv, sv, bv)) return XM_ERR_CUDA;
    }
    return XM_OK;
}

// readout work items: row = (query pair, object, channel half); rows of whole pairs weigh 2, the trailing half pair 1; the
// grid's CTAs are handed out proportionally (largest remainder), every row gets at least one
void k1_plan_items(const K1Geom& g, int n_obj, K1Params& p) {
    p.n_rows = g.qpairs * n_obj * 2;
    if (p.n_rows > MAX_ROWS_TABLE || p.n_rows >= g.grid) {
        p.n_items = p.n_rows;
        if (p.n_rows <= MAX_ROWS_TABLE) for (int r = 0; r < p.n_rows; ++r) p.row_slices[r] = 1;
        return;
    }
    int weight[MAX_ROWS_TABLE]; int wsum = 0;
    for (int r = 0; r < p.n_rows; ++r) {
        const int rpair = (r >> 1) / n_obj;
        weight[r] = (rpair * 2 + 1 < g.qtiles) ? 2 : 1;
        wsum += weight[r];
    }
    int given = 0;
    for (int r = 0; r < p.n_rows; ++r) {
        int s = (int)((long long)g.grid * weight[r] / wsum);
        if (s < 1) s = 1;
        if (s > 255) s = 255;
        p.row_slices[r] = (uint8_t)s; given += s;
    }
    // hand the remaining CTAs to the rows with the largest load per slice
    while (given < g.grid) {
        int best = 0; double best_load = -1.0;
        for (int r = 0; r < p.n_rows; ++r) {
            const double load = (double)weight[r] / p.row_slices[r];
            if (load > best_load && p.row_slices[r] < 255) { best_load = load; best = r; }
        }
        ++p.row_slices[best]; ++given;
    }
    while (given > g.grid) {               // (only if the floor of 1 per row overshot)
        int best = -1; double best_load = 1e30;
        for (int r = 0; r < p.n_rows; ++r) {
            if (p.row_slices[r] <= 1) continue;
            const double load = (double)weight[r] / (p.row_slices[r] - 1);
            if (load < best_load) { best_load = load; best = r; }
        }
        if (best < 0) break;
        --p.row_slices[best]; --given;
    }
    p.n_items = given;
}

int k1_launch(const xm_affinity_args_t* a, const K1Maps& maps, const K1Ws& w, int group, int mode, const float* tau_ext,
              const float* inv_ext, float* top32_out, float* out_f32, bool write_fp16, cudaStream_t stream) {
    const K1Geom g = k1_geom(a->hw);
    const xm_group_t& gr = a->groups[group];
    XM_REQUIRE(g.qpairs + 8 <= NCTR && g.qpairs <= g.grid, "xm_affinity: hw = %d needs more query pairs (%d) than CTAs (%d)", a->hw, g.qpairs, g.grid);
    K1Params p;
    memset(&p, 0, sizeof(p));
    p.seg = w.plan + group; p.bsq = a->bsq;
    p.hw = a->hw; p.hw_pad = a->hw_pad; p.top_k = a->top_k;
    p.obj_begin = gr.obj_begin; p.n_obj = gr.n_obj;
    p.do_usage = (group == 0) ? 1 : 0;
    p.mode = mode;
    p.qtiles = g.qtiles; p.qpairs = g.qpairs; p.nslice1 = g.nslice1;
    k1_plan_items(g, gr.n_obj, p);
    p.ctr = w.ctr; p.candA = w.candA; p.tau_lo = w.tau_lo; p.lists = w.lists; p.lcnt = w.lcnt; p.fin = w.fin; p.partial = w.partial;
    p.tau_ext = tau_ext; p.inv_ext = inv_ext; p.top32_out = top32_out; p.out_f32 = out_f32;
    p.out_chw = write_fp16 ? (__half*)a->readout_chw : nullptr;
    p.out_hwc = write_fp16 ? (__half*)a->readout_hwc : nullptr;
    p.out_obj_total = a->n_obj_total;
    p.dbg = (group == 0 && (mode & MODE_SWEEP_A)) ? a->debug_scores : nullptr;
    tc5_debug_init();
    static XmPerDevice attr_token = {0};
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(k1_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    }
    XM_CHECK_CUDA(tc5_launch(k1_fused, dim3(g.grid), dim3(NTHREADS), SMEM_TOTAL, stream, maps, p));
    XM_CHECK_CUDA(cudaGetLastError());
    xm_count_launches(1);
    return XM_OK;
}

int k1_common_checks(const xm_affinity_args_t* a, const char* who) {
    XM_REQUIRE(a, "%s: null args", who);
    XM_REQUIRE(a->hw > 0 && a->hw_pad == (a->hw + TQ - 1) / TQ * TQ, "%s: hw_pad must be hw rounded up to 128", who);
    XM_REQUIRE(a->top_k > 0 && a->top_k <= XM_MAX_TOPK, "%s: top_k must be in [1,%d]", who, XM_MAX_TOPK);
    XM_REQUIRE(a->n_groups > 0 && a->n_groups <= XM_MAX_GROUPS, "%s: bad n_groups %d", who, a->n_groups);
    XM_REQUIRE(a->qp && a->bsq && a->workspace, "%s: null query/workspace", who);
    XM_REQUIRE(a->workspace_bytes >= xm_affinity_workspace_bytes(a->hw, a->n_obj_total), "%s: workspace too small", who);
    return XM_OK;
}
}  // namespace

extern "C" int xm_affinity_readout(const xm_affinity_args_t* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = k1_common_checks(a, "xm_affinity_readout");
    if (rc != XM_OK) return rc;
    XM_REQUIRE(a->readout_chw || a->readout_hwc, "xm_affinity_readout: no output buffer");
    const K1Ws w = k1_carve(a->workspace, a->hw, a->n_obj_total);
    if (!a->plan_is_resident) {
        // eager convenience path: build the table here and copy it (pageable source: staged before the call returns)
        K1Seg plan[XM_MAX_GROUPS];
        memset(plan, 0, sizeof(plan));
        rc = k1_build_plan(a, plan);
        if (rc != XM_OK) return rc;
        XM_CHECK_CUDA(cudaMemcpyAsync(w.plan, plan, sizeof(K1Seg) * a->n_groups, cudaMemcpyHostToDevice, stream));
    }
    K1Maps maps;
    rc = k1_make_maps(a, maps);
    if (rc != XM_OK) return rc;
    for (int g = 0; g < a->n_groups; ++g) {
        const xm_group_t& gr = a->groups[g];
        XM_REQUIRE(gr.n_obj > 0 && gr.obj_begin >= 0 && gr.obj_begin + gr.n_obj <= a->n_obj_total, "xm_affinity_readout: bad group %d objects", g);
        rc = k1_launch(a, maps, w, g, MODE_FULL, nullptr, nullptr, nullptr, nullptr, true, stream);
        if (rc != XM_OK) return rc;
    }
    return XM_OK;
}

// One-time zeroing of the barrier counters of a fresh workspace (the kernel re-arms them itself on exit).
extern "C" int xm_affinity_workspace_init(void* workspace, int64_t workspace_bytes, int32_t hw, int32_t n_obj_total, void* stream) {
    XM_REQUIRE(workspace && workspace_bytes >= xm_affinity_workspace_bytes(hw, n_obj_total), "xm_affinity_workspace_init: workspace too small");
    const K1Ws w = k1_carve(workspace, hw, n_obj_total);
    XM_CHECK_CUDA(cudaMemsetAsync(w.ctr, 0, NCTR * 4, (cudaStream_t)stream));
    return XM_OK;
}

// ---------------------------------------------------------------------------------------------
// T-sharded memory read: the banks of ONE long video are distributed over R ranks by stored frame.
// Every rank holds the same query; the host interleaves three small collectives between these stages (the same kernel,
// restricted to some of its phases):
//   stage_a : sweep A + merge A -> local lower bound                      ... all_reduce(MAX)  tau_lo[hw_pad]
//   stage_b : sweep B with the global bound -> local 32 largest per query ... all_gather       top32[R][hw_pad][32]
//   merge   : exact global tau, 1/den from the gathered records (every rank, identical result)
//   stage_c : local entries >= tau weighted with the GLOBAL 1/den, readout -> fp32 partial ... all_reduce(SUM) readout_f32
//   cast    : fp32 -> fp16 CHW / NHWC
// Single object group per call (groups[0]); usage stays local to the rank that owns the column.
// ---------------------------------------------------------------------------------------------
static int tshard_setup(const xm_affinity_args_t* a, cudaStream_t stream, K1Maps& maps, K1Ws& w, bool upload_plan) {
    int rc = k1_common_checks(a, "xm_affinity_tshard");
    if (rc != XM_OK) return rc;
    XM_REQUIRE(a->n_groups == 1, "xm_affinity_tshard: exactly one object group per call");
    w = k1_carve(a->workspace, a->hw, a->n_obj_total);
    if (upload_plan) {
        K1Seg plan[XM_MAX_GROUPS];
        memset(plan, 0, sizeof(plan));
        rc = k1_build_plan(a, plan, /*allow_small=*/true);
        if (rc != XM_OK) return rc;
        XM_CHECK_CUDA(cudaMemcpyAsync(w.plan, plan, sizeof(K1Seg), cudaMemcpyHostToDevice, stream));
    }
    return k1_make_maps(a, maps);
}

extern "C" int xm_affinity_tshard_stage_a(const xm_affinity_args_t* a, float* tau_lo_local, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    K1Maps maps; K1Ws w;
    int rc = tshard_setup(a, stream, maps, w, true);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau_lo_local, "xm_affinity_tshard_stage_a: null output");
    rc = k1_launch(a, maps, w, 0, MODE_SWEEP_A, nullptr, nullptr, nullptr, nullptr, false, stream);
    if (rc != XM_OK) return rc;
    XM_CHECK_CUDA(cudaMemcpyAsync(tau_lo_local, w.tau_lo, (size_t)a->hw_pad * 4, cudaMemcpyDeviceToDevice, stream));
    return XM_OK;
}

extern "C" int xm_affinity_tshard_stage_b(const xm_affinity_args_t* a, const float* tau_lo_global, float* top32_local, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    K1Maps maps; K1Ws w;
    int rc = tshard_setup(a, stream, maps, w, false);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau_lo_global && top32_local, "xm_affinity_tshard_stage_b: null pointer");
    return k1_launch(a, maps, w, 0, MODE_SWEEP_B | MODE_EXPORT32 | MODE_EXT_TAU, tau_lo_global, nullptr, top32_local, nullptr, false, stream);
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
    K1Maps maps; K1Ws w;
    int rc = tshard_setup(a, stream, maps, w, false);
    if (rc != XM_OK) return rc;
    XM_REQUIRE(tau && inv_den && readout_f32, "xm_affinity_tshard_stage_c: null pointer");
    return k1_launch(a, maps, w, 0, MODE_SELECT | MODE_READOUT | MODE_EXT_TAU, tau, inv_den, nullptr, readout_f32, false, stream);
}

extern "C" int xm_affinity_tshard_cast(const float* readout_f32, int32_t n_obj, int32_t hw, int32_t hw_pad, void* readout_chw,
                                       void* readout_hwc, void* stream_) {
    XM_REQUIRE(readout_f32 && (readout_chw || readout_hwc) && n_obj >= 1, "xm_affinity_tshard_cast: bad arguments");
    XM_CHECK_CUDA(tc5_launch(k1_cast, dim3((hw + 31) / 32, XM_CV / 32, n_obj), dim3(32, 8), 0, (cudaStream_t)stream_, readout_f32, n_obj, hw,
                             hw_pad, (__half*)readout_chw, (__half*)readout_hwc));
    xm_count_launches(1);
    return XM_OK;
}
