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
//             key tile, which halves the per-SM TMA ingest), every scan thread keeps 8 branch-free running maxima of
//             disjoint column subsets ("slot maxima").  Every key tile is swept (a sampled sweep is cheaper but its bound
//             collapses on real videos, see k1_launch).
//   merge A   tau_lo[q] = k-th largest slot maximum  (a LOWER bound of the true k-th largest score: the maxima belong
//             to distinct memory columns).
//   sweep B   the same S tiles again (bit-identical instruction stream); every score > pred(tau_lo) is appended with its
//             column index to a thread-private list (32 entries per (query, slice, tile parity); a full list keeps its 32
//             largest scores, which keeps the selection exact for any input).
//   merge B   exact top-k per query on 64-bit keys (score, lowest column first), weights exp(S)/sum exp(S) (no max
//             subtraction, memory_util.py:48-49), final (column, weight) list per query.
//   readout   O[q, c] += P[q, n] V[n, c]  as a dense tcgen05 contraction (the reference's dense v @ affinity):
//             CTA = 256 queries x 256 value channels x a slice of the memory columns, O in TMEM (2 x 256 columns),
//             V tiles by TMA, the P tile is zero except for the listed entries, which are scattered into the swizzled
//             smem operand (and cleared again after use).  No score is recomputed in this phase.  The warps that do not
//             build P tiles sum the usage (column sums of P, memory_util.py:62-63) in 2^-40 fixed point.
//   reduce    sum of the column-slice partials -> fp16 CHW / NHWC (or fp32 for the T-sharded mode).
// Math.  Packed operands Kp[n] = (k_n^2, k_n), Qp[q] = (-e_q, 2 k_q e_q) (fp16, 128 wide):
//     S'[q,n] = Qp[q] . Kp[n]                                (tcgen05.mma, fp32 accumulate)
//     S[q,n]  = (S'[q,n] - bsq[q]) * shrinkage[n] / sqrt(64)  (bsq = sum_c e k_q^2, fp32)
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdlib>
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
constexpr int NSLOT = 8;                // slot maxima per scan thread
constexpr int LCAP = 32;                // candidate capacity per scan thread = (column slice, tile parity, query); >= top_k keeps it exact
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
    int a_stride;            // sweep A looks at every a_stride-th key tile only (any subset of real scores gives a valid lower bound)
    int n_rows, n_items;
    unsigned* ctr;
    float* candA;            // [qpairs*256][nslice1][2 * NSLOT]
    float* tau_lo;           // [qpairs*256]
    uint2* lists;            // [qpairs*256][nslice1][LCAP]  (score bits, linear column)
    int* lcnt;               // [qpairs*256][nslice1]
    uint2* fin;              // [qpairs*256][32]             (linear column or -1, weight bits)
    float* partial;          // [n_items][256][256]
    unsigned long long* uacc; // [max_columns] fixed-point (2^-40) column sums of the affinity (usage), order independent
    int uacc_cols;
    const float* tau_ext;    // MODE_EXT_TAU
    const float* inv_ext;
    float* top32_out;        // MODE_EXPORT32
    float* out_f32;          // fp32 [n_obj][hw_pad][512] (T-shard) or null
    __half* out_chw;
    __half* out_hwc;
    int out_obj_total;
    float* dbg;
    unsigned long long* timeline;   // [grid][16] globaltimer stamps of the phase boundaries (diagnostics, always written)
    uint8_t row_slices[MAX_ROWS_TABLE];
};

__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * LOG2E));
    return y;
}
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));       // FMNMX3 (sm_100+)
    return r;
}
__device__ __forceinline__ float score2(uint32_t acc_bits, float bsq8, float ms) {
    // ((S' - b_sq) * shrinkage) / 8 == (S'/8 - b_sq/8) * shrinkage exactly (power-of-two scaling commutes with rounding)
    return fmaf(__uint_as_float(acc_bits), 0.125f, -bsq8) * ms;
}

// global -> shared bulk copy (1-D TMA) completing on an mbarrier; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }
#ifdef K1_TRACE
// cycle accounting of CTA 0 (diagnostics build): every role accumulates clock64() differences in registers and writes its totals
// once per sweep into the trace area behind the time stamps ([role slot][sweep][counter])
#define K1_CLK() clock64()
#define K1_ACC(var, t0) do { var += clock64() - (t0); } while (0)
#define K1_TRACE_OUT(slot, sweep, c0, c1, c2) do { if (blockIdx.x == 0) { unsigned long long* tb_ = p.timeline + 148 * 16 + ((slot) * 2 + (sweep)) * 4; \
    tb_[0] = (unsigned long long)(c0); tb_[1] = (unsigned long long)(c1); tb_[2] = (unsigned long long)(c2); } } while (0)
#else
#define K1_CLK() 0ll
#define K1_ACC(var, t0) do { } while (0)
#define K1_TRACE_OUT(slot, sweep, c0, c1, c2) do { } while (0)
#endif
#define K1_STAMP(i) do { if (threadIdx.x == 0) p.timeline[(size_t)blockIdx.x * 16 + (i)] = global_ns(); } while (0)
// the same from the first worker thread (warp 2, lane 0): stamps 11..15 = merge B done, readout tables built, MMAs done, tile
// drained, list entries counted
#define K1_WSTAMP(i) do { if (threadIdx.x == 64) p.timeline[(size_t)blockIdx.x * 16 + (i)] = global_ns(); } while (0)

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
};
struct ReadSmem {
    alignas(1024) uint8_t v[P2_VSTAGES][256 * 128];          // [stage] 256 channel rows x (64 columns = 128 B)
    alignas(1024) uint8_t p[2][2][TQ * 128];                 // [buffer][q-tile] 128 query rows x (64 columns = 128 B)
    float entw[P2_MAXENT];                                   // weights of this item's entries, bucketed by k-tile
    uint16_t entp[P2_MAXENT];                                // their position in the P tile: q-tile << 13 | row << 6 | column
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
        unsigned long long t0 = 0ull;
        while (ld_acquire_u32(ctr) < target) {
            __nanosleep(40);
            const unsigned long long now = global_ns();
            if (t0 == 0ull) t0 = now;
            if (now - t0 > 4000000000ull) mbar_timeout(tag, target);      // 4 s: a CTA never arrived
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

// Candidate lists are PRIVATE to a scan thread (query, column slice, tile parity): plain appends to global memory, no lock and
// no fence (nobody else reads them before the next inter-CTA barrier).  A full list (LCAP >= top_k entries) keeps its LCAP
// largest scores, which keeps the selection exact for any input; that path only runs when > 32 scores of ONE thread's
// ~500 columns pass the lower bound.
// (all LCAP loads of a scan are issued before the first compare: one L2 round trip per call, not LCAP)
__device__ __noinline__ float list_min(const uint2* lst) {
    float v[LCAP];
#pragma unroll
    for (int u = 0; u < LCAP; ++u) v[u] = __uint_as_float(__ldcg(&lst[u]).x);
    float m = INFINITY;
#pragma unroll
    for (int u = 0; u < LCAP; ++u) m = fminf(m, v[u]);
    return m;
}
__device__ __noinline__ float list_replace_min(uint2* lst, float sc, int lin) {
    float vv[LCAP];
#pragma unroll
    for (int u = 0; u < LCAP; ++u) vv[u] = __uint_as_float(__ldcg(&lst[u]).x);
    float m1 = INFINITY, m2 = INFINITY; int p1 = 0;
#pragma unroll
    for (int u = 0; u < LCAP; ++u) {
        const float v = vv[u];
        if (v < m1) { m2 = m1; m1 = v; p1 = u; } else if (v < m2) { m2 = v; }
    }
    if (sc > m1) {
        __stcg(&lst[p1], make_uint2(__float_as_uint(sc), (uint32_t)lin));
        m1 = fminf(m2, sc);
    }
    return m1;                                                // the smallest listed score: nothing below it matters any more
}
__device__ __forceinline__ void list_append(uint2* lst, int& cnt, float& thr, float sc, int lin) {
    if (cnt < LCAP) {
        __stcg(&lst[cnt], make_uint2(__float_as_uint(sc), (uint32_t)lin));
        if (++cnt == LCAP) thr = fmaxf(thr, list_min(lst));
    } else {
        thr = fmaxf(thr, list_replace_min(lst, sc, lin));
    }
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
// gather the candidate lists of query q (2 * nslice1 thread lists) into scr.  Fast path: every lane reads the counts of its
// lists at once and copies their entries to its own offset (one L2 round trip per step instead of one per list).  Lists that
// together exceed the scratch fall back to a sequential gather that first reduces the scratch to its `keep` largest keys
// (exact).  min_score_ord: entries below it are dropped (T-shard selection).
__device__ int warp_gather_lists(const K1Params& p, int q, uint64_t* scr, int keep, uint32_t min_score_ord, int lane) {
    const int nlists = 2 * p.nslice1;                         // (column slice, tile parity); <= 64
    const int* cntp = p.lcnt + (size_t)q * nlists;
    const uint2* lbase = p.lists + (size_t)q * nlists * LCAP;
    const int c0 = (lane < nlists) ? min(LCAP, __ldcg(cntp + lane)) : 0;
    const int c1 = (lane + 32 < nlists) ? min(LCAP, __ldcg(cntp + lane + 32)) : 0;
    int x0 = c0, x1 = c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
        if (lane >= o) { x0 += y0; x1 += y1; }
    }
    const int tot0 = __shfl_sync(0xffffffffu, x0, 31), total = tot0 + __shfl_sync(0xffffffffu, x1, 31);
    int n = 0;
    if (total <= SCR_CAP) {
        const int off0 = x0 - c0, off1 = tot0 + x1 - c1;
        // loads in batches of 8 per list (independent L2 requests in flight), then the shared-memory stores
        const uint2* l0 = lbase + (size_t)lane * LCAP;
        const uint2* l1 = lbase + (size_t)(lane + 32) * LCAP;
        for (int e0 = 0; e0 < max(c0, c1); e0 += 8) {
            uint2 a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] = (e0 + u < c0) ? __ldcg(l0 + e0 + u) : make_uint2(0u, 0u);
                b[u] = (e0 + u < c1) ? __ldcg(l1 + e0 + u) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (e0 + u < c0) scr[off0 + e0 + u] = make_key(a[u].x, a[u].y);
                if (e0 + u < c1) scr[off1 + e0 + u] = make_key(b[u].x, b[u].y);
            }
        }
        n = total;
        __syncwarp();
    } else {
        for (int s = 0; s < nlists; ++s) {
            const int c = min(LCAP, __ldcg(cntp + s));
            if (c == 0) continue;
            if (n + c > SCR_CAP && n > keep) {
                const uint64_t kth = warp_kth_key(scr, n, keep, lane);
                __syncwarp();
                n = warp_compact_ge(scr, n, kth, lane);
            }
            if (lane < c) { const uint2 ent = __ldcg(lbase + (size_t)s * LCAP + lane); scr[n + lane] = make_key(ent.x, ent.y); }
            n += c;
            __syncwarp();
        }
    }
    if (min_score_ord != 0u) n = warp_compact_ge(scr, n, static_cast<uint64_t>(min_score_ord) << 32, lane);
    return n;
}
// The `want` largest of scr[0..n) (n <= 128, distinct keys) in DESCENDING order into out[0..min(n, want)): every lane ranks its
// (up to 4) keys against all n by broadcast reads -- no cross-lane traffic, independent iterations.  Returns min(n, want).
__device__ int warp_select_sorted(const uint64_t* scr, int n, int want, uint64_t* out, int lane) {
    uint64_t mine[4];
    int rank[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int e = lane + 32 * i; mine[i] = (e < n) ? scr[e] : ~0ull; rank[i] = 0; }
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
        const uint64_t kj = scr[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) rank[i] += (kj > mine[i]) ? 1 : 0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) if (lane + 32 * i < n && rank[i] < want) out[rank[i]] = mine[i];
    __syncwarp();
    return min(n, want);
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
// score scan of one 128-column key tile by one thread (= one query): 8 blocks of 16 columns, the TMEM load of the next block
// in flight while the current one is processed.  Compile-time variants keep the hot loop small (instruction cache!):
//   SWEEP 0: slot maxima   SWEEP 1: collect scores above thr     FULL: every column of the tile is valid     DBG: test dump
// ---------------------------------------------------------------------------------------------
struct ScanCtx {
    float bsq8;
    float thr;
    uint2* glist;
    int cnt;
    int lin0;
    int lo, hi;
    float* dbg;            // column 0 of this tile, this query (DBG only)
    ptrdiff_t dbg_stride;
};

template <int SWEEP, bool FULL, bool DBG>
__device__ __forceinline__ void scan_block16(const uint32_t (&r)[16], const float* msp, int c, float (&slot)[NSLOT], ScanCtx& cx) {
    if (SWEEP == 0) {
        float sc[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 f = *reinterpret_cast<const float4*>(msp + c * 16 + j);
            sc[j] = score2(r[j], cx.bsq8, f.x); sc[j + 1] = score2(r[j + 1], cx.bsq8, f.y);
            sc[j + 2] = score2(r[j + 2], cx.bsq8, f.z); sc[j + 3] = score2(r[j + 3], cx.bsq8, f.w);
        }
        if (!FULL) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = c * 16 + j;
                sc[j] = (n >= cx.lo && n < cx.hi) ? sc[j] : -INFINITY;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) slot[j] = fmax3(slot[j], sc[j], sc[j + 8]);      // 8 slots, one FMNMX3 per two scores
        if (DBG) {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (sc[j] != -INFINITY) cx.dbg[(ptrdiff_t)(c * 16 + j) * cx.dbg_stride] = sc[j];
        }
    } else {
        // 3 instructions per score: t = acc/8 - bsq8, d = thr - t*ms (ONE fma, exact product), sign bit of d shifted into a mask.
        // sign(d) is set whenever the exact product exceeds thr -- a superset of fl(t*ms) > thr -- so the flagged scores are
        // recomputed with the reference expression (score2) and compared again before they are listed.
        uint32_t mask = 0u;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 f = *reinterpret_cast<const float4*>(msp + c * 16 + j);
            const float ms[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float t = fmaf(__uint_as_float(r[j + u]), 0.125f, -cx.bsq8);
                const float d = fmaf(-t, ms[u], cx.thr);
                mask = __funnelshift_l(__float_as_uint(d), mask, 1);       // score j ends up at bit 15 - j
            }
        }
        if (!FULL) {
            const int jlo = max(0, cx.lo - c * 16), jhi = min(16, cx.hi - c * 16);     // valid j in [jlo, jhi)
            const uint32_t valid = (jhi > jlo) ? (((1u << (jhi - jlo)) - 1u) << (16 - jhi)) : 0u;
            mask &= valid;
        }
        if (mask) {                                                         // ~4 % of the lanes
            const int lin = cx.lin0 + c * 16;
            do {
                const int bit = 31 - __clz(mask);
                mask &= ~(1u << bit);
                const int j = 15 - bit;
                // r[j] with a dynamic j: binary select tree (registers cannot be indexed)
                const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8;
                const uint32_t x0 = b0 ? r[1] : r[0], x1 = b0 ? r[3] : r[2], x2 = b0 ? r[5] : r[4], x3 = b0 ? r[7] : r[6];
                const uint32_t x4 = b0 ? r[9] : r[8], x5 = b0 ? r[11] : r[10], x6 = b0 ? r[13] : r[12], x7 = b0 ? r[15] : r[14];
                const uint32_t y0 = b1 ? x1 : x0, y1 = b1 ? x3 : x2, y2 = b1 ? x5 : x4, y3 = b1 ? x7 : x6;
                const uint32_t z0 = b2 ? y1 : y0, z1 = b2 ? y3 : y2;
                const uint32_t acc = b3 ? z1 : z0;
                const float v = score2(acc, cx.bsq8, msp[c * 16 + j]);
                if (v > cx.thr) list_append(cx.glist, cx.cnt, cx.thr, v, lin + j);
            } while (mask);
        }
        __syncwarp();
    }
}

template <int SWEEP, bool FULL, bool DBG>
__device__ __forceinline__ void scan_tile(uint32_t trow, const float* msp, float (&slot)[NSLOT], ScanCtx& cx) {
    uint32_t ra[16], rb[16];
    tmem_ld_32x32b_x16(trow, ra);
#pragma unroll 1
    for (int c = 0; c < 8; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + (c + 1) * 16, rb);
        scan_block16<SWEEP, FULL, DBG>(ra, msp, c, slot, cx);
        tmem_ld_wait();
        if (c + 2 < 8) tmem_ld_32x32b_x16(trow + (c + 2) * 16, ra);
        scan_block16<SWEEP, FULL, DBG>(rb, msp, c + 1, slot, cx);
    }
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
k1_fused(const __grid_constant__ K1Maps maps, const __grid_constant__ K1Params p) {
    extern __shared__ uint8_t smem_raw[];
    // aligned base as an OFFSET from the extern array: the compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    K1_STAMP(0);
    const bool want_usage = p.do_usage && (p.mode & MODE_READOUT) && (sg.usage[0] || sg.usage[1] || sg.usage[2]);
    if (want_usage) {                                         // consumed after the grid barrier that opens the readout
        const int ncol = sg.t64[sg.nseg] * TK;
        for (int i = cta * NTHREADS + threadIdx.x; i < ncol; i += G * NTHREADS) p.uacc[i] = 0ull;
    }

    unsigned pair_uses = 0, grid_uses = 0;

    // ------------------------------------------------------------------ sweeps (A: slot maxima, B: collect)
    const int total128 = sg.t128[sg.nseg];
    const int t_begin = (int)(((long long)total128 * slice) / S1);
    const int t_end = (int)(((long long)total128 * (slice + 1)) / S1);
    const int nt = in_sweep ? (t_end - t_begin) : 0;
    unsigned* pair_ctr = p.ctr + 8 + pair;
    int done_it = 0, done_uses[2] = {0, 0};
    bool q_loaded = false;

    for (int sweep = 0; sweep < 2; ++sweep) {
        const int this_mode = sweep == 0 ? MODE_SWEEP_A : MODE_SWEEP_B;
        if (!(p.mode & this_mode)) continue;
        if (in_sweep) {
            // sweep A may sample every a_stride-th key tile; iteration i of a sweep works on tile t_begin + i * tstride
            const int tstride = (sweep == 0) ? p.a_stride : 1;
            const int n_it = (nt + tstride - 1) / tstride;
            const int it0 = done_it;                          // mbarrier phases continue across the two sweeps
            const int uses0[2] = {done_uses[0], done_uses[1]};   // uses of S buffer b (iterations with (i & 1) == b) so far
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
                    long long c_wait = 0;
                    for (int i = 0; i < n_it; ++i) {
                        int s, col, lo, hi;
                        locate_tile128(sg, t_begin + i * tstride, s, col, lo, hi);
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const uint32_t ms_bytes = (uint32_t)min(TN, sg.cap[s] - col) * 4u;
                        const long long tw = K1_CLK();
                        mbar_wait(&cm.kempty[st], ph ^ 1, 2);
                        K1_ACC(c_wait, tw);
                        mbar_expect_tx(&cm.kfull[st], 2 * TN * 128 + ms_bytes);
                        const CUtensorMap* km = &maps.k[sg.bank[s]];
                        tma_load_2d(sw.st[st].k[0], km, &cm.kfull[st], 0, col);
                        tma_load_2d(sw.st[st].k[1], km, &cm.kfull[st], 64, col);
                        bulk_g2s(sw.st[st].ms, sg.shr[s] + col, ms_bytes, &cm.kfull[st]);
                    }
                    K1_TRACE_OUT(0, sweep, c_wait, n_it, 0);
                }
            } else if (warp == 1) {
                if (lane == 0) {
                    constexpr uint32_t idesc = make_idesc_f16(TQ, TN);
                    if (!q_loaded) mbar_wait(&cm.qfull, 0, 1);
                    long long c_kfull = 0, c_sempty = 0, c_issue = 0;
                    for (int i = 0; i < n_it; ++i) {
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const int b = i & 1, use = uses0[b] + (i >> 1);
                        long long tw = K1_CLK();
                        mbar_wait(&cm.kfull[st], ph, 3);
                        K1_ACC(c_kfull, tw);
                        for (int h = 0; h < nqh; ++h) {
                            tw = K1_CLK();
                            mbar_wait(&cm.sempty[h][b], (use & 1) ^ 1, 4);
                            K1_ACC(c_sempty, tw);
                            tw = K1_CLK();
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
                            K1_ACC(c_issue, tw);
                        }
                        mma_commit(&cm.kempty[st]);
                    }
                    K1_TRACE_OUT(1, sweep, c_kfull, c_sempty, c_issue);
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
                    ScanCtx cx;
                    cx.bsq8 = bsq8; cx.thr = -INFINITY; cx.cnt = 0; cx.dbg = nullptr; cx.dbg_stride = 0;
                    if (sweep == 1) {
                        const float t = (p.mode & MODE_EXT_TAU) ? p.tau_ext[q] : __ldcg(p.tau_lo + q);
                        cx.thr = (t == -INFINITY) ? -INFINITY : ((t == INFINITY) ? FLT_MAX : __uint_as_float(
                                  t > 0.f ? __float_as_uint(t) - 1u : (t < 0.f ? __float_as_uint(t) + 1u : 0x80000001u)));   // pred(tau_lo)
                    }
                    cx.glist = p.lists + (((size_t)q * S1 + slice) * 2 + b) * LCAP;          // this thread's candidate list
                    long long c_wait = 0, c_scan = 0;
                    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16) + (h * 2 + b) * TN;
                    for (int i = b; i < n_it; i += 2) {
                        int s, col, lo, hi;
                        locate_tile128(sg, t_begin + i * tstride, s, col, lo, hi);
                        const int use = uses0[b] + (i >> 1);
                        const int it = it0 + i, st = it % P1_KSTAGES, ph = (it / P1_KSTAGES) & 1;
                        const float* msp = sw.st[st].ms;
                        long long tw = K1_CLK();
                        mbar_wait(&cm.kfull[st], ph, 6);              // the shrinkage slice arrived with the key tile
                        mbar_wait(&cm.sfull[h][b], use & 1, 5);
                        K1_ACC(c_wait, tw);
                        tw = K1_CLK();
                        tc_fence_after();
                        const bool full_tile = (lo <= 0) && (hi >= TN);
                        const int lin0 = linear_col(sg, s, col);
                        cx.lin0 = lin0; cx.lo = lo; cx.hi = hi;
                        if (sweep == 0) {
                            if (p.dbg) {                       // tests / selector only (warp-uniform)
                                cx.dbg = p.dbg + ((ptrdiff_t)sg.col0[s] + (col - sg.begin[s])) * (ptrdiff_t)p.hw_pad + q;
                                cx.dbg_stride = p.hw_pad;
                                scan_tile<0, false, true>(trow, msp, slot, cx);
                            } else if (full_tile) {
                                scan_tile<0, true, false>(trow, msp, slot, cx);
                            } else {
                                scan_tile<0, false, false>(trow, msp, slot, cx);
                            }
                        } else {
                            if (full_tile) scan_tile<1, true, false>(trow, msp, slot, cx);
                            else scan_tile<1, false, false>(trow, msp, slot, cx);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(&cm.sempty[h][b]); mbar_arrive(&cm.kempty[st]); }
                        K1_ACC(c_scan, tw);
                    }
                    if (lane == 0) K1_TRACE_OUT(2 + ww, sweep, c_wait, c_scan, (n_it - b + 1) >> 1);
                    if (sweep == 0) {
                        // 32 slot maxima per (slice, query): 16 from each tile parity
                        float* dst = p.candA + ((size_t)q * S1 + slice) * (2 * NSLOT) + b * NSLOT;
#pragma unroll
                        for (int u = 0; u < NSLOT; u += 4) *reinterpret_cast<float4*>(dst + u) = make_float4(slot[u], slot[u + 1], slot[u + 2], slot[u + 3]);
                    } else {
                        p.lcnt[((size_t)q * S1 + slice) * 2 + b] = cx.cnt;
                    }
                }
            }
            __syncthreads();
            K1_STAMP(sweep == 0 ? 1 : 4);
            q_loaded = true;
            done_it += n_it; done_uses[0] += (n_it + 1) >> 1; done_uses[1] += n_it >> 1;

            // ---------------------------------------------------------------- merge A: tau_lo = k-th largest slot maximum
            if (sweep == 0) {
                cta_group_barrier(pair_ctr, ++pair_uses * S1, 40);
                K1_STAMP(2);
                if (warp >= 2) {
                    for (int ql = slice + ww * S1; ql < nqh * TQ; ql += NWORK * S1) {
                        const int q = pair * QPAIR + ql;
                        float* out = p.tau_lo;
                        if (q >= p.hw) { if (lane == 0) out[q] = INFINITY; continue; }      // padded queries never select anything
                        // every lane keeps the two largest of its S1 slot maxima; the k-th largest of those 64 values is
                        // still a lower bound of the k-th largest score (real scores of distinct columns) and ~3 ranks looser
                        const int nval = S1 * 2 * NSLOT;
                        const float* src = p.candA + (size_t)q * nval;
                        float m1 = -INFINITY, m2 = -INFINITY;
                        float cv[MAX_SLICE1 * 2 * NSLOT / 32];        // all loads in flight at once: one L2 round trip
#pragma unroll
                        for (int j = 0; j < MAX_SLICE1 * 2 * NSLOT / 32; ++j) cv[j] = (lane + 32 * j < nval) ? __ldcg(src + lane + 32 * j) : -INFINITY;
#pragma unroll
                        for (int j = 0; j < MAX_SLICE1 * 2 * NSLOT / 32; ++j) {
                            m2 = fmaxf(m2, fminf(m1, cv[j]));
                            m1 = fmaxf(m1, cv[j]);
                        }
                        const uint32_t o1 = f2ord(m1), o2 = f2ord(m2);
                        uint32_t t = 0u;
#pragma unroll 1
                        for (int bit = 31; bit >= 0; --bit) {
                            const uint32_t trial = t | (1u << bit);
                            int c = ((o1 >= trial) ? 1 : 0) + ((o2 >= trial) ? 1 : 0);
                            c = __reduce_add_sync(0xffffffffu, c);
                            if (c >= p.top_k) t = trial;
                        }
                        // fewer than top_k finite maxima: t stays below the image of -inf -> collect everything
                        const float kth = (t <= f2ord(-INFINITY)) ? -INFINITY : ord2f(t);
                        if (lane == 0) out[q] = kth;
                    }
                }
                if (p.mode & MODE_SWEEP_B) cta_group_barrier(pair_ctr, ++pair_uses * S1, 41);
                K1_STAMP(3);
            } else {
                cta_group_barrier(pair_ctr, ++pair_uses * S1, 42);
                K1_STAMP(5);
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
            long long tw = K1_CLK(), c_g = 0, c_k = 0, c_r = 0;
            int n = warp_gather_lists(p, q, scr, LISTK, min_ord, lane);
            __syncwarp();
            K1_ACC(c_g, tw); tw = K1_CLK();
            const int want = (p.mode & MODE_EXPORT32) ? LISTK : (ext ? LISTK : p.top_k);
            uint64_t key;
            if (n <= 128) {
                uint64_t* out = scr + SCR_CAP / 2;             // n <= 128 leaves the upper half of the scratch free
                n = warp_select_sorted(scr, n, want, out, lane);
                key = (lane < n) ? out[lane] : 0ull;
                K1_ACC(c_k, tw); tw = K1_CLK();
            } else {
                const uint64_t kth = warp_kth_key(scr, n, want, lane);
                __syncwarp();
                n = warp_compact_ge(scr, n, kth, lane);
                __syncwarp();
                K1_ACC(c_k, tw); tw = K1_CLK();
                key = (lane < n) ? scr[lane] : 0ull;
                key = warp_sort_desc(key, lane);
            }
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
            }
            __syncwarp();
            K1_ACC(c_r, tw);
            if (lane == 0) K1_TRACE_OUT(18 + ww, 0, c_g, c_k, c_r);
        }
        K1_WSTAMP(11);
    }

    // ------------------------------------------------------------------ readout: O[q, c] += P[q, n] V[n, c]
    if (p.mode & MODE_READOUT) {
        fence_proxy_async_smem();                             // the merge scratch (generic proxy) aliases the TMA / UMMA buffers below
        __syncthreads();
        K1_STAMP(6);
        // zero both P buffers once (the merge scratch they alias is dead after the barrier above); afterwards only the listed
        // entries are written and cleared again.  Done while waiting for the other CTAs.
        if (warp >= 2) {
            uint4* pz = reinterpret_cast<uint4*>(&rd.p[0][0][0]);
            for (int i = threadIdx.x - 64; i < (int)(sizeof(rd.p) / 16); i += NWORK * 32) pz[i] = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async_smem();
        }
        cta_group_barrier(p.ctr, ++grid_uses * G, 43);
        K1_STAMP(7);
        pdl_launch_dependents();
        const int KT = sg.t64[sg.nseg];
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
                    // N = 256 on purpose: an M128 x N128 SS MMA re-reads the 4 KB P slice per 64 clocks of math (128 B/clk of
                    // shared memory, the whole budget) -- splitting the value tile in two N = 128 halves was measured 1.5x slower
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
                // this pair's final (column, weight) lists: 256 queries x 32 entries = 16 per worker thread, fetched in ONE
                // batch of independent loads (a load-per-iteration loop around the shared-memory atomics below costs a
                // serial L2 round trip per entry, twice)
                constexpr int ENT_PER_THREAD = QPAIR * LISTK / (NWORK * 32);
                uint2 ent[ENT_PER_THREAD];
                {
                    const uint2* fin = p.fin + (size_t)(rpair * QPAIR) * LISTK;
#pragma unroll
                    for (int j = 0; j < ENT_PER_THREAD; ++j) {
                        const int i = wt + j * (NWORK * 32);
                        ent[j] = (i < rnqh * TQ * LISTK) ? __ldcg(fin + i) : make_uint2(0xffffffffu, 0u);
                    }
                }
                for (int kb = k0; kb < k1 || kb == k0; kb += P2_MAXKT) {
                    const int ke = min(k1, kb + P2_MAXKT), nb = ke - kb;
                    if (nb <= 0) break;
                    asm volatile("bar.sync 1, 512;" ::: "memory");     // previous batch's tables are no longer read
                    for (int i = wt; i <= nb; i += NWORK * 32) rd.cur[i] = 0u;
                    asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < ENT_PER_THREAD; ++j) {
                        const int kt = (ent[j].x == 0xffffffffu) ? -1 : (int)(ent[j].x >> 6);
                        if (kt >= kb && kt < ke) atomicAdd(&rd.cur[kt - kb + 1], 1u);
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (item_iter == 0 && kb == k0) K1_WSTAMP(15);
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
#pragma unroll
                    for (int j = 0; j < ENT_PER_THREAD; ++j) {
                        const uint2 e = ent[j];
                        const int kt = (e.x == 0xffffffffu) ? -1 : (int)(e.x >> 6);
                        if (kt >= kb && kt < ke) {
                            const uint32_t pos = atomicAdd(&rd.cur[kt - kb], 1u);
                            const int ql = (wt + j * (NWORK * 32)) / LISTK;            // query inside the pair
                            rd.entp[pos] = (uint16_t)(((uint32_t)(ql >> 7) << 13) | ((uint32_t)(ql & 127) << 6) | (e.x & 63u));
                            rd.entw[pos] = __uint_as_float(e.y);
                        }
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (item_iter == 0 && kb == k0) K1_WSTAMP(12);
                    if (ww < 2) {
                        const int pb = ww;
                        auto p_addr = [&](uint32_t x) -> uint16_t* {        // element (q-tile, row, column) of the swizzled P tile
                            const uint32_t r = (x >> 6) & 127u, c = x & 63u;
                            return reinterpret_cast<uint16_t*>(&rd.p[pb][(x >> 13) & 1u][r * 128 + (((c >> 3) ^ (r & 7u)) << 4) + (c & 7u) * 2]);
                        };
                        for (int n = kb - k0; n < ke - k0; ++n) {
                            const int g = g0 + n;
                            if ((g & 1) != pb) continue;
                            const int use = g >> 1;
                            const int li = (k0 + n) - kb;
                            const int eb = rd.off[li], ee = rd.off[li + 1];
                            mbar_wait(&cm.pempty[pb], (use & 1) ^ 1, 11);
                            for (int e = prev_b + lane; e < prev_e; e += 32) *p_addr(rd.entp[e]) = 0;      // clear the previous use of this buffer
                            for (int e = eb + lane; e < ee; e += 32) *p_addr(rd.entp[e]) = __half_as_ushort(__float2half_rn(rd.entw[e]));
                            prev_b = eb; prev_e = ee;
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&cm.pfull[pb]);
                        }
                        // end of the batch: the entry table is about to be rebuilt -> clear this buffer's last entries now
                        if (prev_e > prev_b) {
                            int last_g = -1;
                            for (int n = ke - k0 - 1; n >= kb - k0; --n) if (((g0 + n) & 1) == pb) { last_g = g0 + n; break; }
                            if (last_g >= 0) {
                                mbar_wait(&cm.pempty[pb], (last_g >> 1) & 1, 12);
                                for (int e = prev_b + lane; e < prev_e; e += 32) *p_addr(rd.entp[e]) = 0;
                                fence_proxy_async_smem();
                            }
                            prev_b = prev_e = 0;
                        }
                    }
                    else if (want_usage && chalf == 0 && obj == 0) {
                        // usage (one CTA per (query pair, k-tile)): the 14 warps that do not build P tiles share the k-tiles.  A
                        // lane owns columns lane and lane + 32 of the tile and walks the tile's (few) entries: column sums in
                        // 2^-40 fixed point, no shared-memory atomics, and integer adds commute, so the result does not depend on
                        // the order in which the entries were bucketed.  Off the P builders' critical path.
                        for (int n = kb - k0 + (ww - 2); n < ke - k0; n += NWORK - 2) {
                            const int li = (k0 + n) - kb;
                            const int eb = rd.off[li], ee = rd.off[li + 1];
                            unsigned long long a0 = 0ull, a1 = 0ull;
                            for (int e = eb; e < ee; ++e) {
                                const uint32_t c = rd.entp[e] & 63u;
                                const unsigned long long wq = (unsigned long long)(rd.entw[e] * 1099511627776.f);
                                if (c == (uint32_t)lane) a0 += wq;
                                if (c == (uint32_t)lane + 32u) a1 += wq;
                            }
                            unsigned long long* ua = p.uacc + (size_t)(k0 + n) * TK;
                            if (a0) atomicAdd(ua + lane, a0);
                            if (a1) atomicAdd(ua + lane + 32, a1);
                        }
                    }
                }
                // epilogue: O (2 x 128 lanes x 256 columns fp32) -> this item's partial tile
                mbar_wait(&cm.ofull, item_iter & 1, 13);
                tc_fence_after();
                K1_WSTAMP(13);
                {
                    const int quad = warp & 3, part = ww >> 2;    // part: q-tile (part >> 1), 128-column half (part & 1)
                    const int h = part >> 1, chh = part & 1;
                    if (h < rnqh) {
                        // partial tile layout = 512-byte UNITS: unit (((h*4+quad)*2+chh)*4+c)*8+j holds channels
                        // chh*128+c*32+j*4..+3 of the 32 queries of one TMEM lane quadrant, lane-major, so every warp store is
                        // one contiguous 512 bytes (a query-major tile made each 16-byte store its own cache line: ~14 us)
                        float4* dst = reinterpret_cast<float4*>(p.partial + (size_t)w * (QPAIR * 256)) +
                                      (size_t)(((h * 4 + quad) * 2 + chh) * 32) * 32 + lane;
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
                            for (int j = 0; j < 8; ++j)
                                __stcg(reinterpret_cast<uint4*>(dst + (c * 8 + j) * 32), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&cm.oempty);
                    K1_WSTAMP(14);
                }
            }
            g0 += nkt;
        }

        // ---------------------------------------------------------------- reduce the column-slice partials
        // One work item per CTA (the usual case): the CTAs of a ROW synchronise among themselves and share that row's
        // reduction, so rows that finish early do not wait for the slowest CTA of the grid.  Otherwise: grid barrier.
        __syncthreads();
        K1_STAMP(8);
        const bool row_sync = p.n_items <= G && p.n_rows <= MAX_ROWS_TABLE;
        int my_row = 0, my_sl = 0, my_nsl = 1;
        if (row_sync) {
            if (cta < p.n_items) {
                while (cm.row_item0[my_row + 1] <= cta) ++my_row;
                my_sl = cta - cm.row_item0[my_row]; my_nsl = p.row_slices[my_row];
                cta_group_barrier(p.ctr + 96 + my_row, my_nsl, 44);
            }
        } else {
            cta_group_barrier(p.ctr, ++grid_uses * G, 44);
        }
        K1_STAMP(9);
        if (warp >= 2 && (!row_sync || cta < p.n_items)) {
            auto reduce_row = [&](int row, int v0, int vstride) {
                const int chalf = row & 1;
                const int obj = (row >> 1) % p.n_obj;
                const int rpair = (row >> 1) / p.n_obj;
                const int rnqh = (rpair * 2 + 1 < p.qtiles) ? 2 : 1;
                int item0, nsl;
                if (p.n_rows <= MAX_ROWS_TABLE) { item0 = cm.row_item0[row]; nsl = p.row_slices[row]; } else { item0 = row; nsl = 1; }
                const float* pbase = p.partial + (size_t)item0 * (QPAIR * 256);
                // one warp per GROUP = 4 units = (32 queries of a lane quadrant) x (16 channels): a lane owns one query, sums
                // 4 float4 per slice (512 contiguous bytes per warp load) and writes every requested output layout
                for (int gi = v0; gi < rnqh * 64; gi += vstride) {
                    const int jh = gi & 1, c = (gi >> 1) & 3, chh = (gi >> 3) & 1, quad = (gi >> 4) & 3, h = gi >> 6;
                    const float4* src = reinterpret_cast<const float4*>(pbase) +
                                        (size_t)((((h * 4 + quad) * 2 + chh) * 4 + c) * 8 + jh * 4) * 32 + lane;
                    float4 acc[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                    for (int s = 0; s < nsl; ++s) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 f = __ldcg(src + (size_t)s * (QPAIR * 256 / 4) + j * 32);
                            acc[j].x += f.x; acc[j].y += f.y; acc[j].z += f.z; acc[j].w += f.w;
                        }
                    }
                    const int q = rpair * QPAIR + h * TQ + quad * 32 + lane;
                    const int c0 = chalf * 256 + chh * 128 + c * 32 + jh * 16;
                    if (p.out_hwc && q < p.hw) {
                        uint4* o = reinterpret_cast<uint4*>(p.out_hwc + ((size_t)(p.obj_begin + obj) * p.hw + q) * XM_CV + c0);
                        o[0] = make_uint4(pack_half2(acc[0].x, acc[0].y), pack_half2(acc[0].z, acc[0].w), pack_half2(acc[1].x, acc[1].y), pack_half2(acc[1].z, acc[1].w));
                        o[1] = make_uint4(pack_half2(acc[2].x, acc[2].y), pack_half2(acc[2].z, acc[2].w), pack_half2(acc[3].x, acc[3].y), pack_half2(acc[3].z, acc[3].w));
                    }
                    if (p.out_f32) {
                        float4* o = reinterpret_cast<float4*>(p.out_f32 + ((size_t)obj * p.hw_pad + q) * XM_CV + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = acc[j];
                    }
                    if (p.out_chw && q < p.hw) {       // reference layout [object][channel][query]: lanes are consecutive queries
                        __half* o = p.out_chw + ((size_t)(p.obj_begin + obj) * XM_CV + c0) * p.hw + q;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            o[(size_t)(4 * j + 0) * p.hw] = __float2half_rn(acc[j].x);
                            o[(size_t)(4 * j + 1) * p.hw] = __float2half_rn(acc[j].y);
                            o[(size_t)(4 * j + 2) * p.hw] = __float2half_rn(acc[j].z);
                            o[(size_t)(4 * j + 3) * p.hw] = __float2half_rn(acc[j].w);
                        }
                    }
                }
            };
            if (row_sync) {
                reduce_row(my_row, my_sl * NWORK + ww, my_nsl * NWORK);
            } else {
                for (int row = 0; row < p.n_rows; ++row) reduce_row(row, cta * NWORK + ww, G * NWORK);
            }
        }
        if (want_usage && row_sync) cta_group_barrier(p.ctr, ++grid_uses * G, 45);     // every k-tile's column sums are complete
        // (3) usage: use_count[column] += sum over the queries of the affinity (memory_util.py:62-63, kv_memory_store.py:96-103)
        if (want_usage) {
            const int ncol = sg.t64[sg.nseg] * TK;
            for (int lin = cta * NTHREADS + threadIdx.x; lin < ncol; lin += G * NTHREADS) {
                const unsigned long long u = __ldcg(p.uacc + lin);
                if (u == 0ull) continue;
                int sgi, col;
                unlinear_col(sg, lin, sgi, col);
                if (sg.usage[sgi] && col >= sg.begin[sgi] && col < sg.end[sgi]) sg.usage[sgi][col] += __ull2float_rn(u) * 9.094947017729282e-13f;   // 2^-40
            }
        }
    } else {
        pdl_launch_dependents();
    }

    // ------------------------------------------------------------------ teardown: last CTA out re-arms the counters
    tc_fence_before();
    __syncthreads();
    K1_STAMP(10);
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

// fp32 [n_obj][hw_pad][512] -> fp16 CHW / NHWC (T-shard, after the all-reduce)
__global__ void k1_cast(const float* __restrict__ src, int n_obj, int hw, int hw_pad, __half* __restrict__ out_chw, __half* __restrict__ out_hwc) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float tile[32][33];
    const int o = blockIdx.z;
    const int c0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int q = q0 + r, c = c0 + tx;
        const float acc = (q < hw) ? src[((size_t)o * hw_pad + q) * XM_CV + c] : 0.f;
        tile[r][tx] = acc;
        if (out_hwc && q < hw) out_hwc[((size_t)o * hw + q) * XM_CV + c] = __float2half_rn(acc);
    }
    if (out_chw) {
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r, q = q0 + tx;
            if (q < hw) out_chw[((size_t)o * XM_CV + c) * hw + q] = __float2half_rn(tile[tx][r]);
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
constexpr int K1_DEFAULT_COLUMNS = 1 << 20;      // memory columns the default workspace size can account usage for
struct K1Ws {
    K1Seg* plan; unsigned* ctr; unsigned long long* timeline; float* candA; float* tau_lo; uint2* lists; int* lcnt; uint2* fin; float* partial;
    unsigned long long* uacc; size_t uacc_cols;
    size_t total;
};
K1Ws k1_carve(void* workspace, int hw, int n_obj_total, int64_t workspace_bytes = -1) {
    const K1Geom g = k1_geom(hw);
    uint8_t* ws = (uint8_t*)workspace;
    size_t off = 0;
    K1Ws w;
    auto take = [&](size_t bytes) { uint8_t* ptr = ws ? ws + off : nullptr; off += align_up(bytes, 256); return ptr; };
    w.plan = (K1Seg*)take(K1_PLAN_BYTES);
    w.ctr = (unsigned*)take(NCTR * 4);
    w.timeline = (unsigned long long*)take((size_t)g.grid * 16 * 8 + 4 * 1024 * 8);       // stamps + diagnostics
    w.candA = (float*)take((size_t)g.qrows * g.nslice1 * 32 * 4);
    w.tau_lo = (float*)take((size_t)g.qrows * 4);
    w.lists = (uint2*)take((size_t)g.qrows * g.nslice1 * 2 * LCAP * 8);
    w.lcnt = (int*)take((size_t)g.qrows * g.nslice1 * 2 * 4);
    w.fin = (uint2*)take((size_t)g.qrows * LISTK * 8);
    const size_t rows_max = (size_t)g.qpairs * (n_obj_total > 0 ? n_obj_total : 1) * 2;
    const size_t items_max = rows_max > (size_t)g.grid ? rows_max : (size_t)g.grid;
    w.partial = (float*)take(items_max * QPAIR * 256 * 4);
    // usage accumulator: one 8-byte cell per (64-padded) memory column; everything the caller provides beyond the fixed part
    // (xm_affinity_workspace_bytes reserves room for K1_DEFAULT_COLUMNS columns)
    w.uacc = (unsigned long long*)(ws ? ws + off : nullptr);
    w.uacc_cols = workspace_bytes < 0 ? (size_t)K1_DEFAULT_COLUMNS : (workspace_bytes > (int64_t)off ? (size_t)(workspace_bytes - (int64_t)off) / 8 : 0);
    w.total = off + (size_t)K1_DEFAULT_COLUMNS * 8;
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
        // cost of a k-tile: a whole pair is bound by its two M128 x N256 MMAs (~0.56 us), the trailing half pair by the 32 KB
        // value tile every CTA has to pull through TMA whatever it multiplies it with (~0.5 us) -- measured, not 2 : 1
        weight[r] = (rpair * 2 + 1 < g.qtiles) ? 16 : 14;
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
    {
        int64_t cols = 0;
        for (int i = 0; i < 3; ++i) if (a->banks[i].keys && a->banks[i].size > 0) cols += a->banks[i].cap + 64;
        XM_REQUIRE((int64_t)w.uacc_cols >= cols, "xm_affinity: workspace accounts usage for %lld memory columns, the banks hold up to %lld "
                   "(allocate xm_affinity_workspace_bytes() + 8 bytes per extra column)", (long long)w.uacc_cols, (long long)cols);
    }
    K1Params p;
    memset(&p, 0, sizeof(p));
    p.seg = w.plan + group; p.bsq = a->bsq;
    p.hw = a->hw; p.hw_pad = a->hw_pad; p.top_k = a->top_k;
    p.obj_begin = gr.obj_begin; p.n_obj = gr.n_obj;
    p.do_usage = (group == 0) ? 1 : 0;
    p.mode = mode;
    p.qtiles = g.qtiles; p.qpairs = g.qpairs; p.nslice1 = g.nslice1;
    // Sweep A looks at EVERY key tile.  Sampling every 2nd tile is a valid bound and 9 us cheaper at the config-2 maximum, but on
    // real videos the best matches of a query sit in a few neighbouring tiles (the same spot of the latest frames); when those
    // are not sampled the bound is loose, thread lists overflow and sweep B falls off a cliff (measured in the clip: 160 us
    // instead of 18).  With every tile the bound leaves 31 candidates per query on average (max ~50) instead of 62 (max ~250).
    // XMEM_K1_A_STRIDE=2 restores the sampled sweep for experiments.
    static const int a_stride_env = [] { const char* e = getenv("XMEM_K1_A_STRIDE"); const int v = e ? atoi(e) : 1; return v >= 1 && v <= 4 ? v : 1; }();
    p.a_stride = a->debug_scores ? 1 : a_stride_env;        // the test dump wants every score
    k1_plan_items(g, gr.n_obj, p);
    p.uacc = w.uacc; p.uacc_cols = (int)(w.uacc_cols > 0x7fffffff ? 0x7fffffff : w.uacc_cols);
    p.ctr = w.ctr; p.timeline = w.timeline; p.candA = w.candA; p.tau_lo = w.tau_lo; p.lists = w.lists; p.lcnt = w.lcnt; p.fin = w.fin; p.partial = w.partial;
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
    const K1Ws w = k1_carve(a->workspace, a->hw, a->n_obj_total, a->workspace_bytes);
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

// diagnostics: copy the phase-boundary time stamps of the LAST launch on this workspace ([grid][16] uint64 ns) to the host
extern "C" int xm_affinity_debug_timeline(void* workspace, int32_t hw, int32_t n_obj_total, unsigned long long* host_out, int32_t max_ctas) {
    XM_REQUIRE(workspace && host_out, "xm_affinity_debug_timeline: null pointer");
    const K1Ws w = k1_carve(workspace, hw, n_obj_total);
    const int n = k1_geom(hw).grid < max_ctas ? k1_geom(hw).grid : max_ctas;
    XM_CHECK_CUDA(cudaMemcpy(host_out, w.timeline, (size_t)n * 16 * 8, cudaMemcpyDeviceToHost));
    return n;
}

// diagnostics: candidates collected per query by the last sweep B on this workspace (sum of the thread lists' counts)
extern "C" int xm_affinity_debug_counts(void* workspace, int32_t hw, int32_t n_obj_total, int32_t* host_out) {
    const K1Ws w = k1_carve(workspace, hw, n_obj_total);
    const K1Geom g = k1_geom(hw);
    const int nl = 2 * g.nslice1;
    int* tmp = (int*)malloc((size_t)g.qrows * nl * 4);
    if (!tmp) return -1;
    cudaError_t e = cudaMemcpy(tmp, w.lcnt, (size_t)g.qrows * nl * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) for (int q = 0; q < hw; ++q) { int t = 0; for (int l = 0; l < nl; ++l) t += tmp[(size_t)q * nl + l]; host_out[q] = t; }
    free(tmp);
    return e == cudaSuccess ? 0 : -2;
}

extern "C" int xm_affinity_debug_trace(void* workspace, int32_t hw, int32_t n_obj_total, unsigned long long* host_out) {
    const K1Ws w = k1_carve(workspace, hw, n_obj_total);
    XM_CHECK_CUDA(cudaMemcpy(host_out, w.timeline + 148 * 16, 4 * 1024 * 8, cudaMemcpyDeviceToHost));
    XM_CHECK_CUDA(cudaMemset(w.timeline + 148 * 16, 0, 4 * 1024 * 8));
    return 0;
}

// One-time zeroing of the barrier counters of a fresh workspace (the kernel re-arms them itself on exit).
extern "C" int xm_affinity_workspace_init(void* workspace, int64_t workspace_bytes, int32_t hw, int32_t n_obj_total, void* stream) {
    XM_REQUIRE(workspace && workspace_bytes >= xm_affinity_workspace_bytes(hw, n_obj_total), "xm_affinity_workspace_init: workspace too small");
    const K1Ws w = k1_carve(workspace, hw, n_obj_total);
    XM_CHECK_CUDA(cudaMemsetAsync(w.ctr, 0, NCTR * 4, (cudaStream_t)stream));
    XM_CHECK_CUDA(cudaMemsetAsync(w.timeline, 0, (size_t)k1_geom(hw).grid * 16 * 8 + 4 * 1024 * 8, (cudaStream_t)stream));
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
    w = k1_carve(a->workspace, a->hw, a->n_obj_total, a->workspace_bytes);
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
