// consolidate.cu — long-term memory maintenance on the device: prototype selection, consolidation
// ("memory potentiation"), least-used eviction and in-arena compaction.
//
// Reference: inference/memory_manager.py:316-390 (compress_features / consolidation), inference/kv_memory_store.py:125-181
// (sieve_by_range, remove_obsolete_features), model/memory_util.py:7-39,55-60 (similarity, full softmax with max subtraction).
// These run once every (max_mid_term_frames - min_mid_term_frames) memory frames on a few thousand columns (1.3 GFLOP at 480p),
// so they are plain fp32 CUDA-core kernels: exact IEEE arithmetic where the reference's selection depends on it (usage ratio,
// thresholds), deterministic reductions (fixed order, no float atomics), no host synchronisation except the one count the
// host-side bookkeeping needs after an eviction.
#include <cfloat>
#include <cmath>
#include "common.h"
#include "tc5.cuh"

using tc5::pdl_wait;
using tc5::pdl_launch_dependents;

namespace {

constexpr int SEL_THREADS = 1024;
constexpr int MAXK = 1024;

__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

__device__ __forceinline__ float usage_ratio(const float* use, const float* life, int i) { return __fdiv_rn(use[i], life[i]); }   // kv_memory_store.py:183-189

// block-wide sum of an int (all threads get the result); `red` holds 32 ints
__device__ __forceinline__ int block_sum(int v, int* red) {
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0;
    if (threadIdx.x < 32) { t = __reduce_add_sync(0xffffffffu, t); if (threadIdx.x == 0) red[0] = t; }
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------------------------------------------
// top-k of use/life (torch.topk(usage, k, sorted=True), memory_manager.py:355): indices by descending ratio, ties by ascending
// index.  One block: exact k-th largest of the 64-bit keys (ratio image << 32 | ~index) by bisection, then a bitonic sort.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS, 1)
usage_topk_kernel(const float* __restrict__ use, const float* __restrict__ life, int n, int k, int* __restrict__ out_idx) {
    __shared__ int red[32];
    __shared__ unsigned long long keys[MAXK];
    __shared__ int cursor;
    pdl_wait();
    pdl_launch_dependents();
    unsigned long long t = 0ull;
    for (int bit = 63; bit >= 0; --bit) {
        const unsigned long long trial = t | (1ull << bit);
        int c = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = ((unsigned long long)f2ord(usage_ratio(use, life, i)) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
            c += (key >= trial) ? 1 : 0;
        }
        if (block_sum(c, red) >= k) t = trial;
    }
    if (threadIdx.x == 0) cursor = 0;
    for (int i = threadIdx.x; i < MAXK; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = ((unsigned long long)f2ord(usage_ratio(use, life, i)) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
        if (key >= t) { const int p = atomicAdd(&cursor, 1); if (p < MAXK) keys[p] = key; }
    }
    __syncthreads();
    // bitonic sort, descending, MAXK entries (zeros sink to the end)
    for (int kk = 2; kk <= MAXK; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const int i = threadIdx.x, l = i ^ j;
            if (l > i) {
                const unsigned long long a = keys[i], b = keys[l];
                const bool desc = ((i & kk) == 0);
                if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[l] = a; }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < k; i += blockDim.x) out_idx[i] = (int)(0xffffffffu - (uint32_t)(keys[i] & 0xffffffffull));
}

// ------------------------------------------------------------------------------------------------------------------
// least-used eviction list (kv_memory_store.py:160-181): thr = the n_remove-th smallest ratio, survivors have ratio > thr.
// keep_idx receives the surviving columns in ascending order, *count their number.  One block.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS, 1)
usage_evict_kernel(const float* __restrict__ use, const float* __restrict__ life, int n, int n_remove, int* __restrict__ keep_idx,
                   int* __restrict__ count) {
    __shared__ int red[32];
    __shared__ int base;
    pdl_wait();
    pdl_launch_dependents();
    // smallest t with count(ratio <= t) >= n_remove  ==  the n_remove-th smallest ratio (bisection on the ordered image)
    uint32_t lo_img = 0u;          // largest image with count(<= img) < n_remove, built bit by bit
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = lo_img | (1u << bit);
        int c = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) c += (f2ord(usage_ratio(use, life, i)) < trial) ? 1 : 0;
        if (block_sum(c, red) < n_remove) lo_img = trial;     // fewer than n_remove values lie strictly below `trial`
    }
    const uint32_t thr = lo_img;   // the n_remove-th smallest value's image: count(< thr) < n_remove <= count(<= thr)
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool keep = (i < n) && f2ord(usage_ratio(use, life, i)) > thr;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        const int wsum = __popc(b);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wsum;
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const int v = red[w]; if (w < (int)(threadIdx.x >> 5)) pre += v; tot += v; }
        if (keep) keep_idx[base + pre + __popc(b & ((1u << (threadIdx.x & 31)) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// ------------------------------------------------------------------------------------------------------------------
// consolidation affinity: one block per prototype q.  S[n] = -(sum_c e_q[c] (k_n[c] - k_q[c])^2) * s[n] / 8 over the candidates
// n in [col_begin, n) (the anisotropic L2 of memory_util.py:7-39 in its un-expanded form), full softmax over n with max
// subtraction (memory_util.py:55-60), and the shrinkage read-out sum_n s[n] aff[n] (memory_manager.py:388).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
consolidate_affinity_kernel(const __half* __restrict__ kp, const float* __restrict__ s, const __half* __restrict__ e, int n,
                            const int* __restrict__ proto_idx, int col_begin, float* __restrict__ aff, long long aff_stride,
                            float* __restrict__ shr_out) {
    __shared__ float kq[XM_CK], eq[XM_CK];
    __shared__ float redf[32];
    pdl_wait();
    pdl_launch_dependents();
    const int q = blockIdx.x;
    const int pidx = proto_idx[q];
    if (pidx < col_begin) return;                     // prototype outside this group's columns (memory_manager.py:357-359)
    if (threadIdx.x < XM_CK) {
        kq[threadIdx.x] = __half2float(kp[(size_t)pidx * 128 + XM_CK + threadIdx.x]);
        eq[threadIdx.x] = e ? __half2float(e[(size_t)pidx * XM_CK + threadIdx.x]) : 0.f;
    }
    __syncthreads();
    float* row = aff + (size_t)q * aff_stride;
    auto block_reduce = [&](float v, bool is_max) -> float {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, u) : v + u; }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = v;
        __syncthreads();
        float t = (threadIdx.x < (blockDim.x >> 5)) ? redf[threadIdx.x] : (is_max ? -INFINITY : 0.f);
        if (threadIdx.x < 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, t, o); t = is_max ? fmaxf(t, u) : t + u; }
            if (threadIdx.x == 0) redf[0] = t;
        }
        __syncthreads();
        return redf[0];
    };
    float m = -INFINITY;
    const bool has_e = e != nullptr;
    for (int i = col_begin + threadIdx.x; i < n; i += blockDim.x) {
        const uint4* kr = reinterpret_cast<const uint4*>(kp + (size_t)i * 128 + XM_CK);
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint4 w = __ldg(kr + u);
            const __half2* h2 = reinterpret_cast<const __half2*>(&w);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float2 f = __half22float2(h2[v]);
                const float q0 = kq[u * 8 + 2 * v], q1 = kq[u * 8 + 2 * v + 1];
                if (has_e) {            // -sum e (k_n - k_q)^2  ==  -a_sq + 2ab - b_sq  (memory_util.py:22-27)
                    const float d0 = f.x - q0, d1 = f.y - q1;
                    acc = fmaf(eq[u * 8 + 2 * v], d0 * d0, acc);
                    acc = fmaf(eq[u * 8 + 2 * v + 1], d1 * d1, acc);
                } else {                // no selection: -|k_n|^2 + 2 k_n.k_q, WITHOUT the |k_q|^2 term (memory_util.py:28-32)
                    acc = fmaf(f.x, f.x - 2.f * q0, acc);
                    acc = fmaf(f.y, f.y - 2.f * q1, acc);
                }
            }
        }
        const float sv = -acc * s[i] * 0.125f;
        row[i] = sv;
        m = fmaxf(m, sv);
    }
    m = block_reduce(m, true);
    float sum = 0.f;
    for (int i = col_begin + threadIdx.x; i < n; i += blockDim.x) { const float ex = expf(row[i] - m); row[i] = ex; sum += ex; }
    sum = block_reduce(sum, false);
    const float inv = __fdiv_rn(1.f, sum);
    float sh = 0.f;
    for (int i = col_begin + threadIdx.x; i < n; i += blockDim.x) { const float a = row[i] * inv; row[i] = a; sh = fmaf(s[i], a, sh); }
    for (int i = threadIdx.x; i < col_begin; i += blockDim.x) row[i] = 0.f;
    sh = block_reduce(sh, false);
    if (shr_out && threadIdx.x == 0) shr_out[q] = sh;
}

// ------------------------------------------------------------------------------------------------------------------
// prototype values: part[split][o][c][j] = sum over this split's candidates n of v[o][c][n] * aff[valid_q[j]][n]
// (memory_manager.py:382-386, `v @ affinity`).  Block = 32 channels x up to 128 prototypes, K chunks of 64 candidates in smem.
// ------------------------------------------------------------------------------------------------------------------
constexpr int PV_C = 32, PV_K = 64, PV_Q = 128;
__global__ void __launch_bounds__(256)
consolidate_values_kernel(const __half* __restrict__ v, long long cap, int col_begin, int n, const float* __restrict__ aff,
                          long long aff_stride, const int* __restrict__ valid_q, int n_valid, int per_split, float* __restrict__ part) {
    __shared__ float sv[PV_C][PV_K + 1];
    __shared__ float sa[PV_K][PV_Q + 1];
    pdl_wait();
    pdl_launch_dependents();
    const int c0 = blockIdx.x * PV_C, o = blockIdx.y, split = blockIdx.z;
    const int k_begin = col_begin + split * per_split, k_end = min(n, k_begin + per_split);
    const int tc = threadIdx.x & 7, tq = threadIdx.x >> 3;          // 4 channels x 4 prototypes per thread
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const __half* vbase = v + ((size_t)o * XM_CV + c0) * cap;
    for (int k0 = k_begin; k0 < k_end; k0 += PV_K) {
        for (int i = threadIdx.x; i < PV_C * PV_K; i += blockDim.x) {
            const int c = i / PV_K, kk = i % PV_K;
            sv[c][kk] = (k0 + kk < k_end) ? __half2float(vbase[(size_t)c * cap + k0 + kk]) : 0.f;
        }
        for (int i = threadIdx.x; i < PV_K * PV_Q; i += blockDim.x) {
            const int j = i / PV_K, kk = i % PV_K;                   // consecutive threads read consecutive candidates of a prototype row
            sa[kk][j] = (j < n_valid && k0 + kk < k_end) ? aff[(size_t)(valid_q ? valid_q[j] : j) * aff_stride + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < PV_K; ++kk) {
            float a4[4], b4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) a4[a] = sv[tc * 4 + a][kk];
#pragma unroll
            for (int b = 0; b < 4; ++b) b4[b] = sa[kk][tq * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
        }
        __syncthreads();
    }
    float* dst = part + (((size_t)split * gridDim.y + o) * XM_CV + c0) * n_valid;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = tq * 4 + b;
            if (j < n_valid) dst[(size_t)(tc * 4 + a) * n_valid + j] = acc[a][b];
        }
}
__global__ void consolidate_values_reduce_kernel(const float* __restrict__ part, int splits, size_t plane, __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp) acc += part[(size_t)sp * plane + i];
        out[i] = __float2half_rn(acc);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// in-arena compaction: column i (first <= i < m) of every array takes the content of column src(i) = keep_idx[i] (or i + shift),
// src(i) >= i.  Two passes through `tmp` (gather, then write back) so that overlapping moves are safe at any grid size.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int src_col(const int* keep_idx, int shift, int i) { return keep_idx ? keep_idx[i] : i + shift; }

__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, int row_words, const int* __restrict__ keep_idx, int shift, int first,
                                   int m, uint32_t* __restrict__ tmp) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t total = (size_t)(m - first) * row_words;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(t / row_words), w = (int)(t % row_words);
        tmp[t] = src[(size_t)src_col(keep_idx, shift, first + r) * row_words + w];
    }
}
__global__ void scatter_rows_kernel(uint32_t* __restrict__ dst, int row_words, int first, int m, const uint32_t* __restrict__ tmp) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t total = (size_t)(m - first) * row_words;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
        dst[(size_t)first * row_words + t] = tmp[t];
}
__global__ void gather_cols_kernel(const __half* __restrict__ v, long long cap, const int* __restrict__ keep_idx, int shift, int first, int m,
                                   __half* __restrict__ tmp) {
    pdl_wait();
    pdl_launch_dependents();
    const int mm = m - first;
    const size_t plane = blockIdx.y;                                  // (object, channel)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mm; i += gridDim.x * blockDim.x)
        tmp[plane * mm + i] = v[plane * cap + src_col(keep_idx, shift, first + i)];
}
__global__ void scatter_cols_kernel(__half* __restrict__ v, long long cap, int first, int m, const __half* __restrict__ tmp) {
    pdl_wait();
    pdl_launch_dependents();
    const int mm = m - first;
    const size_t plane = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mm; i += gridDim.x * blockDim.x)
        v[plane * cap + first + i] = tmp[plane * mm + i];
}

}  // namespace

#define STREAM ((cudaStream_t)stream)

extern "C" int xm_usage_topk(const float* use, const float* life, int32_t n, int32_t k, int32_t* out_idx, void* stream) {
    XM_REQUIRE(use && life && out_idx, "xm_usage_topk: null pointer");
    XM_REQUIRE(n >= 1 && k >= 1 && k <= n && k <= MAXK, "xm_usage_topk: need 1 <= k <= min(n, %d)", MAXK);
    XM_CHECK_CUDA(tc5_launch(usage_topk_kernel, dim3(1), dim3(SEL_THREADS), 0, STREAM, use, life, (int)n, (int)k, (int*)out_idx));
    xm_count_launches(1);
    return XM_OK;
}

extern "C" int xm_usage_evict_list(const float* use, const float* life, int32_t n, int32_t n_remove, int32_t* keep_idx, int32_t* count,
                                   void* stream) {
    XM_REQUIRE(use && life && keep_idx && count, "xm_usage_evict_list: null pointer");
    XM_REQUIRE(n >= 1 && n_remove >= 1 && n_remove <= n, "xm_usage_evict_list: need 1 <= n_remove <= n");
    XM_CHECK_CUDA(tc5_launch(usage_evict_kernel, dim3(1), dim3(SEL_THREADS), 0, STREAM, use, life, (int)n, (int)n_remove, (int*)keep_idx, (int*)count));
    xm_count_launches(1);
    return XM_OK;
}

extern "C" int xm_consolidate_affinity(const void* kp, const float* s, const void* e, int32_t n, const int32_t* proto_idx, int32_t n_proto,
                                       int32_t col_begin, float* aff, int64_t aff_stride, float* shr_out, void* stream) {
    XM_REQUIRE(kp && s && proto_idx && aff, "xm_consolidate_affinity: null pointer");
    XM_REQUIRE(n >= 1 && n_proto >= 1 && col_begin >= 0 && col_begin < n && aff_stride >= n, "xm_consolidate_affinity: bad sizes");
    XM_CHECK_CUDA(tc5_launch(consolidate_affinity_kernel, dim3(n_proto), dim3(256), 0, STREAM, (const __half*)kp, s, (const __half*)e, (int)n,
                             (const int*)proto_idx, (int)col_begin, aff, (long long)aff_stride, shr_out));
    xm_count_launches(1);
    return XM_OK;
}

extern "C" int64_t xm_consolidate_scratch_bytes(int32_t n_obj, int32_t n_valid) {
    return (int64_t)8 * n_obj * XM_CV * n_valid * 4;
}

extern "C" int xm_consolidate_values(const void* v, int64_t cap, int32_t n_obj, int32_t col_begin, int32_t n, const float* aff,
                                     int64_t aff_stride, const int32_t* valid_q, int32_t n_valid, float* scratch, int64_t scratch_bytes,
                                     void* out, void* stream) {
    XM_REQUIRE(v && aff && scratch && out, "xm_consolidate_values: null pointer");
    XM_REQUIRE(n_obj >= 1 && n_valid >= 1 && n_valid <= PV_Q && col_begin >= 0 && col_begin < n && cap >= n, "xm_consolidate_values: bad sizes");
    const int splits = 8;
    XM_REQUIRE(scratch_bytes >= xm_consolidate_scratch_bytes(n_obj, n_valid), "xm_consolidate_values: scratch too small");
    int per_split = ((n - col_begin) + splits - 1) / splits;
    per_split = (per_split + PV_K - 1) / PV_K * PV_K;
    XM_CHECK_CUDA(tc5_launch(consolidate_values_kernel, dim3(XM_CV / PV_C, n_obj, splits), dim3(256), 0, STREAM, (const __half*)v, (long long)cap,
                             (int)col_begin, (int)n, aff, (long long)aff_stride, (const int*)valid_q, (int)n_valid, per_split, scratch));
    const size_t plane = (size_t)n_obj * XM_CV * n_valid;
    XM_CHECK_CUDA(tc5_launch(consolidate_values_reduce_kernel, dim3((unsigned)((plane + 255) / 256)), dim3(256), 0, STREAM, (const float*)scratch,
                             splits, plane, (__half*)out));
    xm_count_launches(2);
    return XM_OK;
}

extern "C" int64_t xm_bank_compact_tmp_bytes(int32_t n_obj_cap, int32_t moved) {
    return (int64_t)moved * (128 * 2 + XM_CK * 2 + 3 * 4 + (int64_t)n_obj_cap * XM_CV * 2) + 1024;
}

extern "C" int xm_bank_compact(void* kp, float* s, void* e, float* use, float* life, void* v, int64_t cap, int32_t n_obj_cap,
                               const int32_t* keep_idx, int32_t shift, int32_t first, int32_t m, void* tmp, int64_t tmp_bytes, void* stream) {
    XM_REQUIRE(kp && s && e && use && life && v && tmp, "xm_bank_compact: null pointer");
    XM_REQUIRE(first >= 0 && m >= first && m <= cap && n_obj_cap >= 1 && (keep_idx || shift >= 0), "xm_bank_compact: bad range");
    const int mm = m - first;
    if (mm == 0) return XM_OK;
    XM_REQUIRE(tmp_bytes >= xm_bank_compact_tmp_bytes(n_obj_cap, mm), "xm_bank_compact: tmp too small");
    const int grid = xm_num_sms() * 4;
    struct Arr { void* p; int words; };
    const Arr arrs[5] = {{kp, 64}, {e, 32}, {s, 1}, {use, 1}, {life, 1}};
    for (int a = 0; a < 5; ++a) {     // one array at a time through the same scratch (stream ordered)
        XM_CHECK_CUDA(tc5_launch(gather_rows_kernel, dim3(grid), dim3(256), 0, STREAM, (const uint32_t*)arrs[a].p, arrs[a].words,
                                 (const int*)keep_idx, (int)shift, (int)first, (int)m, (uint32_t*)tmp));
        XM_CHECK_CUDA(tc5_launch(scatter_rows_kernel, dim3(grid), dim3(256), 0, STREAM, (uint32_t*)arrs[a].p, arrs[a].words, (int)first, (int)m,
                                 (const uint32_t*)tmp));
    }
    const int planes = n_obj_cap * XM_CV;
    XM_CHECK_CUDA(tc5_launch(gather_cols_kernel, dim3((mm + 255) / 256, planes), dim3(256), 0, STREAM, (const __half*)v, (long long)cap,
                             (const int*)keep_idx, (int)shift, (int)first, (int)m, (__half*)tmp));
    XM_CHECK_CUDA(tc5_launch(scatter_cols_kernel, dim3((mm + 255) / 256, planes), dim3(256), 0, STREAM, (__half*)v, (long long)cap, (int)first,
                             (int)m, (const __half*)tmp));
    xm_count_launches(12);
    return XM_OK;
}
