// common.cu — error string, driver entry point for cuTensorMapEncodeTiled, device queries.
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include "common.h"

static thread_local char g_err[512] = "";

void xm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* xm_last_error(void) { return g_err; }
extern "C" int xm_version(void) { return 100; }

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<encode_tiled_fn>(p);
        }
    });
    return fn;
}

int xm_make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    encode_tiled_fn fn = get_encode();
    if (!fn) {
        xm_set_error("cuTensorMapEncodeTiled driver entry point not available (no CUDA driver?)");
        return XM_ERR_CUDA;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        xm_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u] base %p",
                     (int)r, rank, (unsigned long long)dims[0], rank > 1 ? (unsigned long long)dims[1] : 0ull,
                     rank > 2 ? (unsigned long long)dims[2] : 0ull, rank > 3 ? (unsigned long long)dims[3] : 0ull,
                     rank > 4 ? (unsigned long long)dims[4] : 0ull, box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                     rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, base);
        return XM_ERR_CUDA;
    }
    return 0;
}

int xm_num_sms() {
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & 63;
    if (n[slot] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[slot] = v > 0 ? v : 148;
    }
    return n[slot];
}

bool xm_first_use_on_device(XmPerDevice* token) {
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    if (token->done_mask & bit) return false;
    token->done_mask |= bit;
    return true;
}

// ---- trap diagnostics: a host-mapped page kernels write to just before __trap() (tc5.cuh) ----
static int* g_trap_host = nullptr;
static int* g_trap_dev = nullptr;
int* xm_debug_trap_device_ptr() {
    if (!g_trap_host) {
        if (cudaHostAlloc((void**)&g_trap_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
        for (int i = 0; i < 16; ++i) g_trap_host[i] = 0;
        if (cudaHostGetDevicePointer((void**)&g_trap_dev, g_trap_host, 0) != cudaSuccess) return nullptr;
    }
    return g_trap_dev;
}
// out[0..6] = {valid, tag, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, parity}; returns 1 if a trap was recorded
extern "C" int xm_debug_last_trap(int* out) {
    if (!g_trap_host) return 0;
    for (int i = 0; i < 7; ++i) out[i] = g_trap_host[i];
    return g_trap_host[0];
}

// ---- launch accounting (bench.py reports how many of this library's kernels ran in the timed region) ----
#include <atomic>
static std::atomic<long long> g_launches{0};
void xm_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long xm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void xm_add_launch_count(int n) { xm_count_launches(n); }   // kernels replayed from a recorded CUDA graph

// ---- programmatic dependent launch toggle (default on; XMEM_NO_PDL=1 or xm_set_pdl(0) turns it off) ----
#include <cstdlib>
static int g_pdl = -1;
int xm_pdl_enabled() {
    if (g_pdl < 0) { const char* e = getenv("XMEM_NO_PDL"); g_pdl = (e && e[0] == '1') ? 0 : 1; }
    return g_pdl;
}
extern "C" void xm_set_pdl(int on) { g_pdl = on ? 1 : 0; }
