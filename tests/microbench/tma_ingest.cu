// tma_ingest.cu — how many bytes per second can ONE SM pull from L2 through TMA, and how does that scale with the number of
// active SMs, CTAs per SM and 2-CTA multicast?  (NOTES.md: the convolutions look bound by ~65-70 GB/s per active SM;
// this settles whether small grids should be spread over more SMs (cluster split-K) and whether multicast helps at all.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ingest tma_ingest.cu -lcuda && ./tma_ingest
//
// Each CTA streams `iters` boxes of 128 rows x 128 B (16 KB, 128-byte swizzle — the convolution's A tile) from a buffer
// that fits in L2 (default 48 MB) into a `stages`-deep shared-memory ring; a consumer thread just waits for the box and
// releases the stage.  Reported: GB/s per active SM and in total, for several grid sizes; then the same with clusters of
// two CTAs where every box is fetched by one CTA and multicast to both (bytes counted once per RECEIVING CTA).
// Standalone (driver API only for the tensor map), never run in round 1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (++spins > (1u << 26)) __trap();          // never hang the box
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int BOX = 128 * 128;       // bytes
constexpr int MAX_STAGES = 8;

struct Smem {
    alignas(1024) uint8_t ring[MAX_STAGES][BOX];
    alignas(8) uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
};

// MODE 0: every CTA loads its own boxes.  MODE 1: clusters of 2, box i is fetched by CTA (i & 1) and multicast to both.
template <int MODE>
__global__ void __launch_bounds__(64) ingest(const __grid_constant__ CUtensorMap map, int rows_total, int iters, int stages) {
    extern __shared__ uint8_t raw[];
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const uint32_t rank = MODE ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], MODE ? 2 : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE) cluster_sync();
    const int nboxes = rows_total / 128;
    // a different start row per CTA so that the CTAs do not all hit the same L2 lines at the same time
    const int start = (int)((blockIdx.x * 7919u) % (unsigned)nboxes);
    if (threadIdx.x == 0) {                                   // producer
        for (int it = 0; it < iters; ++it) {
            const int st = it % stages, ph = (it / stages) & 1;
            mbar_wait(&sm.empty[st], ph ^ 1);
            mbar_expect_tx(&sm.full[st], BOX);
            const int row = ((start + it) % nboxes) * 128;
            if (!MODE) tma_load_2d(sm.ring[st], &map, &sm.full[st], 0, row);
            else if ((it & 1) == (int)rank) tma_load_2d_mc(sm.ring[st], &map, &sm.full[st], 0, row, (uint16_t)0x3);
        }
    } else if (threadIdx.x == 32) {                           // consumer
        for (int it = 0; it < iters; ++it) {
            const int st = it % stages, ph = (it / stages) & 1;
            mbar_wait(&sm.full[st], ph);
            if (!MODE) mbar_arrive(&sm.empty[st]);
            else { mbar_arrive_cluster(&sm.empty[st], 0); mbar_arrive_cluster(&sm.empty[st], 1); }   // the stage is rewritten in BOTH CTAs
        }
    }
    __syncthreads();
    if (MODE) cluster_sync();
}

static CUtensorMap make_map(void* base, uint64_t rows) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap m;
    cuuint64_t dims[2] = {64, rows}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((Fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

template <int MODE>
static float run(const CUtensorMap& map, int rows, int ctas, int stages, int iters, int extra_smem) {
    const int smem = (int)sizeof(Smem) + 1024 + extra_smem;           // extra_smem pads the CTA to force 1 CTA/SM
    CK(cudaFuncSetAttribute(ingest<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = MODE ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, ingest<MODE>, map, rows, iters, stages));     // warm L2
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; ++r) CK(cudaLaunchKernelEx(&cfg, ingest<MODE>, map, rows, iters, stages));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 5;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 48;
    const int iters = argc > 2 ? atoi(argv[2]) : 2000;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const uint64_t rows = mb * 1024 * 1024 / 128 / 128 * 128;
    void* buf; CK(cudaMalloc(&buf, rows * 128)); CK(cudaMemset(buf, 1, rows * 128));
    CUtensorMap map = make_map(buf, rows);
    printf("%s, %d SMs, buffer %zu MB, %d boxes of 16 KB per CTA\n", prop.name, sms, mb, iters);
    printf("%-10s %6s %7s %9s %12s %12s\n", "mode", "CTAs", "stages", "ms", "GB/s total", "GB/s per SM");
    const int grids[] = {14, 28, 56, 112, sms, 2 * sms};
    for (int stages : {3, 6}) {
        for (int g : grids) {
            const int per_sm = g > sms ? 2 : 1;
            const int pad = per_sm == 1 ? 100 * 1024 - stages * BOX : 0;          // > half of the SM's smem -> one CTA per SM
            const float ms = run<0>(map, (int)rows, g, stages, iters, pad > 0 ? pad : 0);
            const double gbs = (double)g * iters * BOX / (ms * 1e-3) / 1e9;
            printf("%-10s %6d %7d %9.3f %12.1f %12.1f\n", "unicast", g, stages, ms, gbs, gbs / (g > sms ? sms : g));
        }
        for (int g : {28, 56, 112, sms / 2 * 2}) {
            const int pad = 100 * 1024 - stages * BOX;
            const float ms = run<1>(map, (int)rows, g, stages, iters, pad > 0 ? pad : 0);
            const double gbs = (double)g * iters * BOX / (ms * 1e-3) / 1e9;           // bytes RECEIVED by the SMs
            printf("%-10s %6d %7d %9.3f %12.1f %12.1f\n", "multicast2", g, stages, ms, gbs, gbs / g);
        }
    }
    return 0;
}
