// umma_rate.cu — issue/execute rate of tcgen05.mma kind::f16 with both operands in shared memory (SS), M = 128, for
// N = 64 / 128 / 256: clocks per MMA (K = 16) measured from the first issue to the completion commit, one CTA per SM,
// operands resident in smem (no TMA).  Decides the tile shapes of the score sweeps and of the convolutions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../xmem2_b200/csrc -o umma_rate umma_rate.cu ../../xmem2_b200/csrc/common.cu
#include <cstdio>
#include "common.h"
#include "tc5.cuh"
using namespace tc5;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

struct Smem { alignas(1024) uint8_t a[2][128 * 128]; alignas(1024) uint8_t b[2][256 * 128]; alignas(8) uint64_t done; uint32_t tmem_base; };

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int ctas_active, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(sizeof(sm.a) / 4); i += 128) reinterpret_cast<uint32_t*>(sm.a)[i] = 0x3c003c00u;
    for (int i = threadIdx.x; i < (int)(sizeof(sm.b) / 4); i += 128) reinterpret_cast<uint32_t*>(sm.b)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&sm.done, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 0 && lane == 0 && blockIdx.x < ctas_active) {
        constexpr uint32_t idesc = make_idesc_f16(128, N);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t a = make_desc_sw128(smem_u32(sm.a[it & 1]) + j * 32);
                const uint64_t b = make_desc_sw128(smem_u32(sm.b[it & 1]) + j * 32);
                mma_f16_ss(tmem + ((it >> 2) & 1) * N, a, b, idesc, (it | j) ? 1u : 0u);
            }
        }
        t1 = clock64();
        mma_commit(&sm.done);
        mbar_wait(&sm.done, 0, 1);
        t2 = clock64();
        out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
int run(int iters, int active, unsigned long long* out) {
    const int smem = sizeof(Smem) + 1024;
    CK(cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) { k<N><<<148, 128, smem>>>(iters, active, out); CK(cudaDeviceSynchronize()); }
    unsigned long long h[2];
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    const double n = 4.0 * iters;
    printf("M128 N%-3d K16 SS, %3d CTAs: issue %6.1f clk/MMA, complete %6.1f clk/MMA  (ideal %d) -> %5.1f %% of the tensor pipe\n", N, active,
           h[0] / n, h[1] / n, N / 2, 100.0 * (N / 2) / (h[1] / n));
    return 0;
}

int main() {
    unsigned long long* out;
    CK(cudaMalloc(&out, 148 * 16));
    tc5_debug_init();
    for (int active : {1, 148}) {
        if (run<64>(2000, active, out)) return 1;
        if (run<128>(2000, active, out)) return 1;
        if (run<256>(2000, active, out)) return 1;
    }
    return 0;
}
