// tmem_read_bw.cu — how fast can the warps of one SM read TMEM (tcgen05.ld.32x32b.x32)?  Decides whether the score sweeps of
// the fused memory read (k1_affinity.cu) are bound by TMEM reads.  One CTA per SM, 512 TMEM columns, W warps each issuing
// `iters` loads of 32 lanes x 32 columns (4 KB) round-robin over the columns; reports bytes / clock / SM for W = 4, 8, 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_read_bw tmem_read_bw.cu && ./tmem_read_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__global__ void __launch_bounds__(576, 1) k(int nwarps, int iters, unsigned long long* out, float* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tbase;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int i = 0; i < iters; ++i) {
            const uint32_t col = (uint32_t)(((i * 4 + (warp >> 2)) * 32) & 511);
            if (X == 32) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(tmem + lane_base + col) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j] & 0x3f800000u);
            } else {
                uint32_t v[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(tmem + lane_base + col) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += __uint_as_float(v[j] & 0x3f800000u);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    unsigned long long* out; float* sink;
    CK(cudaMalloc(&out, 148 * 8)); CK(cudaMalloc(&sink, 4));
    const int iters = 4000;
    for (int x : {32, 8})
        for (int nw : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (x == 32) k<32><<<148, 576>>>(nw, iters, out, sink); else k<8><<<148, 576>>>(nw, iters, out, sink);
                CK(cudaDeviceSynchronize());
            }
            unsigned long long h[148];
            CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
            const double bytes = (double)nw * iters * 32 * x * 4;
            printf("x%-2d warps %2d: %8llu clks, %7.1f B/clk/SM, %6.1f clks per load per warp\n", x, nw, h[0], bytes / (double)h[0], (double)h[0] / iters);
        }
    return 0;
}
