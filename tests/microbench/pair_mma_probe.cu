// pair_mma_probe.cu — checks, in isolation, the assumptions experimental/conv_igemm_2cta.cu makes about CTA pairs:
//   1. `cp.async.bulk.tensor ... .cta_group::2` issued by the rank-1 CTA into ITS shared memory credits the bytes to the
//      rank-0 CTA's mbarrier when the barrier address has bit 24 cleared (0xFEFFFFFF mask);
//   2. one `tcgen05.mma.cta_group::2` (M = 256, N = 128) issued by the leader reads 128 A rows from EACH CTA and N/2 = 64 B rows
//      from each; rank 0's B rows are output columns 0..63, rank 1's are 64..127;
//   3. `tcgen05.commit.cta_group::2 ... multicast::cluster` with mask 0b11 arrives on the same-offset mbarrier of both CTAs;
//   4. pair-wide TMEM alloc/dealloc by warp 1 of both CTAs.
// Data: A_r[m][k] = (r*128 + m + 3k) mod 1024 (fp16-exact integers), B = [I64 ; 2*I64] (128 x 64), so the accumulator of CTA r
// must be D[m][n] = A_r[m][n] for n < 64 and 2*A_r[m][n-64] for n >= 64.  Prints the number of wrong elements per CTA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../xmem2_b200/csrc -o pair_mma_probe \
//        pair_mma_probe.cu ../../xmem2_b200/csrc/common.cu && ./pair_mma_probe
// Standalone; never run in round 1.  All waits are bounded (a protocol error traps instead of hanging).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

struct Maps { CUtensorMap a, b; };
struct Smem {
    alignas(1024) uint8_t a[128 * 128];
    alignas(1024) uint8_t b[64 * 128];
    alignas(8) uint64_t full;
    uint64_t done;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192) probe(const __grid_constant__ Maps maps, const __half* __restrict__ A, int* __restrict__ wrong) {
    extern __shared__ uint8_t raw[];
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(&sm.full, 1); mbar_init(&sm.done, 1); fence_mbar_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    if (warp == 0 && lane == 0) {
        if (rank == 0) mbar_expect_tx(&sm.full, 2 * (128 * 128 + 64 * 128));      // both CTAs' A tile and B half
        tma_load_2d_2sm(sm.a, &maps.a, &sm.full, 0, (int)rank * 128);              // own 128 A rows
        tma_load_2d_2sm(sm.b, &maps.b, &sm.full, 0, (int)rank * 64);               // own half of the B rows
    } else if (warp == 1 && lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_f16(256, 128);
        mbar_wait(&sm.full, 0, 1);
        tc_fence_after();
        for (int j = 0; j < 4; ++j) {
            const uint64_t ad = make_desc_sw128(smem_u32(sm.a) + j * 32), bd = make_desc_sw128(smem_u32(sm.b) + j * 32);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&sm.done)), "h"((uint16_t)0x3) : "memory");
    } else if (warp >= 2) {
        mbar_wait(&sm.done, 0, 2);                 // rank 1 is released by the leader's multicast commit
        tc_fence_after();
        const int lane_base = (warp & 3) * 32, m = lane_base + lane;
        int bad = 0;
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + ((uint32_t)lane_base << 16) + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) {
                const int n = c0 + j;
                const float a = __half2float(A[((int)rank * 128 + m) * 64 + (n & 63)]);
                bad += (__uint_as_float(r[j]) != (n < 64 ? a : 2.f * a)) ? 1 : 0;
            }
        }
        for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        if (lane == 0) atomicAdd(&wrong[rank], bad);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
    std::vector<__half> ha(256 * 64), hb(128 * 64);
    for (int m = 0; m < 256; ++m) for (int k = 0; k < 64; ++k) ha[m * 64 + k] = __float2half((float)((m + 3 * k) % 1024));
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2half(((n & 63) == k) ? (n < 64 ? 1.f : 2.f) : 0.f);
    __half *da, *db; int* dw;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dw, 8);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dw, 0, 8);
    Maps maps;
    uint64_t da_[2] = {64, 256}, db_[2] = {64, 128}, st[1] = {128};
    uint32_t ba[2] = {64, 128}, bb[2] = {64, 64};
    if (xm_make_tmap_f16(&maps.a, da, 2, da_, st, ba) || xm_make_tmap_f16(&maps.b, db, 2, db_, st, bb)) { printf("tensor map: %s\n", xm_last_error()); return 1; }
    const int smem = (int)sizeof(Smem) + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<2, 192, smem>>>(maps, da, dw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s (a trap = a bounded wait timed out: protocol assumption wrong)\n", cudaGetErrorString(e)); return 1; }
    int w[2]; cudaMemcpy(w, dw, 8, cudaMemcpyDeviceToHost);
    printf("CTA 0: %d wrong of 16384, CTA 1: %d wrong of 16384\nRESULT: pair protocol %s\n", w[0], w[1], (w[0] || w[1]) ? "DOES NOT match the assumptions" : "matches the assumptions");
    return 0;
}
