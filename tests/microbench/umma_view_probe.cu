// umma_view_probe.cu — one question, answered in isolation: can a tcgen05 K-major SWIZZLE_128B operand descriptor address
// the "tap views" of a haloed tile that experimental/conv_igemm_halo.cu relies on?
//   smem: a TMA box [18 rows][16 pixels][64 ch] (36 KB, 128-byte swizzle), exactly what the halo kernel loads;
//   view for tap (kh, kw): 16 groups of 8 pixel-rows, group g at byte offset ((g + kh) * 16 + kw) * 128
//        -> descriptor start = (kh*16 + kw) * 128, SBO = 2048, base_offset = ?
// The probe multiplies every view by a 64x64 identity (B operand), so D[m][c] must equal X[pixel-row of m][c], and counts
// mismatches for base_offset = kw (the guess), base_offset = 0, and for the canonical three-box layout (start on a
// 1024-byte boundary, SBO = 1024), which must pass.  Prints one line per (kh, kw, variant).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../xmem2_b200/csrc -o umma_view_probe \
//        umma_view_probe.cu ../../xmem2_b200/csrc/common.cu && ./umma_view_probe
// Standalone; never run in round 1.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.h"
#include "tc5.cuh"

using namespace tc5;

constexpr int HW_ = 16, HH_ = 18;                      // halo pitch / rows
constexpr int HALO_BYTES = HH_ * HW_ * 128;
constexpr int BOX3 = HH_ * 8 * 128;

struct Maps { CUtensorMap halo, box3, ident; };

struct Smem {
    alignas(1024) uint8_t a[HALO_BYTES];
    alignas(1024) uint8_t a3[3][BOX3];
    alignas(1024) uint8_t b[64 * 128];
    alignas(8) uint64_t full;
    uint64_t done;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t desc_view(uint32_t addr, uint32_t sbo, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// results[(kh*3 + kw)*3 + variant] = number of wrong elements (of 128 x 64)
__global__ void __launch_bounds__(192) probe(const __grid_constant__ Maps maps, const __half* __restrict__ X, int* __restrict__ results) {
    extern __shared__ uint8_t raw[];
    Smem& sm = *reinterpret_cast<Smem*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&sm.full, 1); mbar_init(&sm.done, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&sm.tmem_base, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&sm.full, HALO_BYTES + 3 * BOX3 + 64 * 128);
        tma_load_4d(sm.a, &maps.halo, &sm.full, 0, 0, 0, 0);
        for (int kw = 0; kw < 3; ++kw) tma_load_4d(sm.a3[kw], &maps.box3, &sm.full, 0, kw, 0, 0);
        tma_load_2d(sm.b, &maps.ident, &sm.full, 0, 0);
    }
    mbar_wait(&sm.full, 0, 1);
    __syncthreads();
    int phase = 0;
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
            for (int variant = 0; variant < 3; ++variant) {
                if (warp == 1 && lane == 0) {
                    constexpr uint32_t idesc = make_idesc_f16(128, 64);
                    tc_fence_after();
                    for (int j = 0; j < 4; ++j) {
                        uint64_t ad;
                        if (variant == 2) ad = make_desc_sw128(smem_u32(sm.a3[kw]) + kh * 1024 + j * 32);                       // canonical
                        else ad = desc_view(smem_u32(sm.a) + (kh * HW_ + kw) * 128 + j * 32, HW_ * 128, variant == 0 ? kw : 0);   // the guess / no phase
                        mma_f16_ss(tmem, ad, make_desc_sw128(smem_u32(sm.b) + j * 32), idesc, j ? 1u : 0u);
                    }
                    mma_commit(&sm.done);
                }
                if (warp >= 2) {
                    mbar_wait(&sm.done, phase, 2);
                    tc_fence_after();
                    const int lane_base = (warp & 3) * 32;
                    const int m = lane_base + lane;                 // GEMM row = group g * 8 + x
                    const int g = m >> 3, x = m & 7;
                    const int src_row = (g + kh) * HW_ + kw + x;    // pixel-row of the haloed tile this GEMM row must show
                    int bad = 0;
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem + ((uint32_t)lane_base << 16) + c0, r);
                        tmem_ld_wait();
                        for (int j = 0; j < 32; ++j)
                            bad += (__uint_as_float(r[j]) != __half2float(X[src_row * 64 + c0 + j])) ? 1 : 0;
                    }
                    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
                    if (lane == 0) atomicAdd(&results[(kh * 3 + kw) * 3 + variant], bad);
                    tc_fence_before();
                }
                phase ^= 1;
                __syncthreads();                                   // TMEM is re-used by the next variant
            }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

int main() {
    const int rows = HH_ * HW_;
    std::vector<__half> hx(rows * 64), hi(64 * 64);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < 64; ++c) hx[r * 64 + c] = __float2half((float)((r * 7 + c) % 2048));
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hi[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
    __half *dx, *di; int* dres;
    cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&di, hi.size() * 2); cudaMalloc(&dres, 27 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(di, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dres, 0, 27 * 4);
    Maps maps;
    {   // X as a [C=64][W=16][H=18][B=1] image: the halo box takes all of it, the three-box layout 8 pixels starting at kw
        uint64_t d[4] = {64, (uint64_t)HW_, (uint64_t)HH_, 1}; uint64_t st[3] = {128, (uint64_t)HW_ * 128, (uint64_t)HH_ * HW_ * 128};
        uint32_t b1[4] = {64, (uint32_t)HW_, (uint32_t)HH_, 1}, b3[4] = {64, 8, (uint32_t)HH_, 1};
        if (xm_make_tmap_f16(&maps.halo, dx, 4, d, st, b1) || xm_make_tmap_f16(&maps.box3, dx, 4, d, st, b3)) { printf("tensor map: %s\n", xm_last_error()); return 1; }
        uint64_t d2[2] = {64, 64}; uint64_t s2[1] = {128}; uint32_t b2[2] = {64, 64};
        if (xm_make_tmap_f16(&maps.ident, di, 2, d2, s2, b2)) { printf("tensor map: %s\n", xm_last_error()); return 1; }
    }
    const int smem = (int)sizeof(Smem) + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 192, smem>>>(maps, dx, dres);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    int res[27]; cudaMemcpy(res, dres, sizeof(res), cudaMemcpyDeviceToHost);
    const char* names[3] = {"view, base_offset = kw", "view, base_offset = 0 ", "three-box canonical   "};
    int guess_ok = 1, canon_ok = 1;
    for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) for (int v = 0; v < 3; ++v) {
        const int bad = res[(kh * 3 + kw) * 3 + v];
        printf("kh=%d kw=%d  %s : %5d wrong of 8192  %s\n", kh, kw, names[v], bad, bad ? "" : "OK");
        if (v == 0 && bad) guess_ok = 0;
        if (v == 2 && bad) canon_ok = 0;
    }
    printf("RESULT: tap views with base_offset = kw %s; canonical three-box layout %s\n", guess_ok ? "WORK" : "DO NOT WORK", canon_ok ? "works" : "FAILS (probe bug?)");
    return 0;
}
