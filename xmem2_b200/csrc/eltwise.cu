// eltwise.cu — the small memory-bound kernels around the tensor-core convolutions (NHWC fp16 activations,
// fp32 math).  Each replaces a handful of separate PyTorch launches in the reference:
//   xm_im2col_stem        7x7/s2 stem patches of the key / value encoder (model/resnet.py:120, modules.py:124-136)
//   xm_maxpool3x3s2       nn.MaxPool2d(3,2,1) (+ReLU after it for the value encoder, modules.py:137-138)
//   xm_relu               F.relu on a GroupResBlock input (group_modules.py:47)
//   xm_keyproj_post       shrinkage = d^2+1, selection = sigmoid(e) (modules.py:207-211) + query packing
//   xm_cbam_*             CBAM channel + spatial gates (model/cbam.py:23-77) and the "+ r" of FeatureFusionBlock
//   xm_upsample2x_add     bilinear x2 + broadcast skip add (modules.py:186-189, group_modules.py:15-23)
//   xm_area_down          area (average) down-sampling for HiddenUpdater (modules.py:58-59, group_modules.py:25-26)
//   xm_gru                the non-textbook GRU update (modules.py:63-74, 88-99)
//   xm_upsample4x_aggregate  bilinear x4 + sigmoid + soft aggregation (modules.py:247, network.py:110-115, aggregate.py:6-16)
//   xm_value_append       [hw,512] NHWC value -> [512][cap] column-major bank arena (kv_memory_store.py:70)
#include "common.h"
#include "tc5.cuh"
#include <cuda_fp16.h>

using tc5::pdl_wait;
using tc5::pdl_launch_dependents;

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------- stem im2col
// out[b][yo][xo][(kh*7+kw)*C + c] = in_c(2*yo+kh-3, 2*xo+kw-3), zero padded to kpad.
// channels: 0..2 image (shared by all b), 3 = mask[b], 4 = sum_{j!=b} mask[j]  (only when masks != null)
__global__ void im2col_stem_kernel(const float* __restrict__ image, const float* __restrict__ masks, int n, int H, int W,
                                   int C, int kpad, __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    // one thread = 8 consecutive k of one output pixel -> one 16-byte store (coalesced across the warp)
    const int Ho = H / 2, Wo = W / 2, K8 = kpad / 8;
    const size_t total = (size_t)n * Ho * Wo * K8;
    const size_t plane = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k8 = i % K8;
        size_t p = i / K8;
        const int xo = p % Wo; p /= Wo;
        const int yo = p % Ho;
        const int b = p / Ho;
        float v[8];
        int k = k8 * 8;
        int tap = k / C, c = k - tap * C;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float val = 0.f;
            if (tap < 49) {
                const int y = 2 * yo + tap / 7 - 3, x = 2 * xo + tap % 7 - 3;
                if (y >= 0 && y < H && x >= 0 && x < W) {
                    const size_t off = (size_t)y * W + x;
                    if (c < 3) val = __ldg(image + c * plane + off);
                    else if (c == 3) val = __ldg(masks + b * plane + off);
                    else {
                        for (int j = 0; j < n; ++j) if (j != b) val += __ldg(masks + j * plane + off);
                    }
                }
            }
            v[e] = val;
            if (++c == C) { c = 0; ++tap; }
        }
        uint4 u;
        u.x = *reinterpret_cast<const uint32_t*>(&(const __half2&)__floats2half2_rn(v[0], v[1]));
        u.y = *reinterpret_cast<const uint32_t*>(&(const __half2&)__floats2half2_rn(v[2], v[3]));
        u.z = *reinterpret_cast<const uint32_t*>(&(const __half2&)__floats2half2_rn(v[4], v[5]));
        u.w = *reinterpret_cast<const uint32_t*>(&(const __half2&)__floats2half2_rn(v[6], v[7]));
        reinterpret_cast<uint4*>(out)[i] = u;
    }
}

// Key-encoder stem (3 image channels): K layout k = kh*24 + kw*3 + c (kw<7, c<3; 3 zero slots per kernel row, rows 7 is all
// zero -> 192).  One thread builds one (pixel, kernel row) = 24 halfs = three 16-byte stores.
__global__ void im2col_stem3_kernel(const float* __restrict__ image, int H, int W, __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)Ho * Wo * 8;
    const size_t plane = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kh = i & 7;
        const size_t p = i >> 3;
        const int xo = p % Wo, yo = p / Wo;
        float v[24];
#pragma unroll
        for (int e = 0; e < 24; ++e) v[e] = 0.f;
        const int y = 2 * yo + kh - 3;
        if (kh < 7 && y >= 0 && y < H) {
            const int xb = 2 * xo - 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* rowp = image + c * plane + (size_t)y * W;
#pragma unroll
                for (int kw = 0; kw < 7; ++kw) {
                    const int x = xb + kw;
                    if (x >= 0 && x < W) v[kw * 3 + c] = __ldg(rowp + x);
                }
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + p * 192 + kh * 24);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint4 u;
            u.x = tc5::pack_half2(v[8 * q], v[8 * q + 1]); u.y = tc5::pack_half2(v[8 * q + 2], v[8 * q + 3]);
            u.z = tc5::pack_half2(v[8 * q + 4], v[8 * q + 5]); u.w = tc5::pack_half2(v[8 * q + 6], v[8 * q + 7]);
            dst[q] = u;
        }
    }
}

// ---------------------------------------------------------------- maxpool 3x3 s2 p1 (8 channels / thread)
__global__ void maxpool_kernel(const __half* __restrict__ in, int B, int H, int W, int C, int relu, __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const size_t total = (size_t)B * Ho * Wo * C8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = i % C8;
        size_t p = i / C8;
        const int xo = p % Wo; p /= Wo;
        const int yo = p % Ho;
        const int b = p / Ho;
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = relu ? 0.f : -INFINITY;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = 2 * yo + dy;
            if (y < 0 || y >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = 2 * xo + dx;
                if (x < 0 || x >= W) continue;
                uint4 u = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + y) * W + x) * C + c8 * 8);
                const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 f = __half22float2(h[e]);
                    m[2 * e] = fmaxf(m[2 * e], f.x); m[2 * e + 1] = fmaxf(m[2 * e + 1], f.y);
                }
            }
        }
        __half2 o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __floats2half2_rn(m[2 * e], m[2 * e + 1]);
        *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + yo) * Wo + xo) * C + c8 * 8) = *reinterpret_cast<uint4*>(o);
    }
}

__global__ void relu_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n8) {
    pdl_wait();
    pdl_launch_dependents();
    const __half2 z = __float2half2_rn(0.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        uint4 u = in[i];
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], z);
        out[i] = u;
    }
}

// ---------------------------------------------------------------- key projection post-processing
// proj [hw][pstride] = (key 0..63 | d 64 | e 65..128).  One warp per position.
__global__ void keyproj_post_kernel(const __half* __restrict__ proj, int pstride, int hw, int hw_pad, __half* __restrict__ key,
                                    __half* __restrict__ sel, float* __restrict__ shr, __half* __restrict__ qp,
                                    float* __restrict__ bsq) {
    pdl_wait();
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= hw_pad) return;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        __half k = __float2half(0.f), e = __float2half(0.f);
        if (row < hw) {
            k = proj[(size_t)row * pstride + c];
            e = __float2half_rn(sigmoidf_(__half2float(proj[(size_t)row * pstride + 65 + c])));
            key[(size_t)row * 64 + c] = k;
            sel[(size_t)row * 64 + c] = e;
        }
        if (qp) {
            const __half ke = __hmul(k, e);
            qp[(size_t)row * 128 + c] = __hneg(e);
            qp[(size_t)row * 128 + 64 + c] = __hadd(ke, ke);
        }
        const float kf = __half2float(k);
        acc += __half2float(e) * (kf * kf);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        if (bsq) bsq[row] = acc;
        if (row < hw) {
            const float d = __half2float(proj[(size_t)row * pstride + 64]);
            shr[row] = d * d + 1.f;
        }
    }
}

// ---------------------------------------------------------------- CBAM
// (a) per-(image, channel) partial sum and max over a slice of the pixels.  grid (C/64, B, CBAM_SLICES), block (64, 4)
constexpr int CBAM_SLICES = 16;
__global__ void cbam_pool_kernel(const __half* __restrict__ x, int HW, int C, float* __restrict__ psum, float* __restrict__ pmax) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float ssum[4][64], smax[4][64];
    const int c = blockIdx.x * 64 + threadIdx.x, b = blockIdx.y, sl = blockIdx.z;
    const int per = (HW + CBAM_SLICES - 1) / CBAM_SLICES;
    const int p0 = sl * per, p1 = min(HW, p0 + per);
    float s = 0.f, m = -INFINITY;
    for (int p = p0 + threadIdx.y; p < p1; p += 4) {
        const float v = __half2float(x[((size_t)b * HW + p) * C + c]);
        s += v; m = fmaxf(m, v);
    }
    ssum[threadIdx.y][threadIdx.x] = s; smax[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0) {
        for (int i = 1; i < 4; ++i) { s += ssum[i][threadIdx.x]; m = fmaxf(m, smax[i][threadIdx.x]); }
        psum[((size_t)b * CBAM_SLICES + sl) * C + c] = s;
        pmax[((size_t)b * CBAM_SLICES + sl) * C + c] = m;
    }
}
// (b) scale_c = sigmoid(mlp(avg) + mlp(max)).  One block (C threads) per image; hidden width R = C/16.
// FAST = the fuser's shape (C = 512, R = 32, 512 threads): both weight matrices are pulled into registers BEFORE the
// programmatic-dependency wait (they are constants, nothing upstream writes them), so the 128 KB of weight traffic overlaps the
// pooling kernel's tail and the dependent part is two short reductions instead of a dozen serial memory round trips.
template <bool FAST>
__global__ void __launch_bounds__(512) cbam_mlp_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, int HW,
                                const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                const float* __restrict__ b2, int C, int R, float* __restrict__ scale) {
    extern __shared__ float sh[];          // [2][C] inputs, [2][R] hidden
    float* in0 = sh; float* in1 = sh + C; float* h0 = sh + 2 * C; float* h1 = h0 + R;
    const int b = blockIdx.x, t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    float w1r[2][16]; float4 w2r[8]; float b1r[2], b2r = 0.f;
    if (FAST) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 16; ++e) w1r[j][e] = __ldg(w1 + (size_t)(warp + 16 * j) * 512 + lane + 32 * e);
            b1r[j] = __ldg(b1 + warp + 16 * j);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) w2r[e] = __ldg(reinterpret_cast<const float4*>(w2 + (size_t)t * 32) + e);
        b2r = __ldg(b2 + t);
    }
    pdl_wait();
    pdl_launch_dependents();
    {
        float s = 0.f, m = -INFINITY;
#pragma unroll
        for (int sl = 0; sl < CBAM_SLICES; ++sl) {
            s += psum[((size_t)b * CBAM_SLICES + sl) * C + t];
            m = fmaxf(m, pmax[((size_t)b * CBAM_SLICES + sl) * C + t]);
        }
        in0[t] = s / HW; in1[t] = m;
    }
    __syncthreads();
    if (FAST) {
        // warp w owns hidden units w and w+16 for BOTH pooled inputs (they share the weight row)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) { a0 += w1r[j][e] * in0[lane + 32 * e]; a1 += w1r[j][e] * in1[lane + 32 * e]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
            if (lane == 0) { h0[warp + 16 * j] = fmaxf(a0 + b1r[j], 0.f); h1[warp + 16 * j] = fmaxf(a1 + b1r[j], 0.f); }
        }
        __syncthreads();
        float a = 2.f * b2r;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            a += w2r[e].x * (h0[4 * e] + h1[4 * e]);
            a += w2r[e].y * (h0[4 * e + 1] + h1[4 * e + 1]);
            a += w2r[e].z * (h0[4 * e + 2] + h1[4 * e + 2]);
            a += w2r[e].w * (h0[4 * e + 3] + h1[4 * e + 3]);
        }
        scale[(size_t)b * C + t] = sigmoidf_(a);
        return;
    }
    // hidden layer: one warp per (input, unit) pair, lanes stride over C
    for (int u = warp; u < 2 * R; u += nwarps) {
        const int r = u % R; const float* in = (u < R) ? in0 : in1;
        float a = 0.f;
#pragma unroll 8
        for (int c = lane; c < C; c += 32) a += __ldg(w1 + (size_t)r * C + c) * in[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) (u < R ? h0 : h1)[r] = fmaxf(a + b1[r], 0.f);
    }
    __syncthreads();
    float a = 2.f * b2[t];
#pragma unroll 8
    for (int r = 0; r < R; ++r) a += __ldg(w2 + (size_t)t * R + r) * (h0[r] + h1[r]);
    scale[(size_t)b * C + t] = sigmoidf_(a);
}
// (c) per pixel: max and mean over channels of x * scale_c.  One warp per pixel.
__global__ void cbam_spatial_pool_kernel(const __half* __restrict__ x, const float* __restrict__ scale, int B, int HW, int C,
                                         float* __restrict__ comp) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t pix = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pix >= (size_t)B * HW) return;
    const int b = pix / HW;
    float s = 0.f, m = -INFINITY;
    for (int c = lane; c < C; c += 32) {
        const float v = __half2float(x[pix * C + c]) * scale[(size_t)b * C + c];
        s += v; m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if (lane == 0) { comp[pix * 2] = m; comp[pix * 2 + 1] = s / C; }
}
// (d) s = sigmoid(conv7x7(comp)); out = x * (1 + scale_c * s)  [= x + CBAM(x)], plus a ReLU'd copy
__global__ void cbam_apply_kernel(const __half* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ comp,
                                  const float* __restrict__ w7, float b7, int B, int H, int W, int C, __half* __restrict__ out,
                                  __half* __restrict__ out_relu) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t pix = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pix >= (size_t)B * H * W) return;
    const int xw = pix % W, y = (pix / W) % H, b = pix / ((size_t)H * W);
    float a = 0.f;
    for (int t = lane; t < 98; t += 32) {
        const int ch = t / 49, tap = t % 49;
        const int yy = y + tap / 7 - 3, xx = xw + tap % 7 - 3;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) a += w7[t] * comp[(((size_t)b * H + yy) * W + xx) * 2 + ch];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const float s = sigmoidf_(a + b7);
    for (int c = lane; c < C; c += 32) {
        const float v = __half2float(x[pix * C + c]) * (1.f + scale[(size_t)b * C + c] * s);
        out[pix * C + c] = __float2half_rn(v);
        if (out_relu) out_relu[pix * C + c] = __float2half_rn(fmaxf(v, 0.f));
    }
}

// ---------------------------------------------------------------- bilinear x2 (align_corners=False) + broadcast skip
// 8 channels (16 bytes) per thread
__global__ void upsample2x_add_kernel(const __half* __restrict__ g, const __half* __restrict__ skip, int B, int h, int w, int C,
                                      __half* __restrict__ out, __half* __restrict__ out_relu) {
    pdl_wait();
    pdl_launch_dependents();
    const int H = 2 * h, W = 2 * w, C8 = C / 8;
    const size_t total = (size_t)B * H * W * C8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = i % C8;
        size_t p = i / C8;
        const int x = p % W; p /= W;
        const int y = p % H;
        const int b = p / H;
        const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float fy = sy - y0, fx = sx - x0;
        const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
        const uint4* gb = reinterpret_cast<const uint4*>(g) + (size_t)b * h * w * C8;
        const uint4 a00 = __ldg(gb + ((size_t)y0 * w + x0) * C8 + c8), a01 = __ldg(gb + ((size_t)y0 * w + x1) * C8 + c8);
        const uint4 a10 = __ldg(gb + ((size_t)y1 * w + x0) * C8 + c8), a11 = __ldg(gb + ((size_t)y1 * w + x1) * C8 + c8);
        const uint4 sk = __ldg(reinterpret_cast<const uint4*>(skip) + ((size_t)y * W + x) * C8 + c8);
        const __half2* h00 = reinterpret_cast<const __half2*>(&a00); const __half2* h01 = reinterpret_cast<const __half2*>(&a01);
        const __half2* h10 = reinterpret_cast<const __half2*>(&a10); const __half2* h11 = reinterpret_cast<const __half2*>(&a11);
        const __half2* hs = reinterpret_cast<const __half2*>(&sk);
        uint4 o, orl;
        __half2* ho = reinterpret_cast<__half2*>(&o); __half2* hr = reinterpret_cast<__half2*>(&orl);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 v00 = __half22float2(h00[e]), v01 = __half22float2(h01[e]), v10 = __half22float2(h10[e]), v11 = __half22float2(h11[e]);
            const float2 s2 = __half22float2(hs[e]);
            const float rx = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x + s2.x;
            const float ry = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y + s2.y;
            ho[e] = __floats2half2_rn(rx, ry);
            hr[e] = __floats2half2_rn(fmaxf(rx, 0.f), fmaxf(ry, 0.f));
        }
        reinterpret_cast<uint4*>(out)[i] = o;
        if (out_relu) reinterpret_cast<uint4*>(out_relu)[i] = orl;
    }
}

// ---------------------------------------------------------------- area down-sampling by f (+ optional extra channel)
// out[b][y][x][c] = mean over fxf of in; channel C (if extra) = mean of extra[b][.][.]; channels up to cpad are zero.
// 8 output channels per thread (C and cpad multiples of 8).
__global__ void area_down_kernel(const __half* __restrict__ in, const __half* __restrict__ extra, int B, int H, int W, int C,
                                 int f, int cpad, __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int Ho = H / f, Wo = W / f, P8 = cpad / 8;
    const size_t total = (size_t)B * Ho * Wo * P8;
    const float inv = 1.f / (f * f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c8 = i % P8;
        size_t p = i / P8;
        const int xo = p % Wo; p /= Wo;
        const int yo = p % Ho;
        const int b = p / Ho;
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        if (c8 * 8 < C) {
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) {
                    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + yo * f + dy) * W + xo * f + dx) * C + c8 * 8));
                    const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) { const float2 v = __half22float2(hh[e]); s[2 * e] += v.x; s[2 * e + 1] += v.y; }
                }
        } else if (c8 * 8 == C && extra) {
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) s[0] += __half2float(extra[((size_t)b * H + yo * f + dy) * W + xo * f + dx]);
        }
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) ho[e] = __floats2half2_rn(s[2 * e] * inv, s[2 * e + 1] * inv);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

// ---------------------------------------------------------------- single-output 3x3 convolution (decoder.pred, modules.py:227,239)
// logits[b][y][x] = bias + sum_{tap,c} w[tap][c] * in[b][y+dy][x+dx][c].  One block = 8x8 output pixels: the 10x10xC halo
// tile is staged in shared memory once (instead of 9 L2 reads per pixel); one warp per output row, lanes split channels.
__global__ void __launch_bounds__(256, 3) conv3x3_c1_kernel(const __half* __restrict__ in, const __half* __restrict__ wgt, float bias, int B, int H, int W, int C,
                                  __half* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ uint4 tile_sm[];               // [10][10][C/8] uint4
    const int C8 = C / 8;
    const int tiles_x = (W + 7) / 8, tiles_y = (H + 7) / 8;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; const int b = t / tiles_y;
    const int x0 = tx * 8 - 1, y0 = ty * 8 - 1;
    if (C8 == 32 && blockDim.x == 256) {
        // 100 halo pixels x 32 uint4: thread = (pixel slot, channel octet); the loads are issued in two batches of seven before
        // their stores, so the fill costs two memory round trips instead of thirteen
        const int c8 = threadIdx.x & 31, p0 = threadIdx.x >> 5;
#pragma unroll 1
        for (int j0 = 0; j0 < 14; j0 += 7) {
            uint4 v[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int pix = p0 + 8 * (j0 + j), py = pix / 10, px = pix - py * 10;
                const int y = y0 + py, x = x0 + px;
                v[j] = make_uint4(0u, 0u, 0u, 0u);
                if (pix < 100 && y >= 0 && y < H && x >= 0 && x < W)
                    v[j] = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + y) * W + x) * C) + c8);
            }
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (p0 + 8 * (j0 + j) < 100) tile_sm[(p0 + 8 * (j0 + j)) * 32 + c8] = v[j];
        }
    } else {
        for (int i = threadIdx.x; i < 100 * C8; i += blockDim.x) {
            const int c8 = i % C8, px = (i / C8) % 10, py = i / (C8 * 10);
            const int y = y0 + py, x = x0 + px;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + y) * W + x) * C) + c8);
            tile_sm[i] = v;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;       // 8 warps: warp = output row inside the patch
    // this lane's weights: channels lane*8 .. +7 of every tap (C8 == 32 for the decoder)
    for (int c8 = lane; c8 < C8; c8 += 32) {
        uint4 wv[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) wv[tap] = __ldg(reinterpret_cast<const uint4*>(wgt + tap * C) + c8);
        for (int px = 0; px < 8; ++px) {
            float acc = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint4 a = tile_sm[((warp + tap / 3) * 10 + px + tap % 3) * C8 + c8];
                const __half2* ha = reinterpret_cast<const __half2*>(&a); const __half2* hw2 = reinterpret_cast<const __half2*>(&wv[tap]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 fa = __half22float2(ha[e]), fw = __half22float2(hw2[e]);
                    acc += fa.x * fw.x + fa.y * fw.y;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const int y = ty * 8 + warp, x = tx * 8 + px;
            if (lane == 0 && y < H && x < W) {
                if (c8 < 32) out[((size_t)b * H + y) * W + x] = __float2half_rn(acc + bias);
            }
        }
    }
}

// ---------------------------------------------------------------- GRU (modules.py:68-72): h' = f*h*(1-u) + u*tanh(v)
__global__ void gru_kernel(const __half* __restrict__ values, const float* __restrict__ h, size_t npix, int hd,
                           float* __restrict__ h_out, __half* __restrict__ h_out16) {
    pdl_wait();
    pdl_launch_dependents();
    const size_t total = npix * hd;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / hd; const int c = i % hd;
        const __half* v = values + p * 3 * hd;
        const float f = sigmoidf_(__half2float(v[c]));
        const float u = sigmoidf_(__half2float(v[hd + c]));
        const float nv = tanhf(__half2float(v[2 * hd + c]));
        const float r = f * h[i] * (1.f - u) + u * nv;
        h_out[i] = r;
        h_out16[i] = __float2half_rn(r);
    }
}

// ---------------------------------------------------------------- bilinear x4 + sigmoid + soft aggregation
// logits4 [n][h4][w4] fp16 -> prob [n+1][H][W] fp32 (and logits) with H=4*h4.
__global__ void upsample4x_aggregate_kernel(const __half* __restrict__ l4, int n, int h4, int w4, float* __restrict__ prob,
                                            float* __restrict__ logits_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int H = 4 * h4, W = 4 * w4;
    const size_t total = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = i % W, y = i / W;
        const float sy = fmaxf((y + 0.5f) * 0.25f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.25f - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, h4 - 1), x1 = min(x0 + 1, w4 - 1);
        const float fy = sy - y0, fx = sx - x0;
        float lg[XM_MAX_GROUPS * 4 + 1];
        float bg = 1.f;
        for (int o = 0; o < n; ++o) {
            const __half* b = l4 + (size_t)o * h4 * w4;
            const float v = (1.f - fy) * ((1.f - fx) * __half2float(b[y0 * w4 + x0]) + fx * __half2float(b[y0 * w4 + x1])) +
                            fy * ((1.f - fx) * __half2float(b[y1 * w4 + x0]) + fx * __half2float(b[y1 * w4 + x1]));
            const float pr = sigmoidf_(v);
            bg *= (1.f - pr);
            const float pc = fminf(fmaxf(pr, 1e-7f), 1.f - 1e-7f);
            lg[o + 1] = logf(pc / (1.f - pc));
        }
        const float bc = fminf(fmaxf(bg, 1e-7f), 1.f - 1e-7f);
        lg[0] = logf(bc / (1.f - bc));
        float m = lg[0];
        for (int o = 1; o <= n; ++o) m = fmaxf(m, lg[o]);
        float den = 0.f;
        for (int o = 0; o <= n; ++o) den += expf(lg[o] - m);
        for (int o = 0; o <= n; ++o) {
            prob[(size_t)o * total + i] = expf(lg[o] - m) / den;
            if (logits_out) logits_out[(size_t)o * total + i] = lg[o];
        }
    }
}

// ---------------------------------------------------------------- value [n][hw][512] -> arena [obj][512][cap] at column col0
__global__ void value_append_kernel(const __half* __restrict__ val, int hw, int cap, int col0, __half* __restrict__ arena) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ __half tile[32][34];
    const int o = blockIdx.z;
    const int q0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int q = q0 + r;
        tile[r][threadIdx.x] = (q < hw) ? val[((size_t)o * hw + q) * XM_CV + c0 + threadIdx.x] : __float2half(0.f);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int q = q0 + threadIdx.x;
        if (q < hw) arena[((size_t)o * XM_CV + c0 + r) * cap + col0 + q] = tile[threadIdx.x][r];
    }
}

inline int grid_for(size_t total, int block = 256) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)xm_num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

#define STREAM ((cudaStream_t)stream)

extern "C" int xm_im2col_stem(const float* image, const float* masks, int32_t n, int32_t H, int32_t W, int32_t kpad, void* out, void* stream) {
    XM_REQUIRE(image && out && n >= 1 && H % 2 == 0 && W % 2 == 0, "xm_im2col_stem: bad arguments");
    const int C = masks ? 5 : 3;
    XM_REQUIRE(kpad >= 49 * C && kpad % 64 == 0, "xm_im2col_stem: kpad must be a multiple of 64 >= %d", 49 * C);
    if (!masks && n == 1 && kpad == 192) {        // key-encoder stem: row-padded K layout (kh*24 + kw*3 + c)
        const size_t tot = (size_t)(H / 2) * (W / 2) * 8;
        XM_CHECK_CUDA(tc5_launch(im2col_stem3_kernel, dim3(grid_for(tot)), dim3(256), 0, STREAM, image, H, W, (__half*)out));
        xm_count_launches(1);
        return XM_OK;
    }
    const size_t total = (size_t)n * (H / 2) * (W / 2) * (kpad / 8);
    XM_CHECK_CUDA(tc5_launch(im2col_stem_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, image, masks, n, H, W, C, kpad, (__half*)out));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_maxpool3x3s2(const void* in, int32_t B, int32_t H, int32_t W, int32_t C, int32_t relu, void* out, void* stream) {
    XM_REQUIRE(in && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "xm_maxpool3x3s2: bad arguments");
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    XM_CHECK_CUDA(tc5_launch(maxpool_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, (const __half*)in, B, H, W, C, relu, (__half*)out));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_relu(const void* in, void* out, int64_t n, void* stream) {
    XM_REQUIRE(in && out && n % 8 == 0, "xm_relu: n must be a multiple of 8");
    XM_CHECK_CUDA(tc5_launch(relu_kernel, dim3(grid_for(n / 8)), dim3(256), 0, STREAM, (const uint4*)in, (uint4*)out, (size_t)n / 8));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_keyproj_post(const void* proj, int32_t pstride, int32_t hw, int32_t hw_pad, void* key, void* sel, float* shr,
                               void* qp, float* bsq, void* stream) {
    XM_REQUIRE(proj && key && sel && shr && pstride >= 129 && hw > 0 && hw_pad >= hw, "xm_keyproj_post: bad arguments");
    const int rows = qp ? hw_pad : hw;
    XM_CHECK_CUDA(tc5_launch(keyproj_post_kernel, dim3((rows + 7) / 8), dim3(256), 0, STREAM, (const __half*)proj, pstride, hw, rows, (__half*)key, (__half*)sel, shr,
                                                            (__half*)qp, bsq));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_cbam(const void* x, int32_t B, int32_t H, int32_t W, int32_t C, const float* w1, const float* b1, const float* w2,
                       const float* b2, const float* w7, float b7, float* scratch, void* out, void* out_relu, void* stream) {
    // scratch: psum[B*16*C] | pmax[B*16*C] | scale[B*C] | comp[B*H*W*2]  floats
    XM_REQUIRE(x && out && scratch && C % 64 == 0 && C <= 1024, "xm_cbam: bad arguments");
    const int HW = H * W, R = C / 16;
    float* avg = scratch; float* mx = avg + (size_t)B * CBAM_SLICES * C; float* scale = mx + (size_t)B * CBAM_SLICES * C;
    float* comp = scale + (size_t)B * C;
    XM_CHECK_CUDA(tc5_launch(cbam_pool_kernel, dim3(dim3(C / 64, B, CBAM_SLICES)), dim3(dim3(64, 4)), 0, STREAM, (const __half*)x, HW, C, avg, mx));
    if (C == 512 && R == 32)
        XM_CHECK_CUDA(tc5_launch(cbam_mlp_kernel<true>, dim3(B), dim3(C), (2 * C + 2 * R) * sizeof(float), STREAM, avg, mx, HW, w1, b1, w2, b2, C, R, scale));
    else
        XM_CHECK_CUDA(tc5_launch(cbam_mlp_kernel<false>, dim3(B), dim3(C), (2 * C + 2 * R) * sizeof(float), STREAM, avg, mx, HW, w1, b1, w2, b2, C, R, scale));
    const size_t npix = (size_t)B * HW;
    XM_CHECK_CUDA(tc5_launch(cbam_spatial_pool_kernel, dim3((unsigned)((npix + 7) / 8)), dim3(256), 0, STREAM, (const __half*)x, scale, B, HW, C, comp));
    XM_CHECK_CUDA(tc5_launch(cbam_apply_kernel, dim3((unsigned)((npix + 7) / 8)), dim3(256), 0, STREAM, (const __half*)x, scale, comp, w7, b7, B, H, W, C, (__half*)out,
                                                                     (__half*)out_relu));
    xm_count_launches(3);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_upsample2x_add(const void* g, const void* skip, int32_t B, int32_t h, int32_t w, int32_t C, void* out, void* out_relu,
                                 void* stream) {
    XM_REQUIRE(g && skip && out && C % 8 == 0, "xm_upsample2x_add: C must be a multiple of 8");
    const size_t total = (size_t)B * 4 * h * w * (C / 8);
    XM_CHECK_CUDA(tc5_launch(upsample2x_add_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, (const __half*)g, (const __half*)skip, B, h, w, C, (__half*)out,
                                                               (__half*)out_relu));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_area_down(const void* in, const void* extra, int32_t B, int32_t H, int32_t W, int32_t C, int32_t f, int32_t cpad,
                            void* out, void* stream) {
    XM_REQUIRE(in && out && f >= 1 && H % f == 0 && W % f == 0 && cpad >= C + (extra ? 1 : 0) && C % 8 == 0 && cpad % 8 == 0,
               "xm_area_down: bad arguments");
    const size_t total = (size_t)B * (H / f) * (W / f) * (cpad / 8);
    XM_CHECK_CUDA(tc5_launch(area_down_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, (const __half*)in, (const __half*)extra, B, H, W, C, f, cpad, (__half*)out));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_conv3x3_c1(const void* in, const void* weight_tap_c, float bias, int32_t B, int32_t H, int32_t W, int32_t C, void* out,
                             void* stream) {
    XM_REQUIRE(in && weight_tap_c && out && C == 256, "xm_conv3x3_c1: specialised for C = 256 (decoder.pred)");
    const int blocks = B * ((H + 7) / 8) * ((W + 7) / 8);
    const size_t smem = (size_t)100 * (C / 8) * 16;
    static XmPerDevice attr_token = {0};
    if (xm_first_use_on_device(&attr_token)) {
        XM_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    XM_CHECK_CUDA(tc5_launch(conv3x3_c1_kernel, dim3(blocks), dim3(256), smem, STREAM, (const __half*)in, (const __half*)weight_tap_c,
                             bias, B, H, W, C, (__half*)out));
    xm_count_launches(1);
    return XM_OK;
}

extern "C" int xm_gru(const void* values, const float* h, int64_t npix, int32_t hidden_dim, float* h_out, void* h_out16, void* stream) {
    XM_REQUIRE(values && h && h_out && h_out16 && npix > 0, "xm_gru: bad arguments");
    XM_CHECK_CUDA(tc5_launch(gru_kernel, dim3(grid_for((size_t)npix * hidden_dim)), dim3(256), 0, STREAM, (const __half*)values, h, (size_t)npix, hidden_dim, h_out,
                                                                      (__half*)h_out16));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_upsample4x_aggregate(const void* logits4, int32_t n, int32_t h4, int32_t w4, float* prob, float* logits, void* stream) {
    XM_REQUIRE(logits4 && prob && n >= 1 && n <= XM_MAX_GROUPS * 4, "xm_upsample4x_aggregate: bad arguments");
    XM_CHECK_CUDA(tc5_launch(upsample4x_aggregate_kernel, dim3(grid_for((size_t)16 * h4 * w4)), dim3(256), 0, STREAM, (const __half*)logits4, n, h4, w4, prob, logits));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}

extern "C" int xm_value_append(const void* value_hwc, int32_t n_obj, int32_t hw, void* arena, int64_t cap, int32_t col0, void* stream) {
    XM_REQUIRE(value_hwc && arena && n_obj >= 1 && hw > 0 && col0 >= 0 && col0 + hw <= cap, "xm_value_append: bad arguments");
    XM_CHECK_CUDA(tc5_launch(value_append_kernel, dim3(dim3((hw + 31) / 32, XM_CV / 32, n_obj)), dim3(dim3(32, 8)), 0, STREAM, (const __half*)value_hwc, hw, (int)cap, col0,
                                                                                           (__half*)arena));
    xm_count_launches(1);
    XM_CHECK_CUDA(cudaGetLastError());
    return XM_OK;
}
