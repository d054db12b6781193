// SVTRv2 recogniser: the kernels around the tcgen05 GEMMs.
//
// Every nn.Linear / nn.Conv2d of model/rec2/svtrv2.py runs on igemm.cu's implicit-GEMM kernel (a token matrix is an NHWC
// image whose pixels are tokens: Linear = 1x1 conv, the grouped 3x3 convs of LocalMixing = a 3x3 conv with block-diagonal
// weights, bias / folded BatchNorm / residual add in the epilogue).  This file holds what is left (reference lines in
// model/rec2/svtrv2.py):
//   svtr_stem1_kernel     ConvStem.conv1 + bn1 + GELU (:112-114,120): 3 input channels, fp32 NCHW crops -> NHWC bf16
//   layernorm_kernel      nn.LayerNorm over the channel axis (:93,95,173,175,184-187,446)
//   gelu_kernel           nn.GELU, exact erf form (:33,51,54,114,117)
//   mhsa_kernel           softmax(q k^T * scale) v per (sequence, head), head_dim 32 (:79-86, :208-215)
//   frm_select_kernel     FRM vertical cross-attention of the selecting token over the H rows of a column (:236-243)
//   decimate_rows_kernel  rows 0, 2, 4, .. of an NHWC map (the first PatchMerging path; the plan now uses the GEMM kernel's
//                         stride (2, 1) mode, the entry point stays for callers that hold a stride-1 output)
//   log_softmax_tbc       F.log_softmax over classes + the (B, T, C) -> (T, B, C) permute (:536-537)
// First correct path: CUDA-core attention (fp32 accumulate, K / V of one head resident in shared memory as fp32); moving it
// to tcgen05 is the next step for this row.
#include "internal.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

namespace b200 {

__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* f) {
    return make_uint4(pack2_bf16(f[0], f[1]), pack2_bf16(f[2], f[3]), pack2_bf16(f[4], f[5]), pack2_bf16(f[6], f[7]));
}
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. 2^-15 of a bf16 ulp at |erf| ~ 1): one MUFU.RCP, one MUFU.EX2
// and seven FMAs instead of erff's ~30 instructions -- the GELU kernels are issue-bound, not HBM-bound, with erff.
__device__ __forceinline__ float erf_as(float x) {
    const float a = fabsf(x);
    float t;  // MUFU.RCP (1 ulp): __frcp_rn would be a ~25-instruction IEEE sequence
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = 1.0f - p * t * __expf(-a * a);
    return copysignf(e, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f)); }

// ------------------------------------------------------------------------------------------ stem conv1
// one thread = one output pixel, all `mid` (<= 64) output channels; weights [mid][27] (ci, ky, kx) + folded scale/shift in smem
__global__ void __launch_bounds__(128) svtr_stem1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         __nv_bfloat16* __restrict__ y, int batch, int h, int wd, int mid) {
    extern __shared__ float s_w[];  // mid * 27 + 2 * mid
    for (int i = threadIdx.x; i < mid * 27; i += blockDim.x) s_w[i] = w[i];
    for (int i = threadIdx.x; i < mid; i += blockDim.x) {
        s_w[mid * 27 + i] = scale[i];
        s_w[mid * 28 + i] = shift[i];
    }
    __syncthreads();
    const int oh = h / 2, ow = wd / 2;
    const int64_t total = static_cast<int64_t>(batch) * oh * ow;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % ow), oy = static_cast<int>((i / ow) % oh), n = static_cast<int>(i / (static_cast<int64_t>(ow) * oh));
        float in[27];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
                    in[c * 9 + ky * 3 + kx] = (iy >= 0 && iy < h && ix >= 0 && ix < wd)
                                                  ? __ldg(x + ((static_cast<int64_t>(n) * 3 + c) * h + iy) * wd + ix) : 0.0f;
                }
        __nv_bfloat16* o = y + i * mid;
        for (int c0 = 0; c0 < mid; c0 += 8) {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* wr = s_w + (c0 + j) * 27;
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < 27; ++k) acc = fmaf(in[k], wr[k], acc);
                r[j] = gelu_erf(fmaf(acc, s_w[mid * 27 + c0 + j], s_w[mid * 28 + c0 + j]));
            }
            *reinterpret_cast<uint4*>(o + c0) = pack8_bf16(r);
        }
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm
// 16 lanes per row of D channels (two rows per warp), kLnChunks 16-byte chunks per lane (D <= 128 * kLnChunks); fp32
// two-pass statistics like torch's layer_norm.  D = 128 / 256 / 384 keep every lane busy (one warp per row left half of
// them idle at D = 128: 1.4 TB/s).
template <int kLnChunks>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int64_t rows, int d, int x_ld, int y_ld, float eps) {
    const int sub = threadIdx.x & 15;
    const int chunks = d >> 3;
    const int64_t stride = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 4;
    const int64_t rows_pad = (rows + 1) & ~static_cast<int64_t>(1);  // both halves of a warp iterate together (shuffles)
    for (int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 4; r < rows_pad; r += stride) {
        const bool live = r < rows;
        const __nv_bfloat16* xr = x + (live ? r : 0) * x_ld;
        float v[kLnChunks][8];
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < kLnChunks; ++k) {
            const int c = sub + 16 * k;
            if (c < chunks) {
                unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(xr) + c), v[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += v[k][j];
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / static_cast<float>(d);
        float sq = 0.0f;
#pragma unroll
        for (int k = 0; k < kLnChunks; ++k)
            if (sub + 16 * k < chunks)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = v[k][j] - mean;
                    sq = fmaf(t, t, sq);
                }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / static_cast<float>(d) + eps);
        if (!live) continue;
        __nv_bfloat16* yr = y + r * y_ld;
#pragma unroll
        for (int k = 0; k < kLnChunks; ++k) {
            const int c = sub + 16 * k;
            if (c < chunks) {
                float o8[8];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) o8[j] = fmaf((v[k][j] - mean) * rstd, g[j], b[j]);
                *(reinterpret_cast<uint4*>(yr) + c) = pack8_bf16(o8);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ GELU (+ residual), in place allowed
// y = gelu(x) [+ res]: the residual form closes a LocalMixing block (x + act2(bn2(conv2(.))), svtrv2.py:62,99)
__global__ void __launch_bounds__(256) gelu_kernel(const uint4* x, const uint4* __restrict__ res, uint4* y, int64_t n8) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float f[8];
        unpack8_bf16(x[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
        if (res) {
            float r[8];
            unpack8_bf16(__ldg(res + i), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
        y[i] = pack8_bf16(f);
    }
}

// ------------------------------------------------------------------------------------------ multi-head self-attention
// qkv: [seq][n][3 * d] bf16 with channel = which * d + head * 32 + e (the reshape(B, N, 3, heads, hd) of the reference);
// out: [seq][n][d], channel = head * 32 + e.  grid = (ceil(n / 128), heads, seq); one thread = one query; the head's K and V
// live in shared memory as fp32 (n * 256 bytes); online softmax in fp32.
__global__ void __launch_bounds__(128) mhsa_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n,
                                                   int d, float scale) {
    extern __shared__ float s_kv[];  // K: [n][32], V: [n][32]
    float* s_k = s_kv;
    float* s_v = s_kv + static_cast<size_t>(n) * 32;
    const int head = blockIdx.y;
    const int64_t seq = blockIdx.z;
    const __nv_bfloat16* base = qkv + seq * n * 3 * d;
    for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {  // 8 chunks of 8 channels per token: 4 of K, 4 of V
        const int tok = i >> 3, part = i & 7;
        const int which = 1 + (part >> 2), c8 = part & 3;
        float f[8];
        unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(tok) * 3 * d + which * d + head * 32 + c8 * 8)), f);
        float* dst = (which == 1 ? s_k : s_v) + tok * 32 + c8 * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n) return;
    float q[32];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(qi) * 3 * d + head * 32 + c8 * 8)), q + c8 * 8);
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) q[e] *= scale;
    float m = -INFINITY, l = 0.0f, acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.0f;
    for (int j = 0; j < n; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(s_k + j * 32);
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 k4 = kr[c];
            s = fmaf(q[4 * c], k4.x, s); s = fmaf(q[4 * c + 1], k4.y, s); s = fmaf(q[4 * c + 2], k4.z, s); s = fmaf(q[4 * c + 3], k4.w, s);
        }
        if (s > m) {  // rescale only when the running maximum moves
            const float r = __expf(m - s);
            l *= r;
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] *= r;
            m = s;
        }
        const float p = __expf(s - m);
        l += p;
        const float4* vr = reinterpret_cast<const float4*>(s_v + j * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v4 = vr[c];
            acc[4 * c] = fmaf(p, v4.x, acc[4 * c]); acc[4 * c + 1] = fmaf(p, v4.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(p, v4.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(p, v4.w, acc[4 * c + 3]);
        }
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] *= inv;
    __nv_bfloat16* o = out + (seq * n + qi) * d + head * 32;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(o + c8 * 8) = pack8_bf16(acc + c8 * 8);
}

// ------------------------------------------------------------------------------------------ attention on the tensor cores
// Same contract as mhsa_kernel, for n a multiple of 16.  One block per (sequence, head): K ([n][32], rows padded to 80 bytes)
// and V^T ([32][n], rows padded by 16 bytes) of the head are staged once in shared memory as bf16 (both paddings make the
// fragment loads bank-conflict free); each warp owns 16 queries and walks the keys in chunks of 64 with an online softmax
// (fp32 statistics and accumulators): S = Q K^T by mma.sync.m16n8k16 (bf16, fp32 accumulate), P re-used in registers as
// the A operand of O += P V.  This is the warp-level tensor-core path; a tcgen05 / TMEM version (128 queries x 256 keys
// per instruction) is the next step -- head_dim 32 gives the P V product only N = 32 columns to work with.
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float a0, float a1, float b, float c) {
    // (d0, d1) = (a0, a1) * b + c on the packed-fp32 pipe
    uint64_t r, av, bv, cv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cv) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(av), "l"(bv), "l"(cv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(r));
}

constexpr int kKStride = 40;  // bf16 elements per K row in shared memory (32 + 8 padding)

__global__ void __launch_bounds__(256, 3) mhsa_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n,
                                                       int d, float scale_log2e) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __nv_bfloat16* s_k = reinterpret_cast<__nv_bfloat16*>(s_raw);  // [n][kKStride]
    const int vt_stride = n + 8;                                    // bf16 elements per V^T row
    __nv_bfloat16* s_vt = s_k + static_cast<size_t>(n) * kKStride;  // [32][vt_stride]
    const int head = blockIdx.x;
    const int64_t seq = blockIdx.y;
    const __nv_bfloat16* base = qkv + seq * n * 3 * d + head * 32;
    for (int i = threadIdx.x; i < n * 4; i += blockDim.x) {
        const int key = i >> 2, c8 = i & 3;
        const __nv_bfloat16* src = base + static_cast<int64_t>(key) * 3 * d + c8 * 8;
        *reinterpret_cast<uint4*>(s_k + key * kKStride + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(src + d));
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + 2 * d));
        const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_vt[(c8 * 8 + j) * vt_stride + key] = ve[j];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    for (int q0 = warp * 16; q0 < n; q0 += nwarps * 16) {
        // Q fragments: two k-steps of 16 channels
        uint32_t qa[2][4];
        const __nv_bfloat16* qr0 = base + static_cast<int64_t>(q0 + g) * 3 * d + tq * 2;
        const __nv_bfloat16* qr1 = qr0 + static_cast<int64_t>(8) * 3 * d;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qr0 + ks * 16));
            qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qr1 + ks * 16));
            qa[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qr0 + ks * 16 + 8));
            qa[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qr1 + ks * 16 + 8));
        }
        float o[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.0f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;  // rows g and g + 8
        for (int k0 = 0; k0 < n; k0 += 64) {
            const int tiles = min(8, (n - k0) >> 3);  // n % 16 == 0: an even number of 8-key tiles
            float sc[8][4];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.0f;
                if (t < tiles) {
                    // one ldmatrix.x4 = the B fragments of both k-steps of this 8-key tile: matrices at channel offsets
                    // 0 / 8 / 16 / 24, row = key (80-byte row stride: the eight 16-byte rows fall in distinct bank groups)
                    uint32_t kb[4];
                    ldmatrix_x4(kb, s_k + (k0 + t * 8 + (lane & 7)) * kKStride + (lane >> 3) * 8);
                    mma_bf16_16816(sc[t], qa[0], kb[0], kb[1]);
                    mma_bf16_16816(sc[t], qa[1], kb[2], kb[3]);
                }
            }
            // running maxima on the RAW scores (scale > 0 keeps the order); exp2(score * c - max * c) is one FFMA + one
            // MUFU.EX2 per element (ex2.approx: exp2f() costs four instructions for denormal handling nobody needs here)
            float mx0 = m0, mx1 = m1;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < tiles) {
                    mx0 = fmaxf(mx0, fmaxf(sc[t][0], sc[t][1]));
                    mx1 = fmaxf(mx1, fmaxf(sc[t][2], sc[t][3]));
                }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float r0 = ex2_approx((m0 - mx0) * scale_log2e), r1 = ex2_approx((m1 - mx1) * scale_log2e);
            m0 = mx0; m1 = mx1;
            const float nm0 = -m0 * scale_log2e, nm1 = -m1 * scale_log2e;
            l0 *= r0; l1 *= r1;
#pragma unroll
            for (int t = 0; t < 4; ++t) { o[t][0] *= r0; o[t][1] *= r0; o[t][2] *= r1; o[t][3] *= r1; }
            uint32_t pa[4][4];  // P as A fragments: one per 16-key k-step
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
                if (t < tiles) {
                    float a0, a1, a2, a3;
                    ffma2_bcast(a0, a1, sc[t][0], sc[t][1], scale_log2e, nm0);
                    ffma2_bcast(a2, a3, sc[t][2], sc[t][3], scale_log2e, nm1);
                    p0 = ex2_approx(a0); p1 = ex2_approx(a1); p2 = ex2_approx(a2); p3 = ex2_approx(a3);
                }
                l0 += p0 + p1;  // fp32 row sums; P itself is rounded to bf16 for the tensor core (relative 2^-9, unbiased)
                l1 += p2 + p3;
                pa[t >> 1][(t & 1) * 2] = pack2_bf16(p0, p1);
                pa[t >> 1][(t & 1) * 2 + 1] = pack2_bf16(p2, p3);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                if (kk * 2 < tiles) {
                    // two ldmatrix.x4 = the B fragments of the four 8-channel output tiles for this 16-key step (matrix j =
                    // channel tile j, row = channel, 16 bytes = 8 keys; low and high key halves)
                    uint32_t vlo[4], vhi[4];
                    const __nv_bfloat16* vr = s_vt + ((lane >> 3) * 8 + (lane & 7)) * vt_stride + k0 + kk * 16;
                    ldmatrix_x4(vlo, vr);
                    ldmatrix_x4(vhi, vr + 8);
#pragma unroll
                    for (int t = 0; t < 4; ++t) mma_bf16_16816(o[t], pa[kk], vlo[t], vhi[t]);
                }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        __nv_bfloat16* or0 = out + (seq * n + q0 + g) * d + head * 32 + tq * 2;
        __nv_bfloat16* or1 = or0 + static_cast<int64_t>(8) * d;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            *reinterpret_cast<uint32_t*>(or0 + t * 8) = pack2_bf16(o[t][0] * i0, o[t][1] * i0);
            *reinterpret_cast<uint32_t*>(or1 + t * 8) = pack2_bf16(o[t][2] * i1, o[t][3] * i1);
        }
    }
}

// ------------------------------------------------------------------------------------------ FRM selecting-token attention
// q: fp32 [d] (v_q(v_norm_q(select_token)): parameters only, computed at weight-preparation time); kv: [b][h][w][2 * d] bf16
// (channel = which * d + head * 32 + e); out: [b][w][d] bf16.
__global__ void __launch_bounds__(128) frm_select_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                                         __nv_bfloat16* __restrict__ out, int batch, int h, int w, int d, float scale) {
    // one thread per (b, x, head): adjacent threads take adjacent heads, so a warp reads contiguous 64-byte K (and V) rows
    const int heads = d >> 5;
    const int64_t total = static_cast<int64_t>(batch) * w * heads;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int head = static_cast<int>(i % heads);
        const int x = static_cast<int>((i / heads) % w);
        const int n = static_cast<int>(i / (static_cast<int64_t>(heads) * w));
        float qe[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(q + head * 32) + c);
            qe[4 * c] = t.x * scale; qe[4 * c + 1] = t.y * scale; qe[4 * c + 2] = t.z * scale; qe[4 * c + 3] = t.w * scale;
        }
        float m = -INFINITY, l = 0.0f, acc[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = 0.0f;
        for (int y = 0; y < h; ++y) {
            const __nv_bfloat16* t = kv + ((static_cast<int64_t>(n) * h + y) * w + x) * 2 * d + head * 32;
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float f[8];
                unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(t) + c), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s = fmaf(qe[c * 8 + j], f[j], s);
            }
            const float mn = fmaxf(m, s);
            const float r = __expf(m - mn), pr = __expf(s - mn);
            l = l * r + pr;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float f[8];
                unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(t + d) + c), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[c * 8 + j] = fmaf(pr, f[j], acc[c * 8 + j] * r);
            }
            m = mn;
        }
        const float inv = 1.0f / l;
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] *= inv;
        __nv_bfloat16* o = out + (static_cast<int64_t>(n) * w + x) * d + head * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) *(reinterpret_cast<uint4*>(o) + c) = pack8_bf16(acc + c * 8);
    }
}

// ------------------------------------------------------------------------------------------ PatchMerging row decimation
__global__ void __launch_bounds__(256) decimate_rows_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int batch, int oh,
                                                            int row_u4) {  // row_u4 = w * c / 8
    const int64_t total = static_cast<int64_t>(batch) * oh * row_u4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / row_u4;  // output row index over (batch, oh)
        const int64_t n = r / oh, yy = r % oh;
        y[i] = x[(n * 2 * oh + 2 * yy) * row_u4 + (i - r * row_u4)];
    }
}

// ------------------------------------------------------------------------------------------ log_softmax + (B,T,C)->(T,B,C)
__global__ void __launch_bounds__(256) log_softmax_tbc_kernel(const float* __restrict__ logits, float* __restrict__ out, int batch,
                                                              int t_len, int c, int ld) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = static_cast<int64_t>(batch) * t_len;
    for (int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; r < rows;
         r += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const float* x = logits + r * ld;
        float m = -INFINITY;
        for (int i = lane; i < c; i += 32) m = fmaxf(m, __ldg(x + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.0f;
        for (int i = lane; i < c; i += 32) s += expf(__ldg(x + i) - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lse = m + logf(s);
        const int64_t b = r / t_len, t = r % t_len;
        float* o = out + (t * batch + b) * c;
        for (int i = lane; i < c; i += 32) o[i] = __ldg(x + i) - lse;
    }
}

static int grid_of(int64_t items, int per_block) {
    int64_t g = (items + per_block - 1) / per_block;
    if (g > 148 * 16) g = 148 * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace b200

using namespace b200;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int b200_svtr_stem1(const float* x, const float* w, const float* scale, const float* shift, void* y, int batch, int h,
                               int wd, int mid, void* stream) {
    if (!x || !w || !scale || !shift || !y) return set_error("b200_svtr_stem1: null pointer");
    if (batch <= 0 || h <= 0 || wd <= 0 || (h & 1) || (wd & 1) || mid <= 0 || (mid % 8) || mid > 256)
        return set_error("b200_svtr_stem1: need even H, W and mid a multiple of 8 (<= 256)");
    const int64_t total = static_cast<int64_t>(batch) * (h / 2) * (wd / 2);
    svtr_stem1_kernel<<<grid_of(total, 128), 128, mid * 29 * sizeof(float), ST(stream)>>>(
        x, w, scale, shift, reinterpret_cast<__nv_bfloat16*>(y), batch, h, wd, mid);
    return check_cuda(cudaGetLastError(), "b200_svtr_stem1 launch");
}

extern "C" int b200_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int d, int x_ld, int y_ld,
                              float eps, void* stream) {
    if (!x || !y || !gamma || !beta) return set_error("b200_layernorm: null pointer");
    if (rows <= 0 || d <= 0 || (d % 8) || d > 1024 || (x_ld % 8) || (y_ld % 8) || x_ld < d || y_ld < d)
        return set_error("b200_layernorm: need D a multiple of 8 (<= 1024) and leading dimensions multiples of 8");
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
    const int grid = grid_of(rows, 16);
    if (d <= 128) layernorm_kernel<1><<<grid, 256, 0, ST(stream)>>>(xi, yo, gamma, beta, rows, d, x_ld, y_ld, eps);
    else if (d <= 256) layernorm_kernel<2><<<grid, 256, 0, ST(stream)>>>(xi, yo, gamma, beta, rows, d, x_ld, y_ld, eps);
    else if (d <= 384) layernorm_kernel<3><<<grid, 256, 0, ST(stream)>>>(xi, yo, gamma, beta, rows, d, x_ld, y_ld, eps);
    else layernorm_kernel<8><<<grid, 256, 0, ST(stream)>>>(xi, yo, gamma, beta, rows, d, x_ld, y_ld, eps);
    return check_cuda(cudaGetLastError(), "b200_layernorm launch");
}

extern "C" int b200_gelu(const void* x, const void* residual, void* y, int64_t n, void* stream) {
    if (!x || !y) return set_error("b200_gelu: null pointer");
    if (n <= 0 || (n % 8)) return set_error("b200_gelu: element count must be a positive multiple of 8");
    gelu_kernel<<<grid_of(n / 8, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(residual),
                                                            reinterpret_cast<uint4*>(y), n / 8);
    return check_cuda(cudaGetLastError(), "b200_gelu launch");
}

extern "C" int b200_mhsa(const void* qkv, void* out, int n_seq, int n, int d, float scale, void* stream) {
    if (!qkv || !out) return set_error("b200_mhsa: null pointer");
    if (n_seq <= 0 || n <= 0 || n > 768 || d <= 0 || (d % 32)) return set_error("b200_mhsa: need 1 <= N <= 768 and D a multiple of 32 (head_dim 32)");
    if (n_seq > 65535) return set_error("b200_mhsa: at most 65535 sequences per launch");
    if ((n % 16) == 0 && !getenv("B200_MHSA_SIMT")) {  // tensor-core path
        const int smem_mma = (n * kKStride + 32 * (n + 8)) * 2;
        // the opt-in is per device (and per function): set it on every call that needs it, like b200_igemm_create does --
        // a process-wide flag would leave the second GPU of a two-device process without it
        if (smem_mma > 48 * 1024 &&
            check_cuda(cudaFuncSetAttribute(mhsa_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (768 * kKStride + 32 * 776) * 2),
                       "b200_mhsa smem opt-in")) return 1;
        const int warps = n / 16 < 8 ? n / 16 : 8;  // 8 warps x 84 registers: three blocks per SM hide the mma / exp latencies
        mhsa_mma_kernel<<<dim3(d / 32, n_seq), warps * 32, smem_mma, ST(stream)>>>(
            reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), n, d, scale * 1.4426950408889634f);
        return check_cuda(cudaGetLastError(), "b200_mhsa launch");
    }
    const int smem = n * 64 * static_cast<int>(sizeof(float));
    if (smem > 48 * 1024 &&
        check_cuda(cudaFuncSetAttribute(mhsa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 768 * 64 * 4), "b200_mhsa smem opt-in")) return 1;
    const int threads = n < 128 ? ((n + 31) / 32) * 32 : 128;
    dim3 grid((n + threads - 1) / threads, d / 32, n_seq);
    mhsa_kernel<<<grid, threads, smem, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), n, d, scale);
    return check_cuda(cudaGetLastError(), "b200_mhsa launch");
}

extern "C" int b200_frm_select(const float* q, const void* kv, void* out, int batch, int h, int w, int d, float scale, void* stream) {
    if (!q || !kv || !out) return set_error("b200_frm_select: null pointer");
    if (batch <= 0 || h <= 0 || w <= 0 || d <= 0 || (d % 32)) return set_error("b200_frm_select: bad shape");
    frm_select_kernel<<<grid_of(static_cast<int64_t>(batch) * w * (d / 32), 128), 128, 0, ST(stream)>>>(
        q, reinterpret_cast<const __nv_bfloat16*>(kv), reinterpret_cast<__nv_bfloat16*>(out), batch, h, w, d, scale);
    return check_cuda(cudaGetLastError(), "b200_frm_select launch");
}

extern "C" int b200_decimate_rows(const void* x, void* y, int batch, int in_h, int w, int c, void* stream) {
    if (!x || !y) return set_error("b200_decimate_rows: null pointer");
    if (batch <= 0 || in_h <= 0 || (in_h & 1) || w <= 0 || c <= 0 || ((w * c) % 8)) return set_error("b200_decimate_rows: bad shape");
    const int row_u4 = w * c / 8;
    decimate_rows_kernel<<<grid_of(static_cast<int64_t>(batch) * (in_h / 2) * row_u4, 256), 256, 0, ST(stream)>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), batch, in_h / 2, row_u4);
    return check_cuda(cudaGetLastError(), "b200_decimate_rows launch");
}

extern "C" int b200_log_softmax_tbc(const float* logits, float* out, int batch, int t_len, int c, int ld, void* stream) {
    if (!logits || !out) return set_error("b200_log_softmax_tbc: null pointer");
    if (batch <= 0 || t_len <= 0 || c <= 0 || ld < c) return set_error("b200_log_softmax_tbc: bad shape");
    log_softmax_tbc_kernel<<<grid_of(static_cast<int64_t>(batch) * t_len, 8), 256, 0, ST(stream)>>>(logits, out, batch, t_len, c, ld);
    return check_cuda(cudaGetLastError(), "b200_log_softmax_tbc launch");
}
