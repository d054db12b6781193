// Implicit-GEMM convolution family for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed operands,
// warp-specialised persistent CTAs.  One kernel template covers
//   * plain 1x1 / 3x3 convolutions, stride 1 or 2 (A operand = NHWC activations loaded by TMA boxes,
//     one box per (tap, 64-channel slice); zero padding comes from TMA out-of-bounds fill),
//   * the modulated deformable conv v2 (A operand gathered by producer warps: 4-corner bilinear blend,
//     mask modulation, bf16 pack, written straight into the 128B-swizzled UMMA layout),
//   * the 7x7 s2 stem on a space-to-depth buffer (A rows are contiguous 128 B runs, copied by the
//     producer warps),
// all with a fused epilogue (folded BN scale/shift or bias, residual add, ReLU, bf16 / fp32 store,
// 2x2 pixel-shuffle store for ConvTranspose2d 2x2 s2).
//
// GEMM view:  D[M = output pixels of one 128-pixel tile][N = out channels] =
//             A[M][K = taps * Cin] * B[N][K]^T,   K ordered (tap, cin),  BLOCK_K = 64 (one SW128 atom).
//
// Reference call sites replaced: see include/dbnet_b200.h.
#include "common.cuh"
#include "internal.h"

#include <cstdio>
#include <cstring>
#include <new>

namespace b200 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB per stage
constexpr int kGatherWarps = 8;                 // producer warps of the gather (DCN / stem) variant
constexpr int kBaseWarps = 6;                   // TMA, MMA, 4 epilogue
constexpr int kMaxStages = 8;
constexpr int kMaxSmem = 227 * 1024;

struct IgemmParams {
    int mode, batch, in_h, in_w, in_c, out_h, out_w, out_c;
    int ksize, stride, flags, block_n;
    int tile_w, tile_h, tiles_x, tiles_y;
    int num_n_blocks, num_tiles, k_blocks, cchunks;
    int out_ld, res_ld, shuf_c, stages;
    int pk_h, pk_w;  // STEM: rows / columns of the packed space-to-depth buffer
    const __nv_bfloat16* x;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    void* y;
    const float* offsets;
};

struct TmapPack {
    CUtensorMap a[4];
    CUtensorMap b;
};

struct SmemLayout {
    uint8_t* a;
    uint8_t* b;
    uint64_t* full;
    uint64_t* empty;
    uint64_t* tfull;
    uint64_t* tempty;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ SmemLayout carve(uint8_t* raw, int stages, int b_bytes) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    SmemLayout s;
    s.a = base;
    s.b = base + stages * kABytes;
    s.full = reinterpret_cast<uint64_t*>(s.b + stages * b_bytes);
    s.empty = s.full + kMaxStages;
    s.tfull = s.empty + kMaxStages;
    s.tempty = s.tfull + 2;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.tempty + 2);
    return s;
}

struct TileCoord {
    int n, y0, x0, nb;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
    TileCoord t;
    t.nb = tile % p.num_n_blocks;
    int m = tile / p.num_n_blocks;
    const int per_img = p.tiles_x * p.tiles_y;
    t.n = m / per_img;
    m -= t.n * per_img;
    const int ty = m / p.tiles_x;
    t.y0 = ty * p.tile_h;
    t.x0 = (m - ty * p.tiles_x) * p.tile_w;
    return t;
}

// ------------------------------------------------------------------------------------------ epilogue
// v[32] = fp32 accumulators of one pixel (TMEM lane) for 32 consecutive GEMM columns starting at col0.
__device__ __forceinline__ void epilogue_chunk(const IgemmParams& p, const uint32_t (&v)[32], int col0, int n, int oy,
                                               int ox) {
    if (col0 >= p.out_c) return;
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0 + i));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0 + i));
        f[i + 0] = fmaf(__uint_as_float(v[i + 0]), sc.x, sh.x);
        f[i + 1] = fmaf(__uint_as_float(v[i + 1]), sc.y, sh.y);
        f[i + 2] = fmaf(__uint_as_float(v[i + 2]), sc.z, sh.z);
        f[i + 3] = fmaf(__uint_as_float(v[i + 3]), sc.w, sh.w);
    }
    const int ncols = min(32, p.out_c - col0);  // multiple of 8
    int64_t opix;
    int ch0 = col0;
    if (p.flags & B200_EPI_SHUFFLE2) {
        const int sub = col0 / p.shuf_c;  // dy*2+dx ; a 32-column chunk never straddles two sub-pixels
        ch0 = col0 - sub * p.shuf_c;
        opix = (static_cast<int64_t>(n) * (2 * p.out_h) + (2 * oy + (sub >> 1))) * (2 * p.out_w) + (2 * ox + (sub & 1));
    } else {
        opix = (static_cast<int64_t>(n) * p.out_h + oy) * p.out_w + ox;
    }
    if (p.flags & (B200_EPI_RES | B200_EPI_RES_UP2)) {
        int64_t rpix;
        if (p.flags & B200_EPI_RES_UP2)
            rpix = (static_cast<int64_t>(n) * (p.out_h >> 1) + (oy >> 1)) * (p.out_w >> 1) + (ox >> 1);
        else
            rpix = opix;
        const __nv_bfloat16* r = p.residual + rpix * p.res_ld + ch0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            if (i < ncols) {
                const uint4 u = ldg128(r + i);
                f[i + 0] += bf16lo(u.x); f[i + 1] += bf16hi(u.x);
                f[i + 2] += bf16lo(u.y); f[i + 3] += bf16hi(u.y);
                f[i + 4] += bf16lo(u.z); f[i + 5] += bf16hi(u.z);
                f[i + 6] += bf16lo(u.w); f[i + 7] += bf16hi(u.w);
            }
        }
    }
    if (p.flags & B200_EPI_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
    }
    if (p.flags & B200_EPI_OUT_F32) {
        float* o = reinterpret_cast<float*>(p.y) + opix * p.out_ld + ch0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            if (i < ncols) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
    } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.y) + opix * p.out_ld + ch0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            if (i < ncols) {
                uint4 u;
                u.x = pack_bf16(f[i + 0], f[i + 1]);
                u.y = pack_bf16(f[i + 2], f[i + 3]);
                u.z = pack_bf16(f[i + 4], f[i + 5]);
                u.w = pack_bf16(f[i + 6], f[i + 7]);
                *reinterpret_cast<uint4*>(o + i) = u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel
template <bool kGather>
__global__ void __launch_bounds__((kBaseWarps + (kGather ? kGatherWarps : 0)) * 32, 1)
igemm_kernel(const __grid_constant__ TmapPack tm, const IgemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b_bytes = p.block_n * 128;
    const SmemLayout sm = carve(smem_raw, p.stages, b_bytes);
    const int stages = p.stages;
    const uint32_t tmem_cols = (2 * p.block_n < 32) ? 32u : static_cast<uint32_t>(2 * p.block_n);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm.b);
        if (!kGather) {
            tma_prefetch_desc(&tm.a[0]);
            if (p.stride == 2 && p.ksize == 3) {
                tma_prefetch_desc(&tm.a[1]);
                tma_prefetch_desc(&tm.a[2]);
                tma_prefetch_desc(&tm.a[3]);
            }
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&sm.full[s], kGather ? 1 + kGatherWarps : 1);
                mbar_init(&sm.empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&sm.tfull[a], 1);
                mbar_init(&sm.tempty[a], 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(sm.tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const TileCoord t = decode_tile(p, tile);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&sm.full[s], (kGather ? 0 : kABytes) + b_bytes);
                    if (!kGather) {
                        const int tap = kb / p.cchunks;
                        const int c0 = (kb - tap * p.cchunks) * kBlockK;
                        int ax = t.x0, ay = t.y0;
                        const CUtensorMap* ma = &tm.a[0];
                        if (p.ksize == 3) {
                            const int ky = tap / 3, kx = tap - ky * 3;
                            if (p.stride == 1) {
                                ax += kx - 1;
                                ay += ky - 1;
                            } else {
                                // input pixel (2*oy + ky - 1, 2*ox + kx - 1) on the (row parity, column parity) view
                                ma = &tm.a[(((ky + 1) & 1) << 1) | ((kx + 1) & 1)];
                                ax += (kx == 0) ? -1 : 0;
                                ay += (ky == 0) ? -1 : 0;
                            }
                        }
                        tma_load_4d(sm.a + s * kABytes, ma, &sm.full[s], c0, ax, ay, t.n);
                    }
                    tma_load_2d(sm.b + s * b_bytes, &tm.b, &sm.full[s], kb * kBlockK, t.nb * p.block_n);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(static_cast<uint32_t>(p.block_n));
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&sm.tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.block_n);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&sm.full[s], ph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(sm.a + s * kABytes));
                    const uint64_t db = umma_desc_sw128(smem_u32(sm.b + s * b_bytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // +32 bytes per UMMA_K step inside the 128B swizzle atom (encoded >> 4)
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&sm.empty[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(&sm.tfull[as]);
            }
        }
    } else if (warp < kBaseWarps) {
        // ===================================================================== epilogue (4 warps)
        const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const int ty = row / p.tile_w;
        const int tx = row - ty * p.tile_w;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const TileCoord t = decode_tile(p, tile);
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int oy = t.y0 + ty, ox = t.x0 + tx;
            const bool valid = (oy < p.out_h) && (ox < p.out_w);
            mbar_wait(&sm.tfull[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as * p.block_n);
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c, v);
                tmem_ld_wait();
                if (valid) epilogue_chunk(p, v, t.nb * p.block_n + c, t.n, oy, ox);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tempty[as]);
        }
    } else if (kGather) {
        // ===================================================================== gather producers
        constexpr int kRowsPerWarp = kBlockM / kGatherWarps;  // 16
        constexpr int kPasses = kRowsPerWarp / 4;             // 4 rows per pass (8 lanes x 16 B = one row)
        const int pw = warp - kBaseWarps;
        const int sub = lane >> 3;
        const int l8 = lane & 7;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(p, tile);
            int oy[kPasses], ox[kPasses];
            bool valid[kPasses];
            uint32_t soff[kPasses];
#pragma unroll
            for (int q = 0; q < kPasses; ++q) {
                const int r = pw * kRowsPerWarp + q * 4 + sub;
                const int ty = r / p.tile_w;
                oy[q] = t.y0 + ty;
                ox[q] = t.x0 + (r - ty * p.tile_w);
                valid[q] = (oy[q] < p.out_h) && (ox[q] < p.out_w);
                soff[q] = static_cast<uint32_t>(r * 128 + ((l8 ^ (r & 7)) << 4));
            }
            if (p.mode == B200_IGEMM_STEM) {
                // A row for (pixel, tap row ty) = 64 contiguous bf16 of the packed buffer
                const __nv_bfloat16* img = p.x + static_cast<int64_t>(t.n) * p.pk_h * p.pk_w * 16;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    uint4 u[kPasses];
#pragma unroll
                    for (int q = 0; q < kPasses; ++q) {
                        u[q] = make_uint4(0, 0, 0, 0);
                        if (valid[q])
                            u[q] = ldg128(img + (static_cast<int64_t>(oy[q] + kb) * p.pk_w + ox[q]) * 16 + l8 * 8);
                    }
                    mbar_wait(&sm.empty[s], ph ^ 1);
                    uint8_t* a = sm.a + s * kABytes;
#pragma unroll
                    for (int q = 0; q < kPasses; ++q) *reinterpret_cast<uint4*>(a + soff[q]) = u[q];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.full[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            } else {
                // modulated deformable sampling (torchvision deform_conv2d semantics, model/det/dcn.py:48-57)
                const __nv_bfloat16* img = p.x + static_cast<int64_t>(t.n) * p.in_h * p.in_w * p.in_c + l8 * 8;
                const float* offp[kPasses];
#pragma unroll
                for (int q = 0; q < kPasses; ++q)
                    offp[q] = p.offsets + ((static_cast<int64_t>(t.n) * p.out_h + oy[q]) * p.out_w + ox[q]) * 32;
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    int o1[kPasses], o2[kPasses], o3[kPasses], o4[kPasses];
                    float w1[kPasses], w2[kPasses], w3[kPasses], w4[kPasses];
#pragma unroll
                    for (int q = 0; q < kPasses; ++q) {
                        o1[q] = o2[q] = o3[q] = o4[q] = 0;
                        w1[q] = w2[q] = w3[q] = w4[q] = 0.0f;
                        if (valid[q]) {
                            const float dy = __ldg(offp[q] + 2 * tap);
                            const float dx = __ldg(offp[q] + 2 * tap + 1);
                            const float ml = __ldg(offp[q] + 18 + tap);
                            const float m = 1.0f / (1.0f + expf(-ml));
                            const float h = static_cast<float>(oy[q] * p.stride - 1 + ky) + dy;
                            const float w = static_cast<float>(ox[q] * p.stride - 1 + kx) + dx;
                            if (h > -1.0f && w > -1.0f && h < static_cast<float>(p.in_h) &&
                                w < static_cast<float>(p.in_w)) {
                                const float hf = floorf(h), wf = floorf(w);
                                const int hl = static_cast<int>(hf), wl = static_cast<int>(wf);
                                const int hh = hl + 1, wh = wl + 1;
                                const float lh = h - hf, lw = w - wf;
                                const float uh = 1.0f - lh, uw = 1.0f - lw;
                                const bool t_ok = hl >= 0, b_ok = hh <= p.in_h - 1;
                                const bool l_ok = wl >= 0, r_ok = wh <= p.in_w - 1;
                                if (t_ok && l_ok) { o1[q] = (hl * p.in_w + wl) * p.in_c; w1[q] = m * (uh * uw); }
                                if (t_ok && r_ok) { o2[q] = (hl * p.in_w + wh) * p.in_c; w2[q] = m * (uh * lw); }
                                if (b_ok && l_ok) { o3[q] = (hh * p.in_w + wl) * p.in_c; w3[q] = m * (lh * uw); }
                                if (b_ok && r_ok) { o4[q] = (hh * p.in_w + wh) * p.in_c; w4[q] = m * (lh * lw); }
                            }
                        }
                    }
                    for (int cc = 0; cc < p.cchunks; ++cc) {
                        const __nv_bfloat16* src = img + cc * kBlockK;
                        uint4 res[kPasses];
#pragma unroll
                        for (int q = 0; q < kPasses; ++q) {
                            const uint4 a = ldg128(src + o1[q]);
                            const uint4 b = ldg128(src + o2[q]);
                            const uint4 c = ldg128(src + o3[q]);
                            const uint4 d = ldg128(src + o4[q]);
                            const float A = w1[q], B = w2[q], C = w3[q], D = w4[q];
#define B200_BLEND(f) pack_bf16(A * bf16lo(a.f) + B * bf16lo(b.f) + C * bf16lo(c.f) + D * bf16lo(d.f), \
                                A * bf16hi(a.f) + B * bf16hi(b.f) + C * bf16hi(c.f) + D * bf16hi(d.f))
                            res[q].x = B200_BLEND(x);
                            res[q].y = B200_BLEND(y);
                            res[q].z = B200_BLEND(z);
                            res[q].w = B200_BLEND(w);
#undef B200_BLEND
                        }
                        mbar_wait(&sm.empty[s], ph ^ 1);
                        uint8_t* a = sm.a + s * kABytes;
#pragma unroll
                        for (int q = 0; q < kPasses; ++q) *reinterpret_cast<uint4*>(a + soff[q]) = res[q];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.full[s]);
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const char* what) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

}  // namespace b200

using namespace b200;

struct b200_igemm {
    TmapPack tm;
    IgemmParams p;
    int gather;
    int grid;
    int threads;
    int smem_bytes;
};

static int pick_tile_w(int out_h, int out_w) {
    int best = 128;
    long best_area = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
        const int th = 128 / tw;
        const long area = (long)((out_w + tw - 1) / tw) * tw * (long)((out_h + th - 1) / th) * th;
        if (best_area < 0 || area < best_area) {
            best_area = area;
            best = tw;
        }
    }
    return best;
}

extern "C" int b200_igemm_create(const b200_igemm_desc* d, b200_igemm** out) {
    if (!d || !out) return set_error("b200_igemm_create: null argument");
    *out = nullptr;
    if (d->mode < B200_IGEMM_CONV || d->mode > B200_IGEMM_DCN) return set_error("igemm: bad mode %d", d->mode);
    if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0) return set_error("igemm: bad shape");
    if (!d->x || !d->w || !d->scale || !d->shift || !d->y) return set_error("igemm: null tensor pointer");
    if (d->out_c <= 0 || (d->out_c % 8) != 0) return set_error("igemm: out_c must be a positive multiple of 8");
    if (d->stride != 1 && d->stride != 2) return set_error("igemm: stride must be 1 or 2");
    int taps;
    if (d->mode == B200_IGEMM_STEM) {
        if (d->ksize != 7 || d->stride != 2 || (d->in_h & 1) || (d->in_w & 1))
            return set_error("igemm: STEM is 7x7 stride 2 on even-sized images");
        taps = 4;
    } else {
        if (d->ksize != 1 && d->ksize != 3) return set_error("igemm: ksize must be 1 or 3");
        if (d->mode == B200_IGEMM_DCN && d->ksize != 3) return set_error("igemm: DCN is 3x3");
        if (d->in_c <= 0 || (d->in_c % 64) != 0) return set_error("igemm: in_c must be a multiple of 64");
        if (d->stride == 2 && ((d->in_h & 1) || (d->in_w & 1)))
            return set_error("igemm: stride 2 needs even input size");
        taps = d->ksize * d->ksize;
    }
    if (d->mode == B200_IGEMM_DCN && !d->offsets) return set_error("igemm: DCN needs offsets");
    if ((d->flags & (B200_EPI_RES | B200_EPI_RES_UP2)) && !d->residual) return set_error("igemm: residual missing");
    int block_n = d->block_n;
    if (block_n == 0) {
        block_n = 256;
        while (block_n > 32 && block_n / 2 >= d->n_pad) block_n >>= 1;
        if (d->n_pad % block_n) block_n = 32;
    }
    if (block_n != 32 && block_n != 64 && block_n != 128 && block_n != 256)
        return set_error("igemm: block_n must be 32, 64, 128 or 256");
    if (d->n_pad <= 0 || (d->n_pad % block_n) != 0) return set_error("igemm: n_pad must be a multiple of block_n");
    if (d->out_c > d->n_pad) return set_error("igemm: out_c > n_pad");
    if ((d->flags & B200_EPI_SHUFFLE2) && (d->shuf_c <= 0 || d->shuf_c % 32 != 0 || d->out_c != 4 * d->shuf_c))
        return set_error("igemm: SHUFFLE2 needs out_c == 4*shuf_c, shuf_c multiple of 32");
    if ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->w) | reinterpret_cast<uintptr_t>(d->y) |
         reinterpret_cast<uintptr_t>(d->residual) | reinterpret_cast<uintptr_t>(d->scale) |
         reinterpret_cast<uintptr_t>(d->shift) | reinterpret_cast<uintptr_t>(d->offsets)) & 15)
        return set_error("igemm: all tensor pointers must be 16-byte aligned");
    if ((d->out_ld % 8) || ((d->flags & (B200_EPI_RES | B200_EPI_RES_UP2)) && (d->res_ld % 8)))
        return set_error("igemm: out_ld / res_ld must be multiples of 8");

    b200_igemm* op = new (std::nothrow) b200_igemm;
    if (!op) return set_error("igemm: out of host memory");
    memset(op, 0, sizeof(*op));
    IgemmParams& p = op->p;
    p.mode = d->mode;
    p.batch = d->batch;
    p.in_h = d->in_h;
    p.in_w = d->in_w;
    p.in_c = d->in_c;
    p.out_h = d->in_h / d->stride;
    p.out_w = d->in_w / d->stride;
    if (d->stride == 1) { p.out_h = d->in_h; p.out_w = d->in_w; }
    p.out_c = d->out_c;
    p.ksize = d->ksize;
    p.stride = d->stride;
    p.flags = d->flags;
    p.block_n = block_n;
    p.tile_w = pick_tile_w(p.out_h, p.out_w);
    p.tile_h = kBlockM / p.tile_w;
    p.tiles_x = (p.out_w + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (p.out_h + p.tile_h - 1) / p.tile_h;
    p.num_n_blocks = d->n_pad / block_n;
    p.num_tiles = p.batch * p.tiles_x * p.tiles_y * p.num_n_blocks;
    p.cchunks = (d->mode == B200_IGEMM_STEM) ? 1 : d->in_c / kBlockK;
    p.k_blocks = taps * p.cchunks;
    p.out_ld = d->out_ld;
    p.res_ld = d->res_ld;
    p.shuf_c = d->shuf_c;
    p.pk_h = d->in_h / 2 + 3;
    p.pk_w = d->in_w / 2 + 3;
    p.x = reinterpret_cast<const __nv_bfloat16*>(d->x);
    p.scale = d->scale;
    p.shift = d->shift;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
    p.y = d->y;
    p.offsets = d->offsets;
    op->gather = (d->mode != B200_IGEMM_CONV);

    const int b_bytes = block_n * 128;
    const int budget = kMaxSmem - 1024 /*alignment*/ - 256 /*barriers*/;
    int stages = budget / (kABytes + b_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    p.stages = stages;
    op->smem_bytes = stages * (kABytes + b_bytes) + 1024 + 256;
    op->threads = (kBaseWarps + (op->gather ? kGatherWarps : 0)) * 32;

    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        delete op;
        return set_error("igemm: cannot query device: %s", cudaGetErrorString(cudaGetLastError()));
    }
    op->grid = p.num_tiles < sms ? p.num_tiles : sms;

    // ---- tensor maps
    const int ktot = p.k_blocks * kBlockK;
    {
        const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->n_pad};
        const uint64_t str[1] = {(uint64_t)ktot * 2};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)block_n};
        if (encode_map(&op->tm.b, d->w, 2, dims, str, box, "weights")) { delete op; return 1; }
    }
    if (!op->gather) {
        const uint64_t C = d->in_c, W = d->in_w, H = d->in_h;
        const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1u};
        if (d->stride == 1) {
            const uint64_t dims[4] = {C, W, H, (uint64_t)d->batch};
            const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
            if (encode_map(&op->tm.a[0], d->x, 4, dims, str, box, "activations")) { delete op; return 1; }
        } else {
            // stride 2: four (row parity, column parity) strided views of the NHWC tensor
            const uint64_t dims[4] = {C, W / 2, H / 2, (uint64_t)d->batch};
            const uint64_t str[3] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2};
            const int nviews = (d->ksize == 3) ? 4 : 1;
            for (int v = 0; v < nviews; ++v) {
                const int py = v >> 1, px = v & 1;
                const char* base = reinterpret_cast<const char*>(d->x) + ((uint64_t)py * W + px) * C * 2;
                if (encode_map(&op->tm.a[v], base, 4, dims, str, box, "strided activations")) { delete op; return 1; }
            }
        }
    }
    // the attribute is per kernel function, not per op: always raise it to the full 227 KB opt-in
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) {
        delete op;
        return set_error("igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    }
    *out = op;
    return 0;
}

extern "C" int b200_igemm_launch(const b200_igemm* op, void* stream) {
    if (!op) return set_error("b200_igemm_launch: null op");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (op->gather)
        igemm_kernel<true><<<op->grid, op->threads, op->smem_bytes, st>>>(op->tm, op->p);
    else
        igemm_kernel<false><<<op->grid, op->threads, op->smem_bytes, st>>>(op->tm, op->p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("igemm launch failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int b200_igemm_destroy(b200_igemm* op) {
    delete op;
    return 0;
}

extern "C" int b200_igemm_info(const b200_igemm* op, int32_t* out_h, int32_t* out_w, int32_t* tile_w, int32_t* tile_h,
                               int32_t* block_n, int32_t* num_tiles, int32_t* smem_bytes) {
    if (!op) return set_error("b200_igemm_info: null op");
    if (out_h) *out_h = op->p.out_h;
    if (out_w) *out_w = op->p.out_w;
    if (tile_w) *tile_w = op->p.tile_w;
    if (tile_h) *tile_h = op->p.tile_h;
    if (block_n) *block_n = op->p.block_n;
    if (num_tiles) *num_tiles = op->p.num_tiles;
    if (smem_bytes) *smem_bytes = op->smem_bytes;
    return 0;
}
