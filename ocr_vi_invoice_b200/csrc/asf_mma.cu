// Adaptive scale fusion (ScaleFeatureSelection.forward, model/det/neck.py:57-79) with the bilinear upsampling AND the
// attention-weighted sum done by the tensor cores.
//
//   out[px, c] = s0[px] * p2[px, c] + sum_{l = 3,4,5} s_l[px] * bilinear_align_corners(p_l)[px, c]
//              = s0[px] * p2[px, c] + sum_q S[px, q] * P[q, c]
//
// where q runs over the source pixels of the three low-resolution levels inside the window an output tile (16 x 8 pixels)
// can touch (at most 10 x 6, 6 x 4 and 4 x 3 pixels, padded to K = 64 + 32 + 16), P[q, :] are their 256 channels and S is
// a 128 x 112 matrix with twelve non-zeros per row: softmax weight x bilinear corner weight, rounded to bf16.  So one
// output tile is ONE 128 x 256 x 112 UMMA chain (7 instructions) instead of 12 corner loads + 12 x 32 FMAs per thread:
// the CUDA-core kernel (asf_combine_kernel) needs ~900 issue slots and 56 L1 wavefronts per pixel and runs at 2.6 TB/s;
// this one is bound by the HBM traffic of reading p2 and writing the fused map.
//   * A operand = S, built per tile in shared memory (K-major, 128B swizzle) by four "builder" warps, one output pixel per
//     thread: interpolate the per-level logit contributions z_l = W_l p_l (computed at native resolution by four tiny GEMMs,
//     see b200_asf_combine), softmax over the four levels, scatter the 12 weights.
//   * B operand = the level windows exactly as TMA lands them: rows = source pixels (K), 64 channels = 128 contiguous bytes
//     per row, i.e. an MN-major operand (instruction-descriptor bit 16; tools/mnmajor_probe.cu pins LBO / SBO).  No
//     transposition anywhere.
//   * epilogue: acc + s0 * p2 (the p2 tile arrives by TMA in the very staging buffer the result overwrites), bf16, TMA store.
// Warp roles (16 warps): 0 window loader, 1 MMA issuer, 2 epilogue DMA (p2 loads + output stores), 3 idle, 4-7 builders,
// 8-15 two epilogue groups.  TMEM: two 256-column accumulators, so tile i+1's MMAs overlap tile i's epilogue.
#include "common.cuh"
#include "internal.h"
#include "dcn2.h"  // encode_map_typed

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace b200 {

namespace {

constexpr int kTileW = 16, kTileH = 8;
// source-pixel windows of levels 3, 4, 5 (see the window bound in b200_asf_plan_create); constexpr functions, not arrays:
// device code cannot index a host constexpr array
__host__ __device__ constexpr int win_w(int l) { return l == 0 ? 10 : (l == 1 ? 6 : 4); }
__host__ __device__ constexpr int win_h(int l) { return l == 0 ? 6 : (l == 1 ? 4 : 3); }
__host__ __device__ constexpr int win_q(int l) { return l == 0 ? 64 : (l == 1 ? 32 : 16); }     // padded to a multiple of 16 (UMMA K)
__host__ __device__ constexpr int k_off(int l) { return l == 0 ? 0 : (l == 1 ? 64 : 96); }       // K offset of the level inside S
__host__ __device__ constexpr int lvl_off(int l) {                                               // byte offset inside a window buffer
    return l == 0 ? 0 : (l == 1 ? 4 * 64 * 128 : 4 * 64 * 128 + 4 * 32 * 128);
}
constexpr int kWinBytes = 4 * (64 + 32 + 16) * 128;          // 56 KB: four 64-channel chunks of every level
constexpr int kWinTx = 4 * (10 * 6 + 6 * 4 + 4 * 3) * 128;   // bytes the 12 boxes deliver
constexpr int kSBytes = 2 * 16384;                           // S: two k-blocks of 64 (the second holds K = 64..111)
constexpr int kChunkBytes = 16384;
constexpr int kEpiBufs = 4;  // 16 KB each; five would overshoot the 227 KB opt-in by 512 bytes
constexpr int kThreads = 16 * 32;
constexpr int kSmemBytes = 1024 + 2 * kWinBytes + kSBytes + kEpiBufs * kChunkBytes + 4 * 128 * 4 + 512;

struct AsfLevelM {
    const float* z;
    int h, w;
    float sy, sx;
};
struct AsfMmaParams {
    AsfLevelM lv[4];
    const float* b_att;
    int batch, zld, h2, w2, tiles_x, tiles_y, num_tiles;
};
struct AsfMmaTmaps {
    CUtensorMap win[3];  // levels 3..5: box {64, kWinW, kWinH, 1}
    CUtensorMap p2;      // box {64, 16, 8, 1}
    CUtensorMap out;     // box {64, 16, 8, 1}
};

struct Smem {
    uint8_t* win;   // [2][kWinBytes]
    uint8_t* s;     // [kSBytes]
    uint8_t* epi;   // [kEpiBufs][16 KB]
    float* s0;      // [4][128]
    uint64_t* win_full;   // [2]
    uint64_t* win_empty;  // [2]
    uint64_t* s_full;     // [1]
    uint64_t* s_empty;    // [1]
    uint64_t* tfull;      // [2]
    uint64_t* tempty;     // [2]
    uint64_t* res_full;   // [kEpiBufs]
    uint64_t* out_ready;  // [kEpiBufs]
    uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* raw) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    Smem s;
    s.win = base;
    s.s = s.win + 2 * kWinBytes;
    s.epi = s.s + kSBytes;
    s.s0 = reinterpret_cast<float*>(s.epi + kEpiBufs * kChunkBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.s0 + 4 * 128);
    s.win_full = bars;
    s.win_empty = bars + 2;
    s.s_full = bars + 4;
    s.s_empty = bars + 5;
    s.tfull = bars + 6;
    s.tempty = bars + 8;
    s.res_full = bars + 10;
    s.out_ready = s.res_full + kEpiBufs;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.out_ready + kEpiBufs);
    return s;
}

struct Tile {
    int n, y0, x0;
};
__device__ __forceinline__ Tile decode(const AsfMmaParams& p, int tile) {
    Tile t;
    const int per_img = p.tiles_x * p.tiles_y;
    t.n = tile / per_img;
    const int m = tile - t.n * per_img;
    const int ty = m / p.tiles_x;
    t.y0 = ty * kTileH;
    t.x0 = (m - ty * p.tiles_x) * kTileW;
    return t;
}

// MN-major shared-memory operand, 128-byte swizzle: rows = K, 64 elements of N contiguous (128 B) per row;
// LBO = bytes between consecutive 64-element N groups, SBO = bytes between groups of 8 K rows
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ void sts_b16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) asf_mma_kernel(const __grid_constant__ AsfMmaTmaps tm, const AsfMmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const Smem sm = carve(smem_raw);
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm.win[0]);
        tma_prefetch_desc(&tm.win[1]);
        tma_prefetch_desc(&tm.win[2]);
        tma_prefetch_desc(&tm.p2);
        tma_prefetch_desc(&tm.out);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int a = 0; a < 2; ++a) {
                mbar_init(&sm.win_full[a], 1);
                mbar_init(&sm.win_empty[a], 1);
                mbar_init(&sm.tfull[a], 1);
                mbar_init(&sm.tempty[a], 8);
            }
            mbar_init(sm.s_full, 4);
            mbar_init(sm.s_empty, 1);
            for (int a = 0; a < kEpiBufs; ++a) {
                mbar_init(&sm.res_full[a], 1);
                mbar_init(&sm.out_ready[a], 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(sm.tmem_slot, 512);
        tmem_relinquish();
    }
    // window rows past the boxes (K padding) and S must be finite: zero everything the MMAs can read
    for (int i = threadIdx.x; i < (2 * kWinBytes + kSBytes) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sm.win)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ===================================================================== window loader
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            mbar_wait(&sm.win_empty[buf], ((it >> 1) & 1) ^ 1);
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x));
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.win_full[buf], kWinTx);
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const int wy = static_cast<int>(p.lv[l + 1].sy * static_cast<float>(t.y0));
                    const int wx = static_cast<int>(p.lv[l + 1].sx * static_cast<float>(t.x0));
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        tma_load_4d(sm.win + buf * kWinBytes + lvl_off(l) + c * win_q(l) * 128, &tm.win[l], &sm.win_full[buf], c * 64, wx,
                                    wy, t.n);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        const uint32_t idesc = umma_idesc_bf16(256u) | (1u << 16);  // B operand MN-major
        const uint32_t s0a = smem_u32(sm.s), w0a = smem_u32(sm.win);
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            mbar_wait(&sm.tempty[buf], ((it >> 1) & 1) ^ 1);
            mbar_wait(&sm.win_full[buf], (it >> 1) & 1);
            mbar_wait(sm.s_full, it & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * 256);
            const uint32_t wb = w0a + buf * kWinBytes;
            if (elect_one()) {
                // K = 0..63: level 3 (4 steps of 16 source pixels = 2 KB of window rows each)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(d_tmem, umma_desc_sw128(s0a) + 2 * k, umma_desc_mn_sw128(wb + lvl_off(0) + k * 2048, win_q(0) * 128, 1024), idesc,
                              k ? 1u : 0u);
                // K = 64..95: level 4, K = 96..111: level 5 (second k-block of S)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    umma_bf16(d_tmem, umma_desc_sw128(s0a + 16384) + 2 * k, umma_desc_mn_sw128(wb + lvl_off(1) + k * 2048, win_q(1) * 128, 1024),
                              idesc, 1u);
                umma_bf16(d_tmem, umma_desc_sw128(s0a + 16384) + 4, umma_desc_mn_sw128(wb + lvl_off(2), win_q(2) * 128, 1024), idesc, 1u);
                umma_commit(&sm.win_empty[buf]);
                umma_commit(sm.s_empty);
                umma_commit(&sm.tfull[buf]);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ===================================================================== epilogue DMA: p2 chunk loads + output stores
        // chunk g (4 per tile) uses buffer g % B; after the store of chunk g is queued the store of chunk g-1 has read its
        // buffer (wait_group.read 1), which is then refilled with the p2 chunk of g-1+B
        const int B = kEpiBufs;
        int ld_it = 0, ld_chunk = 0;
        auto issue_load = [&](int slot) {
            if (ld_it >= my_tiles) return;
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + ld_it * static_cast<int>(gridDim.x));
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.res_full[slot], kChunkBytes);
                tma_load_4d(sm.epi + slot * kChunkBytes, &tm.p2, &sm.res_full[slot], ld_chunk * 64, t.x0, t.y0, t.n);
            }
            __syncwarp();
            if (++ld_chunk == 4) { ld_chunk = 0; ++ld_it; }
        };
        for (int i = 0; i < B; ++i) issue_load(i);
        int g = 0, slot = 0, rslot = 0;
        uint32_t par = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x));
            for (int c = 0; c < 4; ++c, ++g) {
                mbar_wait(&sm.out_ready[slot], par);
                if (elect_one()) {
                    tma_store_4d(&tm.out, sm.epi + slot * kChunkBytes, c * 64, t.x0, t.y0, t.n);
                    bulk_commit();
                    if (g >= 1) bulk_wait_read1();
                }
                __syncwarp();
                if (g >= 1) {
                    issue_load(rslot);
                    if (++rslot == B) rslot = 0;
                }
                if (++slot == B) { slot = 0; par ^= 1; }
            }
        }
        if (elect_one()) bulk_wait_all();
        __syncwarp();
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== builders: one output pixel per thread
        const int r = (warp - 4) * 32 + lane;
        const int ty = r >> 4, tx = r & 15;
        const uint32_t srow = smem_u32(sm.s) + static_cast<uint32_t>(r) * 128;
        const uint32_t swz = static_cast<uint32_t>(r & 7);
        const float b0 = __ldg(p.b_att), b1 = __ldg(p.b_att + 1), b2 = __ldg(p.b_att + 2), b3 = __ldg(p.b_att + 3);
        for (int it = 0; it < my_tiles; ++it) {
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x));
            const int y = t.y0 + ty, x = t.x0 + tx;
            const bool valid = (y < p.h2) && (x < p.w2);
            int kidx[3][4];
            float cw[3][4];
            float4 lg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) lg = __ldg(reinterpret_cast<const float4*>(p.lv[0].z + (static_cast<int64_t>(t.n) * p.h2 + y) * p.w2 * p.zld + static_cast<int64_t>(x) * p.zld));
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const AsfLevelM& L = p.lv[l + 1];
                // same float arithmetic as asf_combine_kernel (and as the loader's window origin)
                const float fy = L.sy * static_cast<float>(y), fx = L.sx * static_cast<float>(x);
                const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
                const int y1 = y0 + ((y0 < L.h - 1) ? 1 : 0), x1 = x0 + ((x0 < L.w - 1) ? 1 : 0);
                const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
                const float hy = 1.0f - ly, hx = 1.0f - lx;
                const int wy = static_cast<int>(L.sy * static_cast<float>(t.y0)), wx = static_cast<int>(L.sx * static_cast<float>(t.x0));
                cw[l][0] = hy * hx; cw[l][1] = hy * lx; cw[l][2] = ly * hx; cw[l][3] = ly * lx;
                kidx[l][0] = k_off(l) + (y0 - wy) * win_w(l) + (x0 - wx);
                kidx[l][1] = k_off(l) + (y0 - wy) * win_w(l) + (x1 - wx);
                kidx[l][2] = k_off(l) + (y1 - wy) * win_w(l) + (x0 - wx);
                kidx[l][3] = k_off(l) + (y1 - wy) * win_w(l) + (x1 - wx);
                if (valid) {
                    const int64_t base = static_cast<int64_t>(t.n) * L.h * L.w;
                    const int pi[4] = {y0 * L.w + x0, y0 * L.w + x1, y1 * L.w + x0, y1 * L.w + x1};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 zc = __ldg(reinterpret_cast<const float4*>(L.z + (base + pi[c]) * p.zld));
                        lg.x = fmaf(cw[l][c], zc.x, lg.x);
                        lg.y = fmaf(cw[l][c], zc.y, lg.y);
                        lg.z = fmaf(cw[l][c], zc.z, lg.z);
                        lg.w = fmaf(cw[l][c], zc.w, lg.w);
                    }
                }
                // a clamped second corner coincides with the first: fold its weight in (it is 0 up to rounding anyway)
                if (x1 == x0) { cw[l][0] += cw[l][1]; cw[l][2] += cw[l][3]; cw[l][1] = 0.0f; cw[l][3] = 0.0f; }
                if (y1 == y0) { cw[l][0] += cw[l][2]; cw[l][1] += cw[l][3]; cw[l][2] = 0.0f; cw[l][3] = 0.0f; }
            }
            const float a0 = lg.x + b0, a1 = lg.y + b1, a2 = lg.z + b2, a3 = lg.w + b3;
            const float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
            const float e0 = expf(a0 - mx), e1 = expf(a1 - mx), e2 = expf(a2 - mx), e3 = expf(a3 - mx);
            const float inv = 1.0f / (e0 + e1 + e2 + e3);
            const float sl[4] = {valid ? e0 * inv : 0.0f, valid ? e1 * inv : 0.0f, valid ? e2 * inv : 0.0f, valid ? e3 * inv : 0.0f};
            // S is single-buffered: the previous tile's MMAs must have read it
            mbar_wait(sm.s_empty, (it & 1) ^ 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                sts128(srow + c * 16, make_uint4(0, 0, 0, 0));
                sts128(srow + 16384 + c * 16, make_uint4(0, 0, 0, 0));
            }
#pragma unroll
            for (int l = 0; l < 3; ++l) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float wgt = sl[l + 1] * cw[l][c];
                    if (wgt != 0.0f) {
                        const int k = kidx[l][c];
                        const uint32_t kb = static_cast<uint32_t>(k >> 6), kk = static_cast<uint32_t>(k & 63);
                        const __nv_bfloat16 hb = __float2bfloat16_rn(wgt);
                        sts_b16(srow + kb * 16384 + (((kk >> 3) ^ swz) << 4) + (kk & 7) * 2, *reinterpret_cast<const uint16_t*>(&hb));
                    }
                }
            }
            sm.s0[(it & 3) * 128 + r] = sl[0];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.s_full);
        }
    } else if (warp >= 8) {
        // ===================================================================== epilogue: two groups of 4 warps
        const int eg = (warp - 8) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const int B = kEpiBufs;
        int g = 0, slot = eg;
        uint32_t par = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            mbar_wait(&sm.tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const float s0 = sm.s0[(it & 3) * 128 + row];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 256);
            for (int c = 0; c < 4; ++c, ++g) {
                if ((g & 1) != eg) continue;
                const uint32_t mine = smem_u32(sm.epi) + static_cast<uint32_t>(slot) * kChunkBytes + static_cast<uint32_t>(row) * 128;
                mbar_wait(&sm.res_full[slot], par);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 64 + h * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t so = (static_cast<uint32_t>(h * 4 + j) ^ swz) << 4;
                        const uint4 u = lds128(mine + so);  // p2 of this very pixel, in place
                        uint4 o;
                        o.x = pack_bf16(fmaf(s0, bf16lo(u.x), __uint_as_float(v[j * 8 + 0])), fmaf(s0, bf16hi(u.x), __uint_as_float(v[j * 8 + 1])));
                        o.y = pack_bf16(fmaf(s0, bf16lo(u.y), __uint_as_float(v[j * 8 + 2])), fmaf(s0, bf16hi(u.y), __uint_as_float(v[j * 8 + 3])));
                        o.z = pack_bf16(fmaf(s0, bf16lo(u.z), __uint_as_float(v[j * 8 + 4])), fmaf(s0, bf16hi(u.z), __uint_as_float(v[j * 8 + 5])));
                        o.w = pack_bf16(fmaf(s0, bf16lo(u.w), __uint_as_float(v[j * 8 + 6])), fmaf(s0, bf16hi(u.w), __uint_as_float(v[j * 8 + 7])));
                        sts128(mine + so, o);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.out_ready[slot]);
                slot += 2;
                if (slot >= B) { slot -= B; par ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200

using namespace b200;

struct b200_asf_plan {
    AsfMmaTmaps tm;
    AsfMmaParams p;
    int grid;
};

extern "C" int b200_asf_plan_supported(int h2, int w2, int h3, int w3, int h4, int w4, int h5, int w5) {
    // the window bounds assume each level is (about) half the previous one; every level must hold its window box
    return h3 * 2 <= h2 + 1 && w3 * 2 <= w2 + 1 && h4 * 2 <= h3 + 1 && w4 * 2 <= w3 + 1 && h5 * 2 <= h4 + 1 && w5 * 2 <= w4 + 1 &&
           h3 >= win_h(0) && w3 >= win_w(0) && h4 >= win_h(1) && w4 >= win_w(1) && h5 >= win_h(2) && w5 >= win_w(2) && h2 >= kTileH &&
           w2 >= kTileW;
}

extern "C" int b200_asf_plan_create(const void* p2, const void* p3, const void* p4, const void* p5, const float* z2, const float* z3,
                                    const float* z4, const float* z5, int zld, const float* b_att, void* out, int batch, int h2,
                                    int w2, int h3, int w3, int h4, int w4, int h5, int w5, b200_asf_plan** plan) {
    if (!plan) return set_error("b200_asf_plan_create: null plan pointer");
    *plan = nullptr;
    if (!p2 || !p3 || !p4 || !p5 || !z2 || !z3 || !z4 || !z5 || !b_att || !out) return set_error("b200_asf_plan_create: null pointer");
    if (batch <= 0 || zld < 4 || (zld % 4)) return set_error("b200_asf_plan_create: bad shape");
    if (!b200_asf_plan_supported(h2, w2, h3, w3, h4, w4, h5, w5))
        return set_error("b200_asf_plan_create: level sizes not supported by the tensor-core kernel (use b200_asf_combine)");
    b200_asf_plan* pl = new (std::nothrow) b200_asf_plan;
    if (!pl) return set_error("b200_asf_plan_create: out of host memory");
    memset(pl, 0, sizeof(*pl));
    AsfMmaParams& P = pl->p;
    const void* ps[4] = {p2, p3, p4, p5};
    const float* zs[4] = {z2, z3, z4, z5};
    const int hs[4] = {h2, h3, h4, h5}, ws[4] = {w2, w3, w4, w5};
    for (int l = 0; l < 4; ++l) {
        P.lv[l].z = zs[l];
        P.lv[l].h = hs[l];
        P.lv[l].w = ws[l];
        P.lv[l].sy = (h2 > 1) ? static_cast<float>(hs[l] - 1) / static_cast<float>(h2 - 1) : 0.0f;
        P.lv[l].sx = (w2 > 1) ? static_cast<float>(ws[l] - 1) / static_cast<float>(w2 - 1) : 0.0f;
    }
    // window bound: rows touched by 8 output rows <= ceil(7 sy) + 2, columns by 16 output columns <= ceil(15 sx) + 2
    for (int l = 1; l < 4; ++l) {
        const int rows = static_cast<int>(7.0f * P.lv[l].sy + 0.9999f) + 2, cols = static_cast<int>(15.0f * P.lv[l].sx + 0.9999f) + 2;
        if (rows > win_h(l - 1) || cols > win_w(l - 1)) {
            delete pl;
            return set_error("b200_asf_plan_create: level %d needs a %d x %d window (kernel has %d x %d)", l + 2, cols, rows, win_w(l - 1),
                             win_h(l - 1));
        }
    }
    P.b_att = b_att;
    P.batch = batch;
    P.zld = zld;
    P.h2 = h2;
    P.w2 = w2;
    P.tiles_x = (w2 + kTileW - 1) / kTileW;
    P.tiles_y = (h2 + kTileH - 1) / kTileH;
    P.num_tiles = batch * P.tiles_x * P.tiles_y;
    int err = 0;
    for (int l = 1; l < 4 && !err; ++l) {
        const uint64_t dims[4] = {256, (uint64_t)ws[l], (uint64_t)hs[l], (uint64_t)batch};
        const uint64_t str[3] = {512, (uint64_t)ws[l] * 512, (uint64_t)hs[l] * ws[l] * 512};
        const uint32_t box[4] = {64u, (uint32_t)win_w(l - 1), (uint32_t)win_h(l - 1), 1u};
        err |= encode_map_typed(&pl->tm.win[l - 1], 0, 1, ps[l], 4, dims, str, box, "ASF level window");
    }
    if (!err) {
        const uint64_t dims[4] = {256, (uint64_t)w2, (uint64_t)h2, (uint64_t)batch};
        const uint64_t str[3] = {512, (uint64_t)w2 * 512, (uint64_t)h2 * w2 * 512};
        const uint32_t box[4] = {64u, (uint32_t)kTileW, (uint32_t)kTileH, 1u};
        err |= encode_map_typed(&pl->tm.p2, 0, 1, p2, 4, dims, str, box, "ASF p2");
        if (!err) err |= encode_map_typed(&pl->tm.out, 0, 1, out, 4, dims, str, box, "ASF output");
    }
    if (err) {
        delete pl;
        return 1;
    }
    cudaError_t e = cudaFuncSetAttribute(asf_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    int dev = 0, sms = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        delete pl;
        return set_error("b200_asf_plan_create: %s", cudaGetErrorString(e));
    }
    pl->grid = P.num_tiles < sms ? P.num_tiles : sms;
    *plan = pl;
    return 0;
}

extern "C" int b200_asf_plan_launch(const b200_asf_plan* pl, void* stream) {
    if (!pl) return set_error("b200_asf_plan_launch: null plan");
    asf_mma_kernel<<<pl->grid, kThreads, kSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(pl->tm, pl->p);
    return check_cuda(cudaGetLastError(), "b200_asf_plan launch");
}

extern "C" int b200_asf_plan_destroy(b200_asf_plan* pl) {
    delete pl;
    return 0;
}
