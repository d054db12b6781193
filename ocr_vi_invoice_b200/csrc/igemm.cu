// Implicit-GEMM convolution family for sm_100a: tcgen05.mma with TMEM accumulators, TMA-fed operands,
// warp-specialised persistent CTAs.  One kernel template covers
//   * plain 1x1 / 3x3 convolutions, stride 1 or 2 (A operand = NHWC activations loaded by TMA boxes; zero padding comes
//     from TMA out-of-bounds fill; stride 2 reads four (row, column)-parity strided views),
//   * "share mode" for narrow (N <= 128) 3x3 stride-1 convolutions and the 7x7 s2 stem: one box of tile_h+2 (stem:
//     tile_h+3) rows feeds the three (four) taps of a kernel column at 1024-byte-aligned row offsets, so the MMA warp pays
//     one barrier round trip per 12-16 MMAs and L2->SM traffic drops 2.4x,
//   * the modulated deformable conv v2 (A operand gathered by 16 producer warps from pre-computed 16-byte
//     b200_dcn_coord records: 2x2-block bilinear blend with mask-modulated bf16 weights in packed bf16x2 FMAs,
//     written straight into the 128B-swizzled UMMA layout; loads software-pipelined one unit ahead, K ordered
//     (channel chunk, tap) so the gathers hit in L1; 512-channel layers run one 128 x 512 tile per pixel block),
// all with a fused epilogue (folded BN scale/shift or bias, residual add, ReLU) that either
//   * stages bf16 64-channel chunks in swizzled shared memory and writes them with TMA stores
//     (pixel-shuffle for ConvTranspose2d 2x2 = four parity views of the output), the residual being
//     prefetched by TMA loads into the very buffer the result overwrites, or
//   * stores directly from registers (fp32 outputs, DCN coordinate records), or
//   * folds the DB head's last transposed conv in (B200_EPI_HEAD_DOT) and writes fp32 logits.
//
// GEMM view:  D[M = output pixels of one 128-pixel tile][N = out channels] =
//             A[M][K = taps * Cin] * B[N][K]^T,   K ordered (64-channel chunk, tap),  BLOCK_K = 64 (one SW128 atom).
//
// Things measured on B200 that shaped this file (tools/ring_probe.cu, mma_probe.cu, sts_probe.cu):
// an mbarrier hand-shake costs ~380 cycles of the waiting thread; tcgen05.mma M=128 K=16 takes 57/63/75/128 cycles at
// N = 32/64/128/256 (operand reads at ~100 B/clk, not math, below N = 256); uniform-datapath instructions inside an
// `if (lane == 0)` region get an ELECT/BRA.U.ANY loop each (so warps 0-3 run converged and elect); a few dozen bytes of
// spills or generic-pointer loads in the epilogue chunk loop cost the HBM-bound layers 10-15 % (so rare epilogues are
// their own instantiations and shared-memory parameters are read with explicit ld.shared).
//
// Reference call sites replaced: see include/dbnet_b200.h.
#include "common.cuh"
#include "internal.h"
#include "dcn2.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace b200 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB per stage (= one 128 x 64 bf16 chunk)
constexpr int kGatherWarps = 16;                // producer warps of the gather (DCN / stem) variant
constexpr int kMaxStages = 8;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kEpiThreads = 128;                // one epilogue group = 4 warps = the 4 TMEM lane quarters
constexpr int kMaxEpiBufs = 12;
// The TMA-fed conv kernel runs four epilogue groups (64-column chunks round-robin): its 1x1 layers are HBM-bound and
// their epilogue math is latency-bound with one warp per scheduler.  The gather kernel keeps one group (its threads
// go to the producers).  Warp 2 is the epilogue DMA warp: it alone issues the residual loads and the output stores.
__host__ __device__ constexpr int epi_groups(bool gather) { return gather ? 1 : 4; }
constexpr int kEpiWarp0 = 4;  // warps: 0 = A producer (TMA), 1 = MMA issuer, 2 = epilogue DMA, 3 = B producer (TMA)
__host__ __device__ constexpr int base_warps(bool gather) { return kEpiWarp0 + 4 * epi_groups(gather); }

struct IgemmParams {
    int mode, batch, in_h, in_w, in_c, out_h, out_w, out_c;
    int ksize, stride, flags, block_n;
    int stride_w;  // column stride (== stride except for the (2, 1) PatchMerging conv)
    int tile_w, tile_h, tiles_x, tiles_y;
    int num_n_blocks, num_tiles, k_blocks, cchunks;
    int out_ld, res_ld, shuf_c;
    int a_stages;  // depth of the A ring
    int pk_h, pk_w;        // STEM: rows / columns of the packed space-to-depth buffer
    int x_ld;              // elements between consecutive input pixels
    int epi_tma, res_tma;  // epilogue stores through TMA / residual prefetched through TMA
    int epi_active;        // epilogue groups that take chunks (the others only keep the TMEM hand-shake going)
    int epi_bufs;          // 16 KB epilogue buffers (residual lands in them, the result is written in place, TMA stores it)
    int ss_shared;         // (always 1) one copy of the layer's whole scale / shift vectors per CTA
    int res_bytes;         // bytes of one residual box (16 KB, or 4 KB for the half-resolution RES_UP2 box)
    int b_resident;        // the whole weight matrix stays in shared memory for the CTA's lifetime (small N x K)
    int kb_split;          // second K segment (x2): k-blocks >= kb_split come from tensor map a[1]; 0 = none
    int mma_group;         // share mode: the MMA warp issues the three kernel-column stages of a chunk as one group (small N)
    int dbg;               // development probes (B200_DEBUG): 1 = no MMA issue, 2 = no A loads, 4 = no epilogue math
    const __nv_bfloat16* x;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    void* y;
    const void* offsets;
    int b_stages;  // depth of the B ring (share mode)
    int a_kb;      // k-blocks one A stage serves: 3 = a (tile_h+2)-row box of one (chunk, kx) feeds the taps ky = 0..2
    int a_bytes;   // bytes of one A stage
    int a_shift;   // share mode: bytes between the A tiles of consecutive ky inside the box (tile_w * 128)
    int b_kb;      // k-blocks of weights per B stage
    int ss_half;   // floats between the scale and the shift vector in shared memory (n_pad, at least 256)
    int n_pad;     // GEMM columns of the layer (all n-blocks)
    int nacc;      // TMEM accumulators: 2 (tile i+1's MMAs overlap tile i's epilogue) or 1 (block_n = 512 fills TMEM)
    int b_rows;    // weight rows one CTA stages per k-block: block_n, or block_n / 2 in pair mode (cta_group::2)
    int num_units; // pair mode: units of work = (pair of pixel tiles, n-block); otherwise == num_tiles
    // B200_EPI_CHAIN (kEpi == 4): a second 1x1 GEMM over the staged output chunks
    int chain_n;        // its N (64)
    int chain_bytes;    // resident second weight matrix: chunks x chain_n x 128 bytes, after the first one in shared memory
    int chain_relu;
    const float* chain_scale;
    const float* chain_shift;
};

struct TmapPack {
    CUtensorMap a[4];
    CUtensorMap b;
    CUtensorMap out[4];  // (B200_EPI_CHAIN: out[1] = the second output)
    CUtensorMap res;
    CUtensorMap b2;      // B200_EPI_CHAIN: second weight matrix, box {64, chain_n}
};

struct SmemLayout {
    uint8_t* a;
    uint8_t* b;
    uint8_t* epi;  // epi_bufs x 16 KB
    float* ss;     // scale[256] then shift[256] of the current n-block (one copy, or one per epilogue group)
    uint64_t* a_full;
    uint64_t* a_empty;
    uint64_t* b_full;
    uint64_t* b_empty;
    uint64_t* tfull;
    uint64_t* tempty;
    uint64_t* res_full;   // [kMaxEpiBufs] residual box landed in the buffer
    uint64_t* slot_free;  // [kMaxEpiBufs] the store that used the buffer has read it (layers without residual)
    uint64_t* out_ready;  // [kMaxEpiBufs] the epilogue group has written its result
    uint64_t* b_ready;    // resident weights have landed
    uint32_t* tmem_slot;
};

template <bool kGather>
__device__ __forceinline__ SmemLayout carve(uint8_t* raw, const IgemmParams& p) {
    constexpr int G = epi_groups(kGather);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    SmemLayout s;
    s.a = base;
    s.b = s.a + p.a_stages * p.a_bytes;
    s.epi = s.b + (p.b_resident ? p.k_blocks : (p.a_kb > 1 ? p.b_stages * 3 : p.a_stages)) * p.b_rows * 128 + p.chain_bytes;
    uint8_t* rest = s.epi + p.epi_bufs * kABytes;
    s.ss = reinterpret_cast<float*>(rest);
    uint8_t* bars = rest + 8 * p.ss_half;
    s.a_full = reinterpret_cast<uint64_t*>(bars);
    s.a_empty = s.a_full + kMaxStages;
    s.b_full = s.a_empty + kMaxStages;
    s.b_empty = s.b_full + kMaxStages;
    s.tfull = s.b_empty + kMaxStages;
    s.tempty = s.tfull + 4;
    s.res_full = s.tempty + 4;
    s.slot_free = s.res_full + kMaxEpiBufs;
    s.out_ready = s.slot_free + kMaxEpiBufs;
    s.b_ready = s.out_ready + kMaxEpiBufs;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.b_ready + 1);
    return s;
}

// Development counters (B200_DEBUG & 16): cycles block 0 spends in each wait, read with b200_igemm_debug_counters().
__device__ unsigned long long g_prof[16];
#ifdef B200_PROF
#define B200_TIMED_WAIT(slot, stmt)                                                     \
    do {                                                                                \
        if (p.dbg & 16) {                                                               \
            const long long t_ = clock64();                                             \
            stmt;                                                                       \
            if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[slot], (unsigned long long)(clock64() - t_)); /* RED: no stall */ \
        } else {                                                                        \
            stmt;                                                                       \
        }                                                                               \
    } while (0)
#else
#define B200_TIMED_WAIT(slot, stmt) do { stmt; } while (0)
#endif

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
// 64-column chunks the TMA-store epilogue writes for n-block nb
__device__ __forceinline__ int chunks_of(const IgemmParams& p, int nb) {
    return min(p.block_n, p.out_c - nb * p.block_n) >> 6;
}

// ------------------------------------------------------------------------------------------ DCN coordinates
// torchvision deform_conv2d sampling of tap k at output pixel (oy, ox) -> one 16-byte b200_dcn_coord record:
// the element offset of the top-left pixel of a 2x2 block clamped inside the image, and the four bilinear weights
// (already multiplied by sigmoid(mask), zero for corners outside the image) as bf16, in block order
// (0,0) (0,1) (1,0) (1,1).  Clamping the block keeps all four addresses in bounds, so the producers need no
// per-corner predicates: a corner torchvision would skip simply carries weight 0.
__device__ __forceinline__ uint4 dcn_make_coord(float dy, float dx, float mlogit, int oy, int ox, int tap, int stride,
                                                int in_h, int in_w, int x_ld) {
    const int ky = tap / 3, kx = tap - ky * 3;
    const float m = 1.0f / (1.0f + expf(-mlogit));
    const float h = static_cast<float>(oy * stride - 1 + ky) + dy;
    const float w = static_cast<float>(ox * stride - 1 + kx) + dx;
    float s00 = 0.0f, s01 = 0.0f, s10 = 0.0f, s11 = 0.0f;
    int off = 0;
    if (h > -1.0f && w > -1.0f && h < static_cast<float>(in_h) && w < static_cast<float>(in_w)) {
        const float hf = floorf(h), wf = floorf(w);
        const int hl = static_cast<int>(hf), wl = static_cast<int>(wf);
        const float lh = h - hf, lw = w - wf;
        const float uh = 1.0f - lh, uw = 1.0f - lw;
        const int br = min(max(hl, 0), max(in_h - 2, 0)), bc = min(max(wl, 0), max(in_w - 2, 0));
        // weight of actual row r in {hl, hl+1} lands in block row r - br (0 or 1) if r is inside the image
        const float wt = (hl >= 0) ? uh : 0.0f, wb = (hl + 1 <= in_h - 1) ? lh : 0.0f;
        const float wlft = (wl >= 0) ? uw : 0.0f, wrgt = (wl + 1 <= in_w - 1) ? lw : 0.0f;
        const float r0 = (hl == br) ? wt : ((hl + 1 == br) ? wb : 0.0f);
        const float r1 = (hl == br + 1) ? wt : ((hl + 1 == br + 1) ? wb : 0.0f);
        const float c0 = (wl == bc) ? wlft : ((wl + 1 == bc) ? wrgt : 0.0f);
        const float c1 = (wl == bc + 1) ? wlft : ((wl + 1 == bc + 1) ? wrgt : 0.0f);
        s00 = m * (r0 * c0); s01 = m * (r0 * c1); s10 = m * (r1 * c0); s11 = m * (r1 * c1);
        off = (br * in_w + bc) * x_ld;
    }
    return make_uint4(static_cast<uint32_t>(off), pack_bf16(s00, s01), pack_bf16(s10, s11), 0u);
}

__device__ __forceinline__ void dcn_store_coords(uint4* dst, const float (&f)[32], int oy, int ox, int stride, int in_h,
                                                 int in_w, int x_ld) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
        dst[tap] = dcn_make_coord(f[2 * tap], f[2 * tap + 1], f[18 + tap], oy, ox, tap, stride, in_h, in_w, x_ld);
}

__global__ void dcn_coords_kernel(const float* __restrict__ raw, uint4* __restrict__ coords, int batch, int out_h,
                                  int out_w, int in_h, int in_w, int x_ld, int stride) {
    const int64_t total = static_cast<int64_t>(batch) * out_h * out_w;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % out_w);
        const int oy = static_cast<int>((i / out_w) % out_h);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(raw + i * 32 + j));
            f[j] = v.x; f[j + 1] = v.y; f[j + 2] = v.z; f[j + 3] = v.w;
        }
        dcn_store_coords(coords + i * 9, f, oy, ox, stride, in_h, in_w, x_ld);
    }
}

// ------------------------------------------------------------------------------------------ direct epilogue
// v[32] = fp32 accumulators of one pixel (TMEM lane) for 32 consecutive GEMM columns starting at col0.
__device__ __forceinline__ void epilogue_direct32(const IgemmParams& p, const uint32_t (&v)[32], int col0, int n, int oy,
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
    if (p.flags & B200_EPI_DCN_COORDS) {
        const int64_t opix = (static_cast<int64_t>(n) * p.out_h + oy) * p.out_w + ox;
        dcn_store_coords(reinterpret_cast<uint4*>(p.y) + opix * 9, f, oy, ox, p.stride, p.in_h, p.in_w, p.x_ld);
        return;
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

// Per-launch constants an epilogue reads with WARP-UNIFORM or compile-time indices travel as a kernel parameter, i.e. in
// the constant bank: an ld.shared.v4 of an address all 32 lanes share still costs four L1-data-pipe wavefronts (one per
// quarter warp), and the head-tail epilogue issued ~100 of them per 64 channels (ncu: LSU wavefronts 78 % of elapsed,
// tensor pipe 11 %).  Constant-bank operands cost none.  Only the head-tail instantiation carries a payload.
struct HeadConsts {
    float w2[256];     // last ConvTranspose2d(64 -> 1): [channel][dy2 * 2 + dx2]
    float scale[256];  // folded BN of the first transposed conv, [sub-pixel][channel]
    float shift[256];
    float bias;
};
struct NoConsts {};
template <int kEpi> struct EpiConsts { using type = NoConsts; };
template <> struct EpiConsts<2> { using type = HeadConsts; };

// ------------------------------------------------------------------------------------------ head tail epilogue
// B200_EPI_HEAD_DOT: the tile is ConvTranspose2d(64 -> 64, k2 s2) of one DB-head branch (N = 4 sub-pixels x 64 channels).
// The 64-channel bf16 map is never stored.  Epilogue group `sub` (one of the four sub-pixels) folds BN + ReLU, rounds to
// bf16 exactly as the stored map would be and writes the 128 x 64 tile into a swizzled staging buffer -- which is a valid
// K-major A operand: the LAST ConvTranspose2d(64 -> 1, k2 s2) (model/det/head.py:15,23) is then a SECOND UMMA
// (128 x 16 x 64, four useful columns = the 2 x 2 output sub-pixels, weights rounded to bf16) issued by warp 3 into its own
// TMEM columns, and the group reads the four fp32 logits back one tile later (software-pipelined, double-buffered staging
// and D2), adds the bias and stores the 4 x 4 patch of the full-resolution logit map.  History: 256 FMAs per thread on the
// CUDA cores with ld.shared operands 0.142 ms -> constant-bank operands 0.118 ms (issue-bound) -> this.
// First half: accumulators -> BN + ReLU -> bf16 staging rows.  scale / shift come from the constant bank (HeadConsts).
__device__ __forceinline__ void head_tail_stage(const IgemmParams& p, uint32_t taddr, uint32_t rowaddr, uint32_t swz, int sub,
                                                uint64_t* tempty, int lane) {
    // The accumulators already hold BN(conv): the BN scale is folded into the weights and the shift was added by one extra
    // K = 16 UMMA step (ones column x [shift_hi, shift_lo] rows, see the kernel prologue), so nothing per-column is
    // fetched here -- with 16 epilogue warps, 128 ld.shared.v4 (four L1 wavefronts each, even when all lanes read one
    // address) or constant-bank loads per thread and tile were what bounded this kernel (0.14 / 0.12 ms).
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(taddr + sub * 64 + half * 32, v);
        tmem_ld_wait();
        if (half == 1) {  // main accumulator fully read by this warp: the next tile's MMAs may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[j * 8 + i]);
            if (p.flags & B200_EPI_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.0f);
            }
            uint4 o;
            o.x = pack_bf16(f[0], f[1]);
            o.y = pack_bf16(f[2], f[3]);
            o.z = pack_bf16(f[4], f[5]);
            o.w = pack_bf16(f[6], f[7]);
            sts128(rowaddr + ((static_cast<uint32_t>(half * 4 + j) ^ swz) << 4), o);
        }
    }
}
// Second half: the four dot products of one pixel (TMEM columns dy2 * 2 + dx2) + bias -> 2 x 2 patch of the logit map.
__device__ __forceinline__ void head_tail_store(const IgemmParams& p, float bias, uint32_t d2addr, int n, int oy, int ox, bool valid,
                                                int sub) {
    uint32_t v[4];
    tmem_ld4(d2addr, v);
    tmem_ld_wait();
    if (valid) {
        float* out = reinterpret_cast<float*>(p.y);
        const int ow = 4 * p.out_w;
        const int64_t row = (static_cast<int64_t>(n) * (4 * p.out_h) + 4 * oy + 2 * (sub >> 1));
        float* o = out + row * ow + 4 * ox + 2 * (sub & 1);
        *reinterpret_cast<float2*>(o) = make_float2(__uint_as_float(v[0]) + bias, __uint_as_float(v[1]) + bias);
        *reinterpret_cast<float2*>(o + ow) = make_float2(__uint_as_float(v[2]) + bias, __uint_as_float(v[3]) + bias);
    }
}

// ------------------------------------------------------------------------------------------ GELU epilogue (kEpi == 3)
// nn.GELU (exact erf form) of the SVTRv2 recogniser (model/rec2/svtrv2.py:33,51,54,114,117) applied to scale * acc + shift,
// BEFORE the optional residual add (LocalMixing: x + act(bn(conv(.)))).  erf by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7, far below a bf16 ulp): one MUFU.RCP + one MUFU.EX2 + seven FMAs per element.
// Two elements at a time on the packed-fp32 pipe (FFMA2 / FMUL2, sm_100): GELU(x) = x * (x >= 0 ? 1 - h : h) with
// h = 0.5 * (a1 t + .. + a5 t^5) * exp(-x^2 / 2), t = 1 / (1 + p |x| / sqrt 2) -- the 0.5 and the 1/sqrt 2 are folded into the
// constants, so a pair costs ~20 issue slots instead of ~36 (the GELU epilogues are issue-bound: ncu 73 % issue active).
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_as2(float& x0, float& x1) {
    const float u0 = fabsf(x0), u1 = fabsf(x1);
    const uint64_t u = pk2(u0, u1);
    float d0, d1;
    upk2(fma2(u, pk2(0.23164189f, 0.23164189f), pk2(1.0f, 1.0f)), d0, d1);  // 0.3275911 / sqrt 2
    const uint64_t t = pk2(rcp_approx(d0), rcp_approx(d1));
    uint64_t q = fma2(pk2(0.5307027145f, 0.5307027145f), t, pk2(-0.7265760135f, -0.7265760135f));  // 0.5 * a5, 0.5 * a4
    q = fma2(q, t, pk2(0.7107068705f, 0.7107068705f));
    q = fma2(q, t, pk2(-0.142248368f, -0.142248368f));
    q = fma2(q, t, pk2(0.127414796f, 0.127414796f));
    q = mul2(q, t);
    float s0, s1;
    upk2(mul2(mul2(u, u), pk2(-0.72134752f, -0.72134752f)), s0, s1);  // -x^2 / 2 * log2 e
    float h0, h1;
    upk2(mul2(q, pk2(ex2_approx_f(s0), ex2_approx_f(s1))), h0, h1);
    x0 *= (x0 >= 0.0f) ? 1.0f - h0 : h0;
    x1 *= (x1 >= 0.0f) ? 1.0f - h1 : h1;
}

// ------------------------------------------------------------------------------------------ kernel
// kVariant: 0 = TMA-fed conv (one A tile + one weight k-block per stage), 1 = gather producers (DCN / stem fallback),
// 2 = TMA-fed 3x3 stride-1 conv in share mode.  Separate instantiations: the loops of one variant must not cost the
// others registers or instructions (the MMA warp and the epilogue warps are latency-critical).
// kEpi: 0 = common epilogue; 1 = adds a nearest-upsampled half-resolution residual (FPN top-down path, keeps eight more
// vector registers alive); 2 = DB-head tail folded in (B200_EPI_HEAD_DOT); 3 = GELU before the residual add
// (B200_EPI_GELU, the recogniser's Linear / conv + GELU layers); 4 = a second 1x1 GEMM chained onto the staged output chunks
// (B200_EPI_CHAIN: conv3 + residual + ReLU of one Bottleneck and conv1 + BN + ReLU of the next in one launch).  Own
// instantiations: the common epilogue must not pay registers (spills) for the rare ones.
// kPair: cta_group::2 -- the two CTAs of a cluster run one 256-pixel x block_n UMMA per K step (variants 0 and 2): each stages
// its own 128-pixel A tile and HALF of the weight rows, rank 0 issues the MMAs and multicasts the commits, each CTA's
// epilogue drains its own TMEM.  Per k-block an SM pulls A + B/2 instead of A + B through the L2 fabric, which is what the
// weight-heavy small-spatial layers (layer3 / layer4 1x1 convs, the top of the neck) are bound by (~42 B/clk per SM).
template <int kVariant, int kEpi = 0, bool kPair = false>
__global__ void __launch_bounds__((base_warps(kVariant == 1) + (kVariant == 1 ? kGatherWarps : 0)) * 32, 1)
igemm_kernel(const __grid_constant__ TmapPack tm, const IgemmParams p, const __grid_constant__ typename EpiConsts<kEpi>::type ec) {
    extern __shared__ uint8_t smem_raw[];
    constexpr bool kGather = (kVariant == 1);
    constexpr bool kShare = (kVariant == 2);
    constexpr bool kUp2 = (kEpi == 1);
    constexpr int kBaseWarps = base_warps(kGather);
    constexpr int G = epi_groups(kGather);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b_bytes = p.b_rows * 128;
    const SmemLayout sm = carve<kGather>(smem_raw, p);
    // unit of work u -> tile index of THIS CTA.  Pair mode: unit = (pixel-tile pair mp, n-block); rank r takes pixel tile
    // 2 mp + r (an odd tile count leaves the last rank-1 tile outside the tensor: TMA zero-fills its loads, clips its stores)
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const int u_first = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int u_stride = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int num_units = kPair ? p.num_units : p.num_tiles;
    auto unit_tile = [&](int u) -> int {
        if (!kPair) return u;
        const int mp = u / p.num_n_blocks;
        return (2 * mp + static_cast<int>(cta_rank)) * p.num_n_blocks + (u - mp * p.num_n_blocks);
    };
    const int stages = p.a_stages;  // A ring depth (the gather producers fill it too)
    // head tail (kEpi 2): one 256-column accumulator + 4 sub-pixels x 2 x 16 columns for the second UMMA's results
    // chain (kEpi 4): one 256-column accumulator + 2 x chain_n columns for the second GEMM
    const uint32_t tmem_cols = (kEpi == 2 || kEpi == 4) ? 512u : ((p.nacc * p.block_n < 32) ? 32u : static_cast<uint32_t>(p.nacc * p.block_n));

    // Programmatic dependent launch: the next launch of the stream may be scheduled onto SMs as this grid's blocks
    // retire; its prologue (below, up to pdl_wait) then overlaps our tail.  Nothing before pdl_wait() touches
    // activations -- only parameters (scale / shift / tail weights), which no kernel writes.
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm.b);
        if (!kGather) {
            tma_prefetch_desc(&tm.a[0]);
            if (p.kb_split) tma_prefetch_desc(&tm.a[1]);
            if (p.stride == 2 && p.ksize == 3) {
                tma_prefetch_desc(&tm.a[1]);
                if (p.stride_w == 2) {
                    tma_prefetch_desc(&tm.a[2]);
                    tma_prefetch_desc(&tm.a[3]);
                }
            }
        }
        if (p.epi_tma) tma_prefetch_desc(&tm.out[0]);
        if (p.res_tma) tma_prefetch_desc(&tm.res);
        if constexpr (kEpi == 4) {
            tma_prefetch_desc(&tm.out[1]);
            tma_prefetch_desc(&tm.b2);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(&sm.a_full[s], kGather ? 1 + kGatherWarps : 1);
                mbar_init(&sm.a_empty[s], 1);
            }
            if constexpr (kShare) {
                for (int s = 0; s < p.b_stages; ++s) {
                    mbar_init(&sm.b_full[s], 1);
                    mbar_init(&sm.b_empty[s], 1);
                }
            }
            for (int a = 0; a < 4; ++a) {  // (share mode at N <= 64 runs four accumulators)
                mbar_init(&sm.tfull[a], 1);
                mbar_init(&sm.tempty[a], (kPair ? 2 : 1) * 4 * G);  // pair: the epilogue warps of BOTH CTAs release rank 0's barrier
            }
            if constexpr (kEpi == 4) {
                for (int a = 0; a < 2; ++a) {
                    mbar_init(&sm.b_full[a], 1);   // d2_full: the chained GEMM's accumulator of tile parity a is complete
                    mbar_init(&sm.b_empty[a], 4 * (p.chain_n >> 6));  // d2_empty: four warps per 64-column tail chunk
                }
            }
            mbar_init(sm.b_ready, 1);
            for (int a = 0; a < kMaxEpiBufs; ++a) {
                mbar_init(&sm.res_full[a], 1);
                mbar_init(&sm.slot_free[a], 1);
                mbar_init(&sm.out_ready[a], 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if constexpr (kPair) {
            tmem_alloc_2sm(sm.tmem_slot, tmem_cols);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(sm.tmem_slot, tmem_cols);
            tmem_relinquish();
        }
    }
    if constexpr (kEpi == 2) {
        // B operand of the second UMMA: w2 as bf16 [16 rows (4 used: dy2 * 2 + dx2)][64 channels], K-major, 128B swizzle
        uint8_t* w2s = reinterpret_cast<uint8_t*>(sm.ss);
        for (int i = threadIdx.x; i < 2048 / 16; i += blockDim.x) reinterpret_cast<uint4*>(w2s)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (threadIdx.x < 256) {
            const int nn = threadIdx.x >> 6, k = threadIdx.x & 63;
            const __nv_bfloat16 hb = __float2bfloat16_rn(ec.w2[k * 4 + nn]);
            *reinterpret_cast<__nv_bfloat16*>(w2s + nn * 128 + (((k >> 3) ^ nn) << 4) + (k & 7) * 2) = hb;
        }
        // BN shift by UMMA: A = [128 x 16] with columns 0 and 1 = 1, B = [256 x 16] with column 0 = bf16(shift), column 1 =
        // bf16(shift - column 0) (two bf16 terms: fp32-grade shift).  Both K-major WITHOUT swizzle (tools/noswizzle_probe.cu):
        // element (r, k) at (r % 8) * 16 + (r / 8) * 128 + (k / 8) * LBO + (k % 8) * 2, LBO = 2048 (A) / 4096 (B).
        uint8_t* ones = w2s + 2048;
        uint8_t* shb = w2s + 8192;
        for (int i = threadIdx.x; i < (4096 + 8192) / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(i < 256 ? ones : shb - 4096)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (threadIdx.x < 128) {
            const int r = threadIdx.x;
            *reinterpret_cast<uint32_t*>(ones + (r & 7) * 16 + (r >> 3) * 128) = 0x3F803F80u;  // (1.0, 1.0) bf16
        }
        if (threadIdx.x < 256) {
            const int nn = threadIdx.x;
            const float sh = ec.shift[nn];
            const __nv_bfloat16 hi = __float2bfloat16_rn(sh);
            const __nv_bfloat16 lo = __float2bfloat16_rn(sh - __bfloat162float(hi));
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(shb + (nn & 7) * 16 + (nn >> 3) * 128);
            dst[0] = hi;
            dst[1] = lo;
        }
        fence_proxy_async_smem();
    }
    if (p.epi_tma) {  // the layer's whole scale / shift vectors, staged once (no per-tile reloads, no named barriers)
        for (int i = threadIdx.x; i < p.n_pad; i += blockDim.x) {
            sm.ss[i] = __ldg(p.scale + i);
            sm.ss[p.ss_half + i] = __ldg(p.shift + i);
        }
        if constexpr (kEpi == 4) {  // the chained GEMM's columns follow the layer's own: "columns n_pad .. n_pad + chain_n"
            for (int i = threadIdx.x; i < p.chain_n; i += blockDim.x) {
                sm.ss[p.n_pad + i] = __ldg(p.chain_scale + i);
                sm.ss[p.ss_half + p.n_pad + i] = __ldg(p.chain_shift + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    pdl_wait();  // the producing launches have completed: activations, residuals and sampling records may be read
#ifdef B200_PROF
    const long long prof_t0 = (p.dbg & 16) ? clock64() : 0;
#endif

    if (warp == 0) {
        // ===================================================================== TMA producer
        // (converged warp, one elected lane issues -- see the MMA warp)
        if (p.b_resident) {  // small weight matrix: fetch it once, every tile multiplies against the same copy
            if (elect_one()) {
                mbar_arrive_expect_tx(sm.b_ready, p.k_blocks * b_bytes + p.chain_bytes);
                for (int kb = 0; kb < p.k_blocks; ++kb)
                    tma_load_2d(sm.b + kb * b_bytes, &tm.b, sm.b_ready, kb * kBlockK, 0);
                if constexpr (kEpi == 4) {  // second weight matrix, one [chain_n x 64] k-block per output chunk
                    uint8_t* w2 = sm.epi - p.chain_bytes;
                    for (int c = 0; c < p.n_pad / 64; ++c)
                        tma_load_2d(w2 + c * p.chain_n * 128, &tm.b2, sm.b_ready, c * kBlockK, 0);
                }
            }
            __syncwarp();
        }
        int s = 0;
        uint32_t ph = 0;
        if constexpr (kShare) {
            // Share mode: one box of tile_h+2 rows per (chunk, kx); the MMA warp reads the taps ky = 0..2 out of it at
            // row offsets, so a 3x3 stride-1 conv pulls 3*(tile_h+2)/tile_h tiles per chunk through L2 instead of 9 and
            // pays one barrier round trip per 3 k-blocks.  The weights travel on their own ring (warp 3).
            const uint32_t tx_bytes = (p.dbg & 2) ? 0u : static_cast<uint32_t>(p.a_bytes);
            const bool stem = (p.mode == B200_IGEMM_STEM);  // one box per tile: rows y0 .. y0+tile_h+2 of the view
            for (int u = u_first; u < num_units; u += u_stride) {
                const int tile = unit_tile(u);
                const TileCoord t = decode_tile(p, tile);
                if (stem) {
                    B200_TIMED_WAIT(0, mbar_wait(&sm.a_empty[s], ph ^ 1));
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&sm.a_full[s], tx_bytes);
                        if (!(p.dbg & 2)) tma_load_4d(sm.a + s * p.a_bytes, &tm.a[0], &sm.a_full[s], 0, t.x0, t.y0, t.n);
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                    continue;
                }
                for (int c0 = 0; c0 < p.in_c; c0 += kBlockK) {
                    for (int kx = 0; kx < 3; ++kx) {
                        B200_TIMED_WAIT(0, mbar_wait(&sm.a_empty[s], ph ^ 1));
                        if constexpr (kPair) {  // both CTAs' boxes land on rank 0's barrier
                            if (elect_one()) {
                                if (cta_rank == 0) mbar_arrive_expect_tx(&sm.a_full[s], 2 * tx_bytes);
                                tma_load_4d_2sm(sm.a + s * p.a_bytes, &tm.a[0], mapa_u32(smem_u32(&sm.a_full[s]), 0), c0,
                                                t.x0 + kx - 1, t.y0 - 1, t.n);
                            }
                        } else if (elect_one()) {
                            mbar_arrive_expect_tx(&sm.a_full[s], tx_bytes);
                            if (!(p.dbg & 2))
                                tma_load_4d(sm.a + s * p.a_bytes, &tm.a[0], &sm.a_full[s], c0, t.x0 + kx - 1, t.y0 - 1, t.n);
                        }
                        __syncwarp();
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        } else {
            // One 128-pixel A tile and one weight k-block per stage, both landing on the stage's barrier.
            const int taps = p.k_blocks / p.cchunks;  // K is ordered (64-channel chunk, tap)
            // pair mode: rank 0's barrier collects the bytes of both CTAs' A tiles and weight halves
            const uint32_t tx_bytes = (kPair ? 2u : 1u) * (((kGather || (p.dbg & 2)) ? 0 : kABytes) + (p.b_resident ? 0 : b_bytes));
            for (int u = u_first; u < num_units; u += u_stride) {
                const int tile = unit_tile(u);
                const TileCoord t = decode_tile(p, tile);
                int tap = 0, c0 = 0;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    int ax = t.x0, ay = t.y0, view = 0;
                    if (!kGather) {
                        if (p.mode == B200_IGEMM_STEM) {
                            ay += tap;  // tap = block row of the 4x4 window; the view's rows overlap by 3 blocks
                        } else if (p.ksize == 3) {
                            const int ky = tap / 3, kx = tap - ky * 3;
                            if (p.stride == 1) {
                                ax += kx - 1;
                                ay += ky - 1;
                            } else if (p.stride_w == 1) {
                                // stride (2, 1): input pixel (2*oy + ky - 1, ox + kx - 1) on the row-parity view
                                view = (ky + 1) & 1;
                                ax += kx - 1;
                                ay += (ky == 0) ? -1 : 0;
                            } else {
                                // input pixel (2*oy + ky - 1, 2*ox + kx - 1) on the (row parity, column parity) view
                                view = (((ky + 1) & 1) << 1) | ((kx + 1) & 1);
                                ax += (kx == 0) ? -1 : 0;
                                ay += (ky == 0) ? -1 : 0;
                            }
                        }
                    }
                    int ca = c0;  // channel offset inside the source tensor of this k-block
                    if (!kGather && p.kb_split && kb >= p.kb_split) {  // second K segment: the other tensor (map a[1])
                        view = 1;
                        ca = c0 - p.kb_split * kBlockK;
                    }
                    B200_TIMED_WAIT(0, mbar_wait(&sm.a_empty[s], ph ^ 1));
                    if constexpr (kPair) {
                        if (elect_one()) {
                            if (cta_rank == 0) mbar_arrive_expect_tx(&sm.a_full[s], tx_bytes);
                            const uint32_t bar0 = mapa_u32(smem_u32(&sm.a_full[s]), 0);
                            tma_load_4d_2sm(sm.a + s * kABytes, &tm.a[view], bar0, ca, ax, ay, t.n);
                            tma_load_2d_2sm(sm.b + s * b_bytes, &tm.b, bar0, kb * kBlockK,
                                            t.nb * p.block_n + static_cast<int>(cta_rank) * p.b_rows);
                        }
                    } else if (elect_one()) {
                        mbar_arrive_expect_tx(&sm.a_full[s], tx_bytes);
                        if (!kGather && !(p.dbg & 2))
                            tma_load_4d(sm.a + s * kABytes, &tm.a[view], &sm.a_full[s], ca, ax, ay, t.n);
                        if (!p.b_resident) {
                            tma_load_2d(sm.b + s * b_bytes, &tm.b, &sm.a_full[s], kb * kBlockK, t.nb * p.block_n);
                            if (kGather && p.block_n > 256)  // a weight box holds at most 256 rows
                                tma_load_2d(sm.b + s * b_bytes + 256 * 128, &tm.b, &sm.a_full[s], kb * kBlockK,
                                            t.nb * p.block_n + 256);
                        }
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                    if (++tap == taps) { tap = 0; c0 += kBlockK; }
                }
            }
        }
    } else if (warp == 1 && (!kPair || cta_rank == 0)) {
        // ===================================================================== MMA issuer (pair mode: rank 0 only)
        // The whole warp runs the loop converged and one elected lane issues: every operand of UTCHMMA / UTCBAR then
        // lives in uniform registers.  (Inside an `if (lane == 0)` region ptxas wraps each uniform-datapath
        // instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, ~80 cycles per MMA: small-N layers were bound by
        // this thread, not by the tensor pipe.)
        // UTCHMMA back-pressures this thread (the tensor-core queue holds only a few instructions), so whatever the loop
        // spends between MMAs is tensor-pipe idle time whenever a k-block is short: each mode gets its own lean loop.
        const uint32_t idesc = umma_idesc_bf16(static_cast<uint32_t>(p.block_n));
        if (p.b_resident) mbar_wait(sm.b_ready, 0);
        const uint32_t a0 = smem_u32(sm.a), b0 = smem_u32(sm.b);
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        if constexpr (kShare) {
            // share mode: one A stage (and one weight stage, unless the weights are resident) = 3 k-blocks = 12 MMAs
            const bool b_ring = !p.b_resident;
            const int bst = p.b_stages, akb = p.a_kb, a_per_tile = p.k_blocks / akb;
            const uint32_t sh = static_cast<uint32_t>(p.a_shift) >> 4, be = static_cast<uint32_t>(b_bytes) >> 4;
            // weight k-blocks between the sub-tiles of a stage: ring stage = 3 consecutive k-blocks; resident 3x3 weights are
            // (chunk, ky, kx) so ky steps by 3; the stem's four window rows are consecutive
            const uint32_t wstep = (b_ring || akb == 4) ? be : 3 * be;
            const uint32_t idesc2 = umma_idesc_bf16_2sm(static_cast<uint32_t>(p.block_n)); (void)idesc2;
            int bs = 0;
            uint32_t bph = 0;
            const int amask = p.nacc - 1, ashift = (p.nacc == 4) ? 2 : 1;  // share mode: 2 or 4 accumulators
            if (p.mma_group == 2) {
                // One stage per tile (stem: 16 instructions; 3x3 64 -> 64 would be three): TWO tiles per issue group, each into
                // its own accumulator (four in flight), for the same reason the kernel columns of a chunk are grouped below --
                // the stem's MMA warp spent ~1000 cycles per tile outside instruction issue against ~1040 inside.
                for (int u = u_first; u < num_units; u += 2 * u_stride, it += 2) {
                    const bool two = (u + u_stride < num_units);
                    int sj[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        sj[j] = s;
                        if (j == 0 || two) {
                            B200_TIMED_WAIT(2, mbar_wait(&sm.tempty[(it + j) & amask], (((it + j) >> ashift) & 1) ^ 1));
                            B200_TIMED_WAIT(3, mbar_wait(&sm.a_full[s], ph));
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
                    tc_fence_after();
#ifdef B200_PROF
                    const long long ti0 = clock64();
#endif
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 0 || two) {
                                const int acc = (it + j) & amask;
                                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.block_n);
                                const uint64_t da = umma_desc_sw128(a0 + sj[j] * p.a_bytes);
                                const uint64_t db = umma_desc_sw128(b0);
#pragma unroll
                                for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
                                    for (int k = 0; k < kBlockK / 16; ++k)
                                        umma_bf16(d_tmem, da + ky * sh + 2 * k, db + ky * wstep + 2 * k, idesc, (ky | k) ? 1u : 0u);
                                }
                                umma_commit(&sm.a_empty[sj[j]]);
                                umma_commit(&sm.tfull[acc]);
                            }
                        }
                    }
                    __syncwarp();
#ifdef B200_PROF
                    if ((p.dbg & 16) && blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[10], (unsigned long long)(clock64() - ti0));
#endif
                }
            } else
            for (int u = u_first; u < num_units; u += u_stride, ++it) {
                const int tile = unit_tile(u); (void)tile;
                const int acc = it & amask;
                B200_TIMED_WAIT(2, mbar_wait(&sm.tempty[acc], ((it >> ashift) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.block_n);
                int wbase = 0, kx = 0;  // resident weights: k-block of (chunk, ky = 0, kx)
                if (p.mma_group) {
                    // 3x3: the three kernel columns of a 64-channel chunk = three consecutive stages are issued as ONE group of
                    // 36 instructions.  What precedes the first UTCHMMA of a group (waits, election, stage bases through
                    // R2UR: ~55 instructions of one warp) leaves the tensor pipe idle once its short queue has drained --
                    // ~480 cycles per 12 instructions of ~57 cycles at N = 32 when every stage was a group of its own; the
                    // descriptor arithmetic of the other 35 instructions sits between instructions the pipe is executing.
                    // N <= 64 only (host: mma_group): the group waits for three stages at once, which needs a deep ring, and
                    // at N >= 128 the queue already covers the gap (grouping cost the N = 256 pair layers 17 %).
                    for (int c = 0; c < a_per_tile; c += 3) {
                        int sj[3], bj[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            sj[j] = s;
                            bj[j] = bs;
                            B200_TIMED_WAIT(3, mbar_wait(&sm.a_full[s], ph));
                            if (b_ring) B200_TIMED_WAIT(4, mbar_wait(&sm.b_full[bs], bph));
                            if (++s == stages) { s = 0; ph ^= 1; }
                            if (b_ring && ++bs == bst) { bs = 0; bph ^= 1; }
                        }
                        tc_fence_after();
#ifdef B200_PROF
                        const long long ti0 = clock64();
#endif
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const uint64_t da = umma_desc_sw128(a0 + sj[j] * p.a_bytes);
                                const uint64_t db = umma_desc_sw128(b0 + (b_ring ? bj[j] * 3 : wbase + j) * b_bytes);
#pragma unroll
                                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                                    for (int k = 0; k < kBlockK / 16; ++k) {
                                        if constexpr (kPair)
                                            umma_bf16_2sm(d_tmem, da + ky * sh + 2 * k, db + ky * wstep + 2 * k, idesc2, (c | j | ky | k) ? 1u : 0u);
                                        else
                                            umma_bf16(d_tmem, da + ky * sh + 2 * k, db + ky * wstep + 2 * k, idesc, (c | j | ky | k) ? 1u : 0u);
                                    }
                                }
                                if constexpr (kPair) {
                                    umma_commit_2sm(&sm.a_empty[sj[j]]);
                                    umma_commit_2sm(&sm.b_empty[bj[j]]);
                                } else {
                                    umma_commit(&sm.a_empty[sj[j]]);
                                    if (b_ring) umma_commit(&sm.b_empty[bj[j]]);
                                }
                            }
                        }
                        __syncwarp();
#ifdef B200_PROF
                        if ((p.dbg & 16) && blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[10], (unsigned long long)(clock64() - ti0));
#endif
                        wbase += 9;
                    }
                } else
                for (int ai = 0; ai < a_per_tile; ++ai) {
                    B200_TIMED_WAIT(3, mbar_wait(&sm.a_full[s], ph));
                    if (b_ring) B200_TIMED_WAIT(4, mbar_wait(&sm.b_full[bs], bph));
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a0 + s * p.a_bytes);
                    const uint64_t db = umma_desc_sw128(b0 + (b_ring ? bs * 3 : wbase) * b_bytes);
#ifdef B200_PROF
                    const long long ti0 = clock64();
#endif
                    if constexpr (kPair) {
                        if (elect_one()) {  // 256 pixels x block_n per instruction: both CTAs' boxes and weight halves
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    umma_bf16_2sm(d_tmem, da + ky * sh + 2 * k, db + ky * wstep + 2 * k, idesc2, (ai | ky | k) ? 1u : 0u);
                            }
                            umma_commit_2sm(&sm.a_empty[s]);
                            umma_commit_2sm(&sm.b_empty[bs]);
                        }
                    } else if (elect_one()) {
                        if (!(p.dbg & 1)) {
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    umma_bf16(d_tmem, da + ky * sh + 2 * k, db + ky * wstep + 2 * k, idesc, (ai | ky | k) ? 1u : 0u);
                            }
                            if (akb == 4) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    umma_bf16(d_tmem, da + 3 * sh + 2 * k, db + 3 * wstep + 2 * k, idesc, 1u);
                            }
                        }
                        umma_commit(&sm.a_empty[s]);
                        if (b_ring) umma_commit(&sm.b_empty[bs]);
                    }
                    __syncwarp();
#ifdef B200_PROF
                    if ((p.dbg & 16) && blockIdx.x == 0 && lane == 0) atomicAdd(&g_prof[10], (unsigned long long)(clock64() - ti0));
#endif
                    if (++s == stages) { s = 0; ph ^= 1; }
                    if (b_ring && ++bs == bst) { bs = 0; bph ^= 1; }
                    if (++kx == 3) { kx = 0; wbase += 7; } else { ++wbase; }
                }
                if (elect_one()) {
                    if constexpr (kPair) umma_commit_2sm(&sm.tfull[acc]);
                    else umma_commit(&sm.tfull[acc]);
                }
                __syncwarp();
            }
        } else {
            // N = 512 (gather variant only): two N = 256 instructions per K step, one accumulator
            const bool wide = kGather && p.block_n > 256;
            const int nacc = kGather ? p.nacc : ((kEpi == 2 || kEpi == 4) ? 1 : 2);
            const uint32_t idesc_l = kPair ? umma_idesc_bf16_2sm(static_cast<uint32_t>(p.block_n)) : (wide ? umma_idesc_bf16(256u) : idesc);
            for (int u = u_first; u < num_units; u += u_stride, ++it) {
                const int tile = unit_tile(u); (void)tile;
                const int acc = (nacc == 2) ? (it & 1) : 0;
                const uint32_t tph = (nacc == 2) ? ((it >> 1) & 1) : (it & 1);
                B200_TIMED_WAIT(2, mbar_wait(&sm.tempty[acc], tph ^ 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.block_n);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    B200_TIMED_WAIT(3, mbar_wait(&sm.a_full[s], ph));
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a0 + s * kABytes);
                    const uint64_t db = umma_desc_sw128(b0 + (p.b_resident ? kb : s) * b_bytes);
                    if constexpr (kPair) {
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)  // one 256 x block_n UMMA: A rows and weight halves of both CTAs
                                umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc_l, (kb | k) ? 1u : 0u);
                            umma_commit_2sm(&sm.a_empty[s]);  // frees the stage in both CTAs
                        }
                    } else if (elect_one()) {
                        if (!(p.dbg & 1)) {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) {
                                // +32 bytes per UMMA_K step inside the 128B swizzle atom (encoded >> 4)
                                umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc_l, (kb | k) ? 1u : 0u);
                                if (wide) umma_bf16(d_tmem + 256, da + 2 * k, db + (256 * 128 >> 4) + 2 * k, idesc_l, (kb | k) ? 1u : 0u);
                            }
                        }
                        umma_commit(&sm.a_empty[s]);
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                if (elect_one()) {
                    if constexpr (kEpi == 2) {  // + BN shift: ones column x shift rows (no-swizzle K-major operands, one K = 16 step)
                        const uint32_t cst = smem_u32(sm.ss);
                        auto noswz = [](uint32_t a, uint32_t lbo) {
                            return static_cast<uint64_t>((a & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
                                   (static_cast<uint64_t>(128 >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
                        };
                        umma_bf16(d_tmem, noswz(cst + 2048, 2048), noswz(cst + 8192, 4096), idesc_l, 1u);
                    }
                    if constexpr (kPair) umma_commit_2sm(&sm.tfull[acc]);
                    else umma_commit(&sm.tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ===================================================================== epilogue DMA warp
        // One elected lane owns every epilogue TMA transfer (the warp runs converged).  Chunk g (CTA-wide order) uses
        // buffer g % B.  After the store of chunk g is queued, the store of chunk g-1 has surely read its buffer
        // (wait_group.read 1), so that buffer is recycled: the residual box of chunk g-1+B is loaded into it (B-1 chunks
        // of prefetch), or it is marked free.
        if constexpr (kEpi == 4) {
            // Chain: a tile is 4 main chunks (residual prefetched, stored to y, multiplied by warp 3) + 1 tail chunk (the
            // chained GEMM's 64 columns: no residual, stored to the second output).  A main chunk's buffer is recycled when
            // BOTH its store has read it and its UMMAs have completed (chain_done = slot_free barrier).
            const int B = p.epi_bufs, NC = p.n_pad >> 6, NT = p.chain_n >> 6;  // main / tail chunks per tile
            int ld_u = u_first, ld_chunk = 0;
            auto issue_load = [&](int slot) {
                if (ld_u >= num_units) return;
                if (ld_chunk < NC && p.res_tma) {
                    const TileCoord t = decode_tile(p, unit_tile(ld_u));
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&sm.res_full[slot], p.res_bytes);
                        tma_load_4d(sm.epi + slot * kABytes, &tm.res, &sm.res_full[slot], ld_chunk * 64, t.x0, t.y0, t.n);
                    }
                } else if (elect_one()) {  // (also the main chunks of a chain without residual)
                    mbar_arrive(&sm.res_full[slot]);  // tail chunk: nothing to load, the buffer is simply free
                }
                __syncwarp();
                if (++ld_chunk == NC + NT) { ld_chunk = 0; ld_u += u_stride; }
            };
            for (int i = 0; i < B; ++i) issue_load(i);
            int g = 0, slot = 0, rslot = 0;
            uint32_t par = 0, cd_par = 0;  // cd_par: per-buffer parity of the chain_done barrier (bit s)
            bool prev_main = false;
            for (int u = u_first; u < num_units; u += u_stride) {
                const TileCoord t = decode_tile(p, unit_tile(u));
                for (int c = 0; c < NC + NT; ++c, ++g) {
                    mbar_wait(&sm.out_ready[slot], par);
                    if (elect_one()) {
                        if (c < NC) tma_store_4d(&tm.out[0], sm.epi + slot * kABytes, c * 64, t.x0, t.y0, t.n);
                        else tma_store_4d(&tm.out[1], sm.epi + slot * kABytes, (c - NC) * 64, t.x0, t.y0, t.n);
                        bulk_commit();
                        if (g >= 1) bulk_wait_read1();
                    }
                    __syncwarp();
                    if (g >= 1) {
                        if (prev_main) {  // chunk g-1 was multiplied too
                            mbar_wait(&sm.slot_free[rslot], (cd_par >> rslot) & 1u);
                            cd_par ^= 1u << rslot;
                        }
                        issue_load(rslot);
                        if (++rslot == B) rslot = 0;
                    }
                    prev_main = (c < NC);
                    if (++slot == B) { slot = 0; par ^= 1; }
                }
            }
            if (elect_one()) bulk_wait_all();
            __syncwarp();
        } else if (p.epi_tma) {
            const int B = p.epi_bufs, L = (B > 1) ? 1 : 0;
            constexpr bool up2 = kUp2;
            int ld_u = u_first, ld_chunk = 0;
            auto issue_load = [&](int slot) {
                if (ld_u >= num_units) return;
                const TileCoord t = decode_tile(p, unit_tile(ld_u));
                if (elect_one()) {
                    mbar_arrive_expect_tx(&sm.res_full[slot], p.res_bytes);
                    tma_load_4d(sm.epi + slot * kABytes, &tm.res, &sm.res_full[slot], t.nb * p.block_n + ld_chunk * 64,
                                up2 ? (t.x0 >> 1) : t.x0, up2 ? (t.y0 >> 1) : t.y0, t.n);
                }
                __syncwarp();
                if (++ld_chunk == chunks_of(p, t.nb)) { ld_chunk = 0; ld_u += u_stride; }
            };
            if (p.res_tma)
                for (int i = 0; i < B; ++i) issue_load(i);
            int g = 0, slot = 0, rslot = 0;
            uint32_t par = 0;
            for (int u = u_first; u < num_units; u += u_stride) {
                const int tile = unit_tile(u);
                const TileCoord t = decode_tile(p, tile);
                const int nchunks = chunks_of(p, t.nb);
                for (int c = 0; c < nchunks; ++c, ++g) {
                    B200_TIMED_WAIT(5, mbar_wait(&sm.out_ready[slot], par));
                    int col0 = t.nb * p.block_n + c * 64, view = 0;
                    if (p.flags & B200_EPI_SHUFFLE2) {
                        view = col0 / p.shuf_c;
                        col0 -= view * p.shuf_c;
                    }
                    if (elect_one()) {
                        tma_store_4d(&tm.out[view], sm.epi + slot * kABytes, col0, t.x0, t.y0, t.n);
                        bulk_commit();
                        if (g >= L) {
                            B200_TIMED_WAIT(6, if (L) bulk_wait_read1(); else bulk_wait_read0());
                            if (!p.res_tma) mbar_arrive(&sm.slot_free[rslot]);
                        }
                    }
                    __syncwarp();
                    if (g >= L) {
                        if (p.res_tma) issue_load(rslot);
                        if (++rslot == B) rslot = 0;
                    }
                    if (++slot == B) { slot = 0; par ^= 1; }
                }
            }
            if (elect_one()) bulk_wait_all();
            __syncwarp();
        }
    } else if (warp < kEpiWarp0) {
        // ===================================================================== warp 3
        if constexpr (kEpi == 2) {
            // head tail: second UMMA per (tile, sub-pixel): staged bf16 tile x w2 -> 16 TMEM columns (see head_tail_stage)
            const uint32_t idesc16 = umma_idesc_bf16(16u);
            const uint64_t dw = umma_desc_sw128(smem_u32(sm.ss));
            const uint32_t a2 = smem_u32(sm.epi);
            int it = 0;
            for (int u = u_first; u < num_units; u += u_stride, ++it) {
#pragma unroll 1
                for (int sub = 0; sub < 4; ++sub) {
                    const int b = sub * 2 + (it & 1);
                    mbar_wait(&sm.out_ready[b], (it >> 1) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a2 + static_cast<uint32_t>(b) * kABytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16(tmem_base + 256u + static_cast<uint32_t>(b * 16), da + 2 * k, dw + 2 * k, idesc16, k ? 1u : 0u);
                        umma_commit(&sm.slot_free[b]);
                    }
                    __syncwarp();
                }
            }
        }
        if constexpr (kEpi == 4) {
            // chain: y2[128 x chain_n] += staged chunk [128 x 64] x W2[chain_n x 64]^T for the tile's main chunks, in order
            const int B = p.epi_bufs, NC = p.n_pad >> 6, NT = p.chain_n >> 6;
            const uint32_t idesc2 = umma_idesc_bf16(static_cast<uint32_t>(p.chain_n));
            const uint32_t a2 = smem_u32(sm.epi), w2 = smem_u32(sm.epi) - static_cast<uint32_t>(p.chain_bytes);
            mbar_wait(sm.b_ready, 0);
            int it = 0, slot = 0;
            uint32_t par = 0;
            for (int u = u_first; u < num_units; u += u_stride, ++it) {
                mbar_wait(&sm.b_empty[it & 1], ((it >> 1) & 1) ^ 1);  // the tile two back has been drained from these columns
                tc_fence_after();
                const uint32_t d2 = tmem_base + 256u + static_cast<uint32_t>((it & 1) * p.chain_n);
                for (int c = 0; c < NC; ++c) {
                    mbar_wait(&sm.out_ready[slot], par);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a2 + static_cast<uint32_t>(slot) * kABytes);
                    const uint64_t db = umma_desc_sw128(w2 + static_cast<uint32_t>(c * p.chain_n * 128));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) umma_bf16(d2, da + 2 * k, db + 2 * k, idesc2, (c | k) ? 1u : 0u);
                        umma_commit(&sm.slot_free[slot]);            // chain_done: the buffer may be recycled
                        if (c == NC - 1) umma_commit(&sm.b_full[it & 1]);  // d2_full
                    }
                    __syncwarp();
                    if (++slot == B) { slot = 0; par ^= 1; }
                }
                for (int j = 0; j < NT; ++j)  // the tile's tail chunks are not multiplied
                    if (++slot == B) { slot = 0; par ^= 1; }
            }
        }
        // weight producer of the share mode:
        // One stage = the three weight k-blocks (ky = 0..2) of one (chunk, kx); the matrix itself stays (chunk, ky, kx).
        if constexpr (kShare) if (!p.b_resident && warp == 3) {  // (rank 1 of a pair: its idle warp 1 lands in this branch too)
            int s = 0;
            uint32_t ph = 0;
            const int bst = p.b_stages;
            for (int u = u_first; u < num_units; u += u_stride) {
                const int tile = unit_tile(u);
                const int n0 = (tile % p.num_n_blocks) * p.block_n;
                for (int kb9 = 0; kb9 < p.k_blocks; kb9 += 9) {
                    for (int kx = 0; kx < 3; ++kx) {
                        B200_TIMED_WAIT(1, mbar_wait(&sm.b_empty[s], ph ^ 1));
                        if constexpr (kPair) {  // each CTA stages its half of the weight rows, rank 0's barrier counts both
                            if (elect_one()) {
                                if (cta_rank == 0) mbar_arrive_expect_tx(&sm.b_full[s], 2 * 3 * b_bytes);
                                const uint32_t bar0 = mapa_u32(smem_u32(&sm.b_full[s]), 0);
#pragma unroll
                                for (int ky = 0; ky < 3; ++ky)
                                    tma_load_2d_2sm(sm.b + (s * 3 + ky) * b_bytes, &tm.b, bar0, (kb9 + ky * 3 + kx) * kBlockK,
                                                    n0 + static_cast<int>(cta_rank) * p.b_rows);
                            }
                        } else if (elect_one()) {
                            mbar_arrive_expect_tx(&sm.b_full[s], 3 * b_bytes);
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky)
                                tma_load_2d(sm.b + (s * 3 + ky) * b_bytes, &tm.b, &sm.b_full[s],
                                            (kb9 + ky * 3 + kx) * kBlockK, n0);
                        }
                        __syncwarp();
                        if (++s == bst) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp < kBaseWarps) {
        // ===================================================================== epilogue (G groups of 4 warps)
        const int eg = (warp - kEpiWarp0) >> 2;  // epilogue group
        const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const int et = ((warp - kEpiWarp0) & 3) * 32 + lane;  // thread index inside the group
        const int ty = row / p.tile_w;
        const int tx = row - ty * p.tile_w;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const int GA = p.epi_active, B = p.epi_bufs;
        const uint32_t ss_s0 = smem_u32(sm.ss), ss_shift = static_cast<uint32_t>(p.ss_half) * 4;
        // residual row this thread reads: its own pixel, or (y>>1, x>>1) of the half-resolution RES_UP2 box
        constexpr bool up2 = kUp2;
        const int rrow = up2 ? ((ty >> 1) * (p.tile_w >> 1) + (tx >> 1)) : row;
        const uint32_t rswz = static_cast<uint32_t>(rrow & 7);
        if constexpr (kEpi == 2) {
            // head tail: group eg = sub-pixel.  Per tile: stage the BN + ReLU'd bf16 tile for the second UMMA, then finish
            // the PREVIOUS tile (its four dot products have had a whole tile's time to come back).
            const int sub = eg;
            const uint32_t a2 = smem_u32(sm.epi) + static_cast<uint32_t>(row) * 128;
            const uint32_t d2 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256u;
            const uint32_t tmain = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            int itx = 0, pn = 0, poy = 0, pox = 0;
            bool pvalid = false;
            for (int u = u_first; u < num_units; u += u_stride, ++itx) {
                const TileCoord t = decode_tile(p, unit_tile(u));
                const int oy = t.y0 + ty, ox = t.x0 + tx;
                const int b = sub * 2 + (itx & 1);
                mbar_wait(&sm.tfull[0], itx & 1);
                tc_fence_after();
                head_tail_stage(p, tmain, a2 + static_cast<uint32_t>(b) * kABytes, swz, sub, &sm.tempty[0], lane);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.out_ready[b]);
                if (itx > 0) {
                    const int pb = sub * 2 + ((itx - 1) & 1);
                    mbar_wait(&sm.slot_free[pb], ((itx - 1) >> 1) & 1);
                    tc_fence_after();
                    head_tail_store(p, ec.bias, d2 + static_cast<uint32_t>(pb * 16), pn, poy, pox, pvalid, sub);
                }
                pn = t.n; poy = oy; pox = ox;
                pvalid = (oy < p.out_h) && (ox < p.out_w);
            }
            if (itx > 0) {
                const int pb = sub * 2 + ((itx - 1) & 1);
                mbar_wait(&sm.slot_free[pb], ((itx - 1) >> 1) & 1);
                tc_fence_after();
                head_tail_store(p, ec.bias, d2 + static_cast<uint32_t>(pb * 16), pn, poy, pox, pvalid, sub);
            }
        } else {
        int it = 0;
        int g = 0;          // CTA-wide 64-column chunk counter (every group counts the same sequence)
        // Chunk g lives in staging buffer g % B; a buffer ALWAYS belongs to the same group (buffer s -> group s % GA).
        // Round-robin by chunk index (g % GA) let two different groups take consecutive uses of one buffer whenever B is not
        // a multiple of GA, and an mbarrier wait only sees the phase PARITY: the group of use k+1 could reach its wait while
        // the residual box of use k had not landed yet (barrier still in phase k, bit != the parity it waits for -> the wait
        // passes at once) and run over the buffer.  Seen as one CTA's chunk g half wrong and chunk g + B wrong, once in ~1000
        // launches of the 64 -> 256 residual conv (B = 7, GA = 4); found with tools/stress_mix.py.  With one owner per buffer
        // the uses of a buffer are in that group's program order, so a waiter is never a phase ahead.
        int slot = 0, owner = 0;  // buffer of chunk g, the group that owns it, and the parity of its barriers ((g / B) & 1)
        uint32_t par = 0;
        for (int u = u_first; u < num_units; u += u_stride, ++it) {
                const int tile = unit_tile(u); (void)tile;
            const TileCoord t = decode_tile(p, tile);
            const int as = (kEpi == 4) ? 0 : (kShare ? (it & (p.nacc - 1)) : ((!kGather || p.nacc == 2) ? (it & 1) : 0));
            const uint32_t aph = (kEpi == 4) ? (it & 1)
                                 : (kShare ? ((it >> (p.nacc == 4 ? 2 : 1)) & 1) : ((!kGather || p.nacc == 2) ? ((it >> 1) & 1) : (it & 1)));
            const int oy = t.y0 + ty, ox = t.x0 + tx;
            const bool valid = (oy < p.out_h) && (ox < p.out_w);
            B200_TIMED_WAIT(7, mbar_wait(&sm.tfull[as], aph));
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as * p.block_n);
            if (!p.epi_tma) {
                if ((it % G) != eg) {
                    // direct stores: the groups take alternate tiles
                } else {
                    for (int c = 0; c < p.block_n; c += 32) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
                        if (valid) epilogue_direct32(p, v, t.nb * p.block_n + c, t.n, oy, ox);
                    }
                }
            } else if (eg < GA) {
                const int nchunks = chunks_of(p, t.nb);
                const uint32_t ss_s = ss_s0 + static_cast<uint32_t>(t.nb * p.block_n) * 4;
                for (int c = 0; c < nchunks; ++c, ++g) {
                    const int my_slot = slot;
                    const uint32_t my_par = par;
                    const bool is_mine = (owner == eg);
                    if (++owner == GA) owner = 0;
                    if (++slot == B) { slot = 0; owner = 0; par ^= 1; }
                    if (!is_mine) continue;
                    const uint32_t buf = smem_u32(sm.epi) + static_cast<uint32_t>(my_slot) * kABytes;
                    const uint32_t mine = buf + static_cast<uint32_t>(row) * 128;
                    const uint32_t rsd = buf + static_cast<uint32_t>(rrow) * 128;
                    B200_TIMED_WAIT(8, if (p.res_tma || kEpi == 4) mbar_wait(&sm.res_full[my_slot], my_par); else mbar_wait(&sm.slot_free[my_slot], my_par ^ 1));
                    uint4 rup[kUp2 ? 8 : 1];
                    if constexpr (up2) {  // the 4 KB half-resolution box is shared by four pixels: read it before anyone overwrites
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            rup[i] = lds128(rsd + ((static_cast<uint32_t>(i) ^ rswz) << 4));
                        named_bar_sync(1 + eg, kEpiThreads);
                    }
                    if (!(p.dbg & 4))
#pragma unroll
                    for (int h = 0; h < 2; ++h) {  // two 32-column halves (register-lean: 19-23 warps per CTA)
                        uint32_t v[32];
                        tmem_ld32(taddr + c * 64 + h * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // 8 columns = one 16-byte piece of the pixel's 128-byte row
                            const int piece = h * 4 + j;
                            const uint32_t so = (static_cast<uint32_t>(piece) ^ swz) << 4;
                            float f[8];
                            const uint32_t sca = ss_s + static_cast<uint32_t>(c * 64 + piece * 8) * 4;
                            const float4 sc0 = lds_f4(sca), sc1 = lds_f4(sca + 16);
                            const float4 sh0 = lds_f4(sca + ss_shift), sh1 = lds_f4(sca + ss_shift + 16);
                            f[0] = fmaf(__uint_as_float(v[j * 8 + 0]), sc0.x, sh0.x);
                            f[1] = fmaf(__uint_as_float(v[j * 8 + 1]), sc0.y, sh0.y);
                            f[2] = fmaf(__uint_as_float(v[j * 8 + 2]), sc0.z, sh0.z);
                            f[3] = fmaf(__uint_as_float(v[j * 8 + 3]), sc0.w, sh0.w);
                            f[4] = fmaf(__uint_as_float(v[j * 8 + 4]), sc1.x, sh1.x);
                            f[5] = fmaf(__uint_as_float(v[j * 8 + 5]), sc1.y, sh1.y);
                            f[6] = fmaf(__uint_as_float(v[j * 8 + 6]), sc1.z, sh1.z);
                            f[7] = fmaf(__uint_as_float(v[j * 8 + 7]), sc1.w, sh1.w);
                            if constexpr (kEpi == 3) {
#pragma unroll
                                for (int i = 0; i < 8; i += 2) gelu_as2(f[i], f[i + 1]);
                            }
                            if (p.res_tma) {
                                uint4 u;
                                if constexpr (up2) u = rup[piece];
                                else u = lds128(mine + so);  // residual of this very pixel, in place
                                f[0] += bf16lo(u.x); f[1] += bf16hi(u.x);
                                f[2] += bf16lo(u.y); f[3] += bf16hi(u.y);
                                f[4] += bf16lo(u.z); f[5] += bf16hi(u.z);
                                f[6] += bf16lo(u.w); f[7] += bf16hi(u.w);
                            }
                            if (p.flags & B200_EPI_RELU) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.0f);
                            }
                            uint4 o;
                            o.x = pack_bf16(f[0], f[1]);
                            o.y = pack_bf16(f[2], f[3]);
                            o.z = pack_bf16(f[4], f[5]);
                            o.w = pack_bf16(f[6], f[7]);
                            sts128(mine + so, o);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.out_ready[my_slot]);
                }
            }
            // accumulator fully read by this warp (or not needed): hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&sm.tempty[as]), 0));
                else mbar_arrive(&sm.tempty[as]);
            }
            if constexpr (kEpi == 4) {
                // tail chunks of the tile: the chained GEMM's chain_n columns (64 each), BN + ReLU, bf16, staged for the DMA warp
                if (eg < GA)
                for (int tj = 0; tj < (p.chain_n >> 6); ++tj) {
                    const int my_slot = slot;
                    const uint32_t my_par = par;
                    const bool is_mine = (owner == eg);
                    if (++owner == GA) owner = 0;
                    if (++slot == B) { slot = 0; owner = 0; par ^= 1; }
                    ++g;
                    if (is_mine) {
                        const uint32_t mine = smem_u32(sm.epi) + static_cast<uint32_t>(my_slot) * kABytes + static_cast<uint32_t>(row) * 128;
                        mbar_wait(&sm.res_full[my_slot], my_par);           // the buffer is free (plain arrival of the DMA warp)
                        mbar_wait(&sm.b_full[it & 1], (it >> 1) & 1);       // d2_full
                        tc_fence_after();
                        const uint32_t t2 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256u +
                                            static_cast<uint32_t>((it & 1) * p.chain_n + tj * 64);
                        const uint32_t ss2 = ss_s0 + static_cast<uint32_t>(p.n_pad + tj * 64) * 4;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v[32];
                            tmem_ld32(t2 + h * 32, v);
                            tmem_ld_wait();
                            if (h == 1) {  // accumulator columns read: the tile after next may overwrite them
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&sm.b_empty[it & 1]);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int piece = h * 4 + j;
                                const uint32_t sca = ss2 + static_cast<uint32_t>(piece * 8) * 4;
                                const float4 sc0 = lds_f4(sca), sc1 = lds_f4(sca + 16);
                                const float4 sh0 = lds_f4(sca + ss_shift), sh1 = lds_f4(sca + ss_shift + 16);
                                float f[8];
                                f[0] = fmaf(__uint_as_float(v[j * 8 + 0]), sc0.x, sh0.x);
                                f[1] = fmaf(__uint_as_float(v[j * 8 + 1]), sc0.y, sh0.y);
                                f[2] = fmaf(__uint_as_float(v[j * 8 + 2]), sc0.z, sh0.z);
                                f[3] = fmaf(__uint_as_float(v[j * 8 + 3]), sc0.w, sh0.w);
                                f[4] = fmaf(__uint_as_float(v[j * 8 + 4]), sc1.x, sh1.x);
                                f[5] = fmaf(__uint_as_float(v[j * 8 + 5]), sc1.y, sh1.y);
                                f[6] = fmaf(__uint_as_float(v[j * 8 + 6]), sc1.z, sh1.z);
                                f[7] = fmaf(__uint_as_float(v[j * 8 + 7]), sc1.w, sh1.w);
                                if (p.chain_relu) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.0f);
                                }
                                uint4 o;
                                o.x = pack_bf16(f[0], f[1]);
                                o.y = pack_bf16(f[2], f[3]);
                                o.z = pack_bf16(f[4], f[5]);
                                o.w = pack_bf16(f[6], f[7]);
                                sts128(mine + ((static_cast<uint32_t>(piece) ^ swz) << 4), o);
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.out_ready[my_slot]);
                    }
                }
            }
        }
        }  // kEpi != 2
    } else if (kGather) {
        // ===================================================================== gather producers (16 warps)
        constexpr int kPasses = kBlockM / (kGatherWarps * 4);  // 2: a pass = 64 rows, 8 lanes x 16 B per row
        const int pw = warp - kBaseWarps;
        const int sub = lane >> 3;
        const int l8 = lane & 7;
        int s = 0;
        uint32_t ph = 0;
        for (int u = u_first; u < num_units; u += u_stride) {
                const int tile = unit_tile(u);
            const TileCoord t = decode_tile(p, tile);
            int oy[kPasses], ox[kPasses];
            bool valid[kPasses];
            uint32_t soff[kPasses];
#pragma unroll
            for (int q = 0; q < kPasses; ++q) {
                const int r = q * 64 + pw * 4 + sub;
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
                    mbar_wait(&sm.a_empty[s], ph ^ 1);
                    uint8_t* a = sm.a + s * kABytes;
#pragma unroll
                    for (int q = 0; q < kPasses; ++q) *reinterpret_cast<uint4*>(a + soff[q]) = u[q];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.a_full[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            } else {
                // modulated deformable sampling from b200_dcn_coord records.  Unit = (k-block, pass); the corner
                // loads of unit u+1 and the record of unit u+2 are in flight while unit u is blended.
                const __nv_bfloat16* img = p.x + static_cast<int64_t>(t.n) * p.in_h * p.in_w * p.x_ld + l8 * 8;
                const int cs = (p.in_w > 1) ? p.x_ld : 0;            // element step to the block's right column
                const int rs = (p.in_h > 1) ? p.in_w * p.x_ld : 0;   // element step to the block's bottom row
                const uint4* rec[kPasses];
#pragma unroll
                for (int q = 0; q < kPasses; ++q) {
                    const int64_t pix = valid[q] ? (static_cast<int64_t>(t.n) * p.out_h + oy[q]) * p.out_w + ox[q] : 0;
                    rec[q] = reinterpret_cast<const uint4*>(p.offsets) + pix * 9;
                }
                // K is ordered (64-channel chunk, tap): the 9 taps x 4 corners of one chunk re-touch the same
                // neighbourhood back to back, so they hit in L1 (the kernel keeps its smem carve-out small).
                // Unit = (k-block, pass).  While unit u is blended, the four corner loads of unit u+1 and the record of
                // unit u+2 are in flight.  The loop is unrolled over 3 k-blocks = 6 units so the record ring (3 slots)
                // and the corner ping-pong (2 slots) are addressed with compile-time indices (k_blocks = 9 * chunks).
                uint4 rcd[3];
                uint4 dat[2][4];
                rcd[0] = ldg128(rec[0]);
                rcd[1] = ldg128(rec[1]);
                {
                    const __nv_bfloat16* b0 = img + static_cast<int>(rcd[0].x);
                    dat[0][0] = ldg128(b0);
                    dat[0][1] = ldg128(b0 + cs);
                    dat[0][2] = ldg128(b0 + rs);
                    dat[0][3] = ldg128(b0 + rs + cs);
                }
                const uint32_t a_base = smem_u32(sm.a);
                int tap0 = 0;                    // tap of k-block kb0 (0, 3 or 6)
                const __nv_bfloat16* src = img;  // channel chunk of k-block kb0
                for (int kb0 = 0; kb0 < p.k_blocks; kb0 += 3) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const int q = j & 1;
                        {   // record of unit u+2: k-block kb0 + (j+2)/2, same pass
                            const int dk = (j + 2) >> 1;
                            if (kb0 + dk < p.k_blocks) {
                                int tp = tap0 + dk;
                                tp = (tp >= 9) ? tp - 9 : tp;
                                rcd[(j + 2) % 3] = ldg128(rec[q] + tp);
                            }
                        }
                        {   // corners of unit u+1: k-block kb0 + (j+1)/2, pass (j+1)&1
                            const int dk = (j + 1) >> 1;
                            if (kb0 + dk < p.k_blocks) {
                                const __nv_bfloat16* b0 =
                                    ((tap0 + dk >= 9) ? src + kBlockK : src) + static_cast<int>(rcd[(j + 1) % 3].x);
                                dat[(j + 1) & 1][0] = ldg128(b0);
                                dat[(j + 1) & 1][1] = ldg128(b0 + cs);
                                dat[(j + 1) & 1][2] = ldg128(b0 + rs);
                                dat[(j + 1) & 1][3] = ldg128(b0 + rs + cs);
                            }
                        }
                        // blend unit u with packed bf16x2 FMAs (one bf16 rounding per FMA)
                        const uint32_t w01 = valid[q] ? rcd[j % 3].y : 0u, w23 = valid[q] ? rcd[j % 3].z : 0u;
                        const __nv_bfloat162 wa = __low2bfloat162(bits_bf162(w01)), wb = __high2bfloat162(bits_bf162(w01));
                        const __nv_bfloat162 wc = __low2bfloat162(bits_bf162(w23)), wd = __high2bfloat162(bits_bf162(w23));
                        const uint4 (&d)[4] = dat[j & 1];
                        uint4 res;
#define B200_BLEND2(f) bf162_bits(__hfma2(wd, bits_bf162(d[3].f), __hfma2(wc, bits_bf162(d[2].f), \
                           __hfma2(wb, bits_bf162(d[1].f), __hmul2(wa, bits_bf162(d[0].f))))))
                        res.x = B200_BLEND2(x);
                        res.y = B200_BLEND2(y);
                        res.z = B200_BLEND2(z);
                        res.w = B200_BLEND2(w);
#undef B200_BLEND2
                        if (q == 0) mbar_wait(&sm.a_empty[s], ph ^ 1);
                        sts128(a_base + static_cast<uint32_t>(s) * kABytes + soff[q], res);
                        if (q == 1) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sm.a_full[s]);
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
                    tap0 += 3;
                    if (tap0 == 9) { tap0 = 0; src += kBlockK; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef B200_PROF
    if ((p.dbg & 16) && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_prof[9], (unsigned long long)(clock64() - prof_t0));
#endif
    if constexpr (kPair) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair's MMAs / multicasts may still touch it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm(tmem_base, tmem_cols);
        else tmem_dealloc(tmem_base, tmem_cols);
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

int encode_map_typed(CUtensorMap* m, int dtype, int swizzle128, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const char* what) {
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
    CUresult r = enc(m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                     const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return 0;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const char* what) {
    return encode_map_typed(m, 0, 1, base, rank, dims, strides_bytes, box, what);
}

// NHWC bf16 tensor (n, h, w, ld) seen as a 4-D TMA tensor over `c` channels with a {64, tile_w, tile_h, 1} box
static int encode_nhwc(CUtensorMap* m, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t n, uint64_t ld,
                       uint64_t xstep, uint64_t ystep, uint64_t row_pixels, uint64_t img_pixels, int tile_w, int tile_h,
                       const char* what) {
    const uint64_t dims[4] = {c, w, h, n};
    const uint64_t str[3] = {xstep * ld * 2, ystep * row_pixels * ld * 2, img_pixels * ld * 2};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)tile_w, (uint32_t)tile_h, 1u};
    return encode_map(m, base, 4, dims, str, box, what);
}

}  // namespace b200

using namespace b200;

struct b200_igemm {
    TmapPack tm;
    IgemmParams p;
    int pair;  // cta_group::2 launch: clusters of two CTAs
    int gather;
    int grid;
    int threads;
    int smem_bytes;
    Dcn2Plan* d2;  // B200_IGEMM_DCN_TMEM: the launch belongs to dcn2.cu
    HeadConsts* hc;  // B200_EPI_HEAD_DOT: constant-bank payload (host copy, passed by value at launch)
};

static int pick_tile_w(int out_h, int out_w, int max_tw) {
    int best = max_tw;
    long best_area = -1;
    for (int tw = max_tw; tw >= 8; tw >>= 1) {
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
    if (d->mode < B200_IGEMM_CONV || d->mode > B200_IGEMM_DCN_TMEM) return set_error("igemm: bad mode %d", d->mode);
    if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0) return set_error("igemm: bad shape");
    if (!d->x || !d->w || !d->scale || !d->shift || !d->y) return set_error("igemm: null tensor pointer");
    if (d->mode == B200_IGEMM_DCN_TMEM) {
        if (d->stride != 1 && d->stride != 2) return set_error("igemm: stride must be 1 or 2");
        if (d->stride == 2 && ((d->in_h | d->in_w) & 1)) return set_error("igemm: stride 2 needs even input size");
        if ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->w) | reinterpret_cast<uintptr_t>(d->y) |
             reinterpret_cast<uintptr_t>(d->scale) | reinterpret_cast<uintptr_t>(d->shift) | reinterpret_cast<uintptr_t>(d->offsets)) & 15)
            return set_error("igemm: all tensor pointers must be 16-byte aligned");
        if ((d->out_ld % 8) || (d->x_ld % 8)) return set_error("igemm: out_ld / x_ld must be multiples of 8");
        Dcn2Plan* d2 = nullptr;
        if (dcn2_create(d, &d2)) return 1;
        b200_igemm* op = new (std::nothrow) b200_igemm;
        if (!op) {
            dcn2_destroy(d2);
            return set_error("igemm: out of host memory");
        }
        memset(op, 0, sizeof(*op));
        op->d2 = d2;
        op->p.out_h = d2->p.out_h; op->p.out_w = d2->p.out_w; op->p.tile_w = d2->p.tile_w; op->p.tile_h = d2->p.tile_h;
        op->p.block_n = d2->p.block_n; op->p.num_tiles = d2->p.num_tiles;
        op->smem_bytes = d2->smem_bytes;
        *out = op;
        return 0;
    }
    if (d->out_c <= 0 || (d->out_c % 8) != 0) return set_error("igemm: out_c must be a positive multiple of 8");
    if (d->stride != 1 && d->stride != 2) return set_error("igemm: stride must be 1 or 2");
    const int stride_w = d->stride_w ? d->stride_w : d->stride;
    if (stride_w != d->stride && !(d->stride == 2 && stride_w == 1 && d->mode == B200_IGEMM_CONV && d->ksize == 3))
        return set_error("igemm: stride_w is for the 3x3 stride (2, 1) convolution only");
    int taps;
    if (d->mode == B200_IGEMM_STEM) {
        if (d->ksize != 7 || d->stride != 2 || (d->in_h & 1) || (d->in_w & 1))
            return set_error("igemm: STEM is 7x7 stride 2 on even-sized images");
        taps = 4;
    } else {
        if (d->ksize != 1 && d->ksize != 3) return set_error("igemm: ksize must be 1 or 3");
        if (d->mode == B200_IGEMM_DCN && d->ksize != 3) return set_error("igemm: DCN is 3x3");
        if (d->in_c <= 0 || (d->in_c % 64) != 0) return set_error("igemm: in_c must be a multiple of 64");
        if (d->stride == 2 && ((d->in_h & 1) || (stride_w == 2 && (d->in_w & 1))))
            return set_error("igemm: stride 2 needs even input size");
        taps = d->ksize * d->ksize;
    }
    const int x_ld = d->x_ld ? d->x_ld : d->in_c;
    if (d->mode != B200_IGEMM_STEM && (x_ld < d->in_c || (x_ld % 8))) return set_error("igemm: bad x_ld");
    if (d->mode == B200_IGEMM_DCN && !d->offsets) return set_error("igemm: DCN needs offsets");
    if (d->x2) {
        const int x2_ld = d->x2_ld ? d->x2_ld : d->in_c2;
        if (d->mode != B200_IGEMM_CONV || d->ksize != 1 || d->stride != 1 || d->in_c2 <= 0 || (d->in_c2 % 64) != 0 ||
            (d->stride2 != 1 && d->stride2 != 2) || x2_ld < d->in_c2 || (x2_ld % 8) || (reinterpret_cast<uintptr_t>(d->x2) & 15) ||
            (d->flags & (B200_EPI_GELU | B200_EPI_HEAD_DOT | B200_EPI_DCN_COORDS)))
            return set_error("igemm: a second K segment (x2) is for 1x1 stride-1 CONV launches, in_c2 a multiple of 64, stride2 1 or 2");
    }
    if ((d->flags & (B200_EPI_RES | B200_EPI_RES_UP2)) && !d->residual) return set_error("igemm: residual missing");
    if ((d->flags & B200_EPI_HEAD_DOT) &&
        (!(d->flags & B200_EPI_SHUFFLE2) || d->shuf_c != 64 || d->out_c != 256 || d->mode != B200_IGEMM_CONV || !d->offsets ||
         (d->flags & (B200_EPI_RES | B200_EPI_RES_UP2 | B200_EPI_OUT_F32))))
        return set_error("igemm: HEAD_DOT is for the 64 -> 4x64 transposed conv with the 257-float tail weights in `offsets`");
    if ((d->flags & B200_EPI_GELU) &&
        (d->mode != B200_IGEMM_CONV || (d->out_c % 64) != 0 ||
         (d->flags & (B200_EPI_RELU | B200_EPI_RES_UP2 | B200_EPI_OUT_F32 | B200_EPI_SHUFFLE2 | B200_EPI_DCN_COORDS | B200_EPI_HEAD_DOT))))
        return set_error("igemm: GELU is for plain bf16 conv outputs with out_c a multiple of 64 (optionally + RES)");
    if (d->flags & B200_EPI_CHAIN) {
        if (d->mode != B200_IGEMM_CONV || d->ksize != 1 || d->stride != 1 || d->out_c != 256 || d->n_pad != 256 || (d->chain_n != 64 && d->chain_n != 128) ||
            !d->chain_w || !d->chain_scale || !d->chain_shift || !d->chain_y || (d->chain_ld % 8) ||
            d->chain_ld < d->chain_n || d->in_c + (d->x2 ? d->in_c2 : 0) > 256 ||
            (d->flags & (B200_EPI_RES_UP2 | B200_EPI_OUT_F32 | B200_EPI_SHUFFLE2 | B200_EPI_DCN_COORDS | B200_EPI_HEAD_DOT | B200_EPI_GELU)))
            return set_error("igemm: CHAIN is for a 1x1 stride-1 conv with out_c == n_pad == 256, in_c <= 256 and chain_n 64 or 128");
        if ((reinterpret_cast<uintptr_t>(d->chain_w) | reinterpret_cast<uintptr_t>(d->chain_y) | reinterpret_cast<uintptr_t>(d->chain_scale) |
             reinterpret_cast<uintptr_t>(d->chain_shift)) & 15)
            return set_error("igemm: chain pointers must be 16-byte aligned");
    }
    if ((d->flags & B200_EPI_DCN_COORDS) && (d->out_c != 32 || d->ksize != 3 || d->mode != B200_IGEMM_CONV))
        return set_error("igemm: DCN_COORDS is for the 3x3 offset conv with out_c == 32");
    int block_n = d->block_n;
    if (block_n == 0) {
        block_n = 256;
        while (block_n > 32 && block_n / 2 >= d->n_pad) block_n >>= 1;
        if (d->n_pad % block_n) block_n = 32;
    }
    // DCN with 512 output channels: one 128 x 512 tile fills TMEM (single accumulator) but every pixel is gathered
    // once instead of once per 256-column block -- the gather, not the MMA, is what this kernel is short of
    if (d->block_n == 0 && d->mode == B200_IGEMM_DCN && d->n_pad == 512 && !getenv("B200_DCN_N256")) block_n = 512;
    if (block_n != 32 && block_n != 64 && block_n != 128 && block_n != 256 && !(block_n == 512 && d->mode == B200_IGEMM_DCN))
        return set_error("igemm: block_n must be 32, 64, 128 or 256 (512 for DCN)");
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
    p.x_ld = x_ld;
    p.out_h = d->in_h / d->stride;
    p.out_w = d->in_w / stride_w;
    p.stride_w = stride_w;
    p.out_c = d->out_c;
    p.ksize = d->ksize;
    p.stride = d->stride;
    p.flags = d->flags;
    p.block_n = block_n;
    p.n_pad = d->n_pad;
    p.ss_half = (d->flags & B200_EPI_HEAD_DOT) ? 2048 : (d->n_pad > 256 ? d->n_pad : 256);  // HEAD_DOT: 16 KB of UMMA constants
    if (d->flags & B200_EPI_CHAIN) {
        p.ss_half = 512;  // the chained GEMM's scale / shift follow the layer's own 256
        p.chain_n = d->chain_n;
        p.chain_bytes = (d->n_pad / 64) * d->chain_n * 128;
        p.chain_relu = d->chain_relu;
        p.chain_scale = d->chain_scale;
        p.chain_shift = d->chain_shift;
    }
    p.nacc = block_n > 256 ? 1 : 2;
    // 3x3 stride-1 convs share one (tile_h+2)-row activation box between the taps ky = 0..2 of a kernel column
    // ("share mode"): tall tiles (8 rows x 16 pixels) make that 2.4x less L2->SM traffic and a third of the barrier
    // round trips (a hand-shake costs ~380 cycles whatever it carries, a 128 x N x 64 MMA block only 2*N).
    // Only narrow GEMMs (N <= 128) take it: at N = 256 a k-block is 512 MMA cycles and hides everything already.
    // ... except N = 256 layers that will run on a CTA pair (K >= 16 k-blocks, see `pair` below): what bounds those is the L2 ->
    // SM fabric (~13 TB/s when every SM pulls; 3x3 256 -> 256 at 256^2 pulled 12.5), and the shared box cuts the activation
    // part from 9 to 3.75 tiles per chunk: 0.772 -> 0.739 ms.
    const bool share = (d->mode == B200_IGEMM_CONV && d->ksize == 3 && d->stride == 1 &&
                        (block_n <= 128 ||
                         (!getenv("B200_NO_SHARE_N256") && block_n == 256 && d->in_c >= 128 && (d->out_c % 64) == 0 &&
                          d->batch * d->in_h * d->in_w > kBlockM /* at least two pixel tiles */ && !getenv("B200_PAIR_MIN_KB") &&
                          (!getenv("B200_PAIR") || atoi(getenv("B200_PAIR")) != 0) &&
                          !(d->flags & (B200_EPI_GELU | B200_EPI_HEAD_DOT | B200_EPI_CHAIN | B200_EPI_RES_UP2 | B200_EPI_OUT_F32 |
                                        B200_EPI_DCN_COORDS | B200_EPI_SHUFFLE2)))) &&
                        !getenv("B200_NO_SHARE"));
    // RES_UP2 reads a half-resolution residual box: the tile needs at least two rows
    // The stem (4x4 window over the space-to-depth buffer) works the same way: the four window rows of a tile are
    // tile_h+3 consecutive rows of the overlapping-stride view, one box and one barrier round trip per tile.
    bool stem_share = (d->mode == B200_IGEMM_STEM && !getenv("B200_STEM_GATHER") && !getenv("B200_NO_SHARE"));
    int max_tw = (d->flags & B200_EPI_RES_UP2) ? 64 : 128;
    if (share || stem_share) max_tw = 16;
    // the deformable gathers live on L1 hits: a squarer tile has a smaller neighbourhood (layer2 128x128: -9 %)
    if (d->mode == B200_IGEMM_DCN) max_tw = (d->in_w / d->stride >= 128) ? 16 : 32;
    if (const char* e = getenv("B200_CONV_TW")) { if (atoi(e) >= 8 && (share || stem_share || atoi(e) < max_tw)) max_tw = atoi(e); }
    if (const char* e = getenv("B200_DEBUG")) p.dbg = atoi(e);
    p.tile_w = pick_tile_w(p.out_h, p.out_w, max_tw);
    p.tile_h = kBlockM / p.tile_w;
    p.a_kb = share ? 3 : (stem_share ? 4 : 1);
    p.a_bytes = share ? (p.tile_h + 2) * p.tile_w * 128 : (stem_share ? (p.tile_h + 3) * p.tile_w * 128 : kABytes);
    p.a_shift = (share || stem_share) ? p.tile_w * 128 : 0;
    p.b_kb = 1;
    p.tiles_x = (p.out_w + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (p.out_h + p.tile_h - 1) / p.tile_h;
    p.num_n_blocks = d->n_pad / block_n;
    p.num_tiles = p.batch * p.tiles_x * p.tiles_y * p.num_n_blocks;
    p.cchunks = (d->mode == B200_IGEMM_STEM) ? 1 : (d->in_c + (d->x2 ? d->in_c2 : 0)) / kBlockK;
    p.kb_split = d->x2 ? d->in_c / kBlockK : 0;
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
    if (d->mode == B200_IGEMM_STEM) {
        // Preferred stem path: no producer warps at all.  A row (pixel, window row ty) = 64 contiguous bf16 of the packed
        // buffer starting at block (oy+ty, ox), i.e. a 4-D tensor {64, W/2, H/2+3, N} whose x stride is ONE block
        // (32 bytes): consecutive rows overlap by three blocks.  TMA only needs 16-byte-multiple strides; if the
        // driver refuses the overlapping view, fall back to the producer-warp copy.
        const uint64_t dims[4] = {64, (uint64_t)(d->in_w / 2), (uint64_t)(d->in_h / 2 + 3), (uint64_t)d->batch};
        const uint64_t pkw = d->in_w / 2 + 3, pkh = d->in_h / 2 + 3;
        const uint64_t str[3] = {32, pkw * 32, pkh * pkw * 32};
        const uint32_t box[4] = {64u, (uint32_t)p.tile_w, (uint32_t)(stem_share ? p.tile_h + 3 : p.tile_h), 1u};
        if (!getenv("B200_STEM_GATHER") && encode_map(&op->tm.a[0], d->x, 4, dims, str, box, "overlapping stem view") == 0) {
            op->gather = 0;
        } else {
            stem_share = false;
            p.a_kb = 1;
            p.a_bytes = kABytes;
            p.a_shift = 0;
        }
    }
    const bool fat = p.a_kb > 1;  // an A stage carries several k-blocks (share variant of the kernel)

    // bf16 outputs whose channel count is a multiple of 64 go out through swizzled staging + TMA stores
    p.epi_tma = !(d->flags & (B200_EPI_OUT_F32 | B200_EPI_DCN_COORDS | B200_EPI_HEAD_DOT)) && (d->out_c % 64) == 0 &&
                (!(d->flags & B200_EPI_SHUFFLE2) || (d->shuf_c % 64) == 0);
    p.res_tma = p.epi_tma && (d->flags & (B200_EPI_RES | B200_EPI_RES_UP2)) && !(d->flags & B200_EPI_SHUFFLE2);
    if ((d->flags & (B200_EPI_RES | B200_EPI_RES_UP2)) && !p.res_tma && p.epi_tma) {
        delete op;
        return set_error("igemm: residual with a pixel-shuffle store is not supported");
    }
    if ((d->flags & B200_EPI_RES_UP2) && ((p.out_h | p.out_w) & 1)) {
        delete op;
        return set_error("igemm: RES_UP2 needs even output size");
    }
    p.res_bytes = (d->flags & B200_EPI_RES_UP2) ? kABytes / 4 : kABytes;

    const int G = epi_groups(op->gather);
    int b_bytes = block_n * 128;
    // Small weight matrices (stem, layer1, ConvT, ...) live in shared memory for the whole launch: no per-tile weight
    // TMA traffic, the ring carries activations only.
    p.b_resident = (p.num_n_blocks == 1 && p.k_blocks * b_bytes <= 72 * 1024 && d->mode != B200_IGEMM_DCN) ? 1 : 0;
    // Pair mode (cta_group::2): layers whose weights stream per tile and whose k-blocks are "fat" in weights relative to
    // activations -- what each SM pulls through the L2 fabric per k-block (A + B, ~42 B/clk per SM when every SM pulls)
    // is what bounds them, and the pair halves the B part.  TMA-fed plain convs with the TMA-store epilogue only.
    const int px_tiles = p.batch * p.tiles_x * p.tiles_y;
    // Measured per layer (batch 16 @1024^2, pair on vs off): 3x3 256->256 at 256^2 0.851 -> 0.769 ms (1.6 PFLOP/s: the SS
    // operand reads, A + B per instruction, drop from 12 to 8 KB), 128^2 0.209 -> 0.189, lat3 (K = 2048) 0.030 -> 0.027;
    // but short-K layers lose to the cross-CTA hand-shakes per tile (conv3 with K = 256: 0.072 -> 0.079, lat0 0.239 ->
    // 0.275, K = 512 layers 0.065 -> 0.072), and K = 1024..2048 1x1 convs do not move.  Hence: at least 12 k-blocks per tile
    // (12 = the 256 + 512 -> 1024 conv3 that carries its projection shortcut as second K segment: 0.096 -> 0.087 ms; without
    // the pair it pulled 94 B/clk per SM through the L2 fabric).
    int pair_min_kb = 12;
    if (const char* e = getenv("B200_PAIR_MIN_KB")) pair_min_kb = atoi(e);
    // Share-mode layers (3x3 stride 1, N = 128) pair as well: the DB head's 256 -> 128 conv pulled 816 KB per tile through the
    // L2 fabric (weights 576 KB of it) at the ~13 TB/s the fabric gives when every SM pulls; the pair halves the weight part.
    bool pair = (d->mode == B200_IGEMM_CONV && (!fat || (share && !(d->flags & B200_EPI_RES_UP2) && !getenv("B200_NO_PAIR_SHARE"))) &&
                 !p.b_resident && p.epi_tma && block_n >= 128 && px_tiles >= 2 &&
                 !(d->flags & (B200_EPI_GELU | B200_EPI_HEAD_DOT | B200_EPI_CHAIN)) && p.k_blocks >= pair_min_kb);
    if (const char* e = getenv("B200_PAIR")) pair = pair && atoi(e) != 0;
    if (share && block_n > 128 && !pair) {
        delete op;
        return set_error("igemm: internal: an N = 256 share-mode plan must run on a CTA pair");
    }
    op->pair = pair ? 1 : 0;
    p.b_rows = pair ? block_n / 2 : block_n;
    p.num_units = pair ? ((px_tiles + 1) / 2) * p.num_n_blocks : p.num_tiles;
    if (pair) b_bytes = p.b_rows * 128;
    p.ss_shared = 1;
    if ((d->flags & B200_EPI_CHAIN) && !p.b_resident) {
        delete op;
        return set_error("igemm: CHAIN needs resident weights");
    }
    const int fixed = 1024 /*alignment*/ + 1024 /*barriers*/ + 8 * p.ss_half /*scale+shift*/ +
                      (p.b_resident ? p.k_blocks * b_bytes : 0) + p.chain_bytes;
    // The DCN gather lives on L1 hits (K ordered chunk-major), so it keeps its carve-out <= 132 KB and leaves the
    // rest of the 256 KB to L1; everything else may use the full 227 KB opt-in.
    int cap = (d->mode == B200_IGEMM_DCN) ? (block_n > 256 ? 227 : 131) * 1024 : kMaxSmem;
    if (const char* e = getenv("B200_DCN_CAP")) { if (d->mode == B200_IGEMM_DCN && atoi(e) >= 100 && atoi(e) <= 227) cap = atoi(e) * 1024; }
    const int a_bytes = p.a_bytes;
    int max_stages = kMaxStages;
    if (const char* e = getenv("B200_STAGES")) { if (atoi(e) >= 2 && atoi(e) < max_stages) max_stages = atoi(e); }
    // Ring depths for `avail` bytes; returns false if the minimum does not fit.
    int a_st = 0, b_st = 0, b_kb = 1;
    auto rings = [&](int avail) -> bool {
        b_kb = 1;
        if (p.b_resident) {
            a_st = avail / a_bytes;
            b_st = 0;
            if (a_st > max_stages) a_st = max_stages;
            return a_st >= 2;
        }
        if (!fat) {  // one A tile and one weight k-block per stage, rings of equal depth
            a_st = b_st = avail / (a_bytes + b_bytes);
            if (a_st > max_stages) a_st = b_st = max_stages;
            return a_st >= 2;
        }
        // share mode: an A stage feeds three k-blocks and so does a weight stage (three k-blocks of block_n rows)
        const int n = avail / (a_bytes + 3 * b_bytes);
        if (n < 2) return false;
        b_kb = 3;
        a_st = b_st = (n > max_stages) ? max_stages : n;
        if (avail - a_st * (a_bytes + 3 * b_bytes) >= a_bytes && a_st < max_stages) ++a_st;
        return true;
    };
    int bufs = 0, ga = 1;
    bool ok;
    if (d->flags & B200_EPI_HEAD_DOT) {
        bufs = 8;  // staging tiles of the second UMMA: 4 sub-pixels x 2 (see head_tail_stage)
        ok = rings(cap - fixed - bufs * kABytes);
    } else if (!p.epi_tma) {
        ok = rings(cap - fixed);
    } else {
        // Split the shared memory between mainloop stages and epilogue buffers by where the time goes:
        // mainloop ~ k_blocks * 2 * block_n cycles per tile (tcgen05 issue rate), epilogue ~ 1600 cycles per 64-column
        // chunk and group.  Enough groups to keep up with the mainloop, one buffer per group plus one draining (plus
        // two more of residual prefetch), the rest to the pipeline stages.
        const int chunks = (block_n < d->out_c ? block_n : d->out_c) / 64;
        const int main_clk = p.k_blocks * 2 * block_n, epi_clk = chunks * ((d->flags & B200_EPI_GELU) ? 2600 : 1600);
        ga = (epi_clk + main_clk - 1) / main_clk;
        if (ga > G) ga = G;
        if (ga < 1) ga = 1;
        bufs = (ga == 1 && !p.res_tma) ? 1 : ga + 1 + (p.res_tma ? 2 : 0);
        ok = rings(cap - fixed - bufs * kABytes);
        const int want = (p.k_blocks >= 4 && !fat) ? 3 : 2;
        while ((!ok || a_st < want) && bufs > 1) {
            --bufs;
            ok = rings(cap - fixed - bufs * kABytes);
        }
        if (ok) {  // leftovers deepen the residual prefetch
            const int used = a_st * a_bytes + (p.b_resident ? 0 : (fat ? b_st * 3 : a_st) * b_bytes);
            bufs = (cap - fixed - used) / kABytes;
            if (bufs > kMaxEpiBufs) bufs = kMaxEpiBufs;
        }
        if (bufs < 1) {
            delete op;
            return set_error("igemm: not enough shared memory for an epilogue buffer");
        }
        if (ga > bufs - 1) ga = bufs > 1 ? bufs - 1 : 1;
        // A staging buffer belongs to one epilogue group (buffer s -> group s % ga, see the kernel): a buffer count that is a
        // multiple of ga keeps the groups evenly loaded.  Layers with a TMA-prefetched residual use the buffers as prefetch depth:
        // round UP if one pipeline stage pays for it (64 -> 256 residual conv: 4 stages + 7 buffers -> 3 + 8, same 6.1 TB/s),
        // else keep the count.  Without a residual extra buffers buy nothing: round DOWN and give the bytes back to the A ring
        // (1x1 256 -> 64: 5 buffers + 6 stages -> 4 + 7; rounding up instead cost that layer 16 %).
        if (ok && ga > 1 && bufs > ga && (bufs % ga) != 0 && !fat && !getenv("B200_EPI_UNEVEN")) {
            const int down = (bufs / ga) * ga, up = down + ga;
            const int stage_bytes = a_bytes + (p.b_resident ? 0 : b_bytes);
            const int min_st = (p.k_blocks >= 4) ? 3 : 2;
            const int give = ((up - bufs) * kABytes + stage_bytes - 1) / stage_bytes;
            if (p.res_tma) {
                // (rounding a residual layer DOWN costs prefetch depth: 128 -> 512 residual conv 0.106 -> 0.124 ms with 4
                // instead of 7 buffers; the uneven split of 7 over 4 groups costs nothing measurable)
                if (up <= kMaxEpiBufs && give <= 1 && a_st - give >= min_st) {
                    a_st -= give;
                    bufs = up;
                }
            } else {
                const int extra = ((bufs - down) * kABytes) / stage_bytes;
                bufs = down;
                a_st += extra;
                if (a_st > max_stages) a_st = max_stages;
            }
            if (!p.b_resident) b_st = a_st;
        }
        if (const char* e = getenv("B200_EPI_BUFS")) { if (atoi(e) >= 2 && atoi(e) <= bufs) { bufs = atoi(e); if (ga > bufs - 1) ga = bufs - 1; } }
    }
    if (!ok) {
        delete op;
        return set_error("igemm: not enough shared memory for two pipeline stages");
    }
    p.a_stages = a_st;
    p.b_stages = b_st;
    p.mma_group = (share && !pair && block_n <= 64 && a_st >= 5 && (p.b_resident || b_st >= 5) && !getenv("B200_NO_MMA_GROUP")) ? 1 : 0;
    // share-mode layers with N <= 64 keep four accumulators in flight (N x 4 <= 256 TMEM columns); the stem (one stage = one
    // tile) issues two tiles per group
    if (fat && !pair && block_n <= 64 && !(d->flags & (B200_EPI_HEAD_DOT | B200_EPI_CHAIN)) && !getenv("B200_NACC2")) {
        p.nacc = 4;
        if (stem_share && p.b_resident && a_st >= 4 && !getenv("B200_NO_MMA_GROUP")) p.mma_group = 2;
    }
    p.b_kb = b_kb;
    p.epi_bufs = bufs;
    p.epi_active = ga;
    op->smem_bytes = fixed + a_st * a_bytes + (p.b_resident ? 0 : (fat ? b_st * 3 : a_st) * b_bytes) + bufs * kABytes;

    op->threads = (base_warps(op->gather) + (op->gather ? kGatherWarps : 0)) * 32;
    if (d->flags & B200_EPI_HEAD_DOT) {  // constants of the head-tail epilogue: one small synchronous read-back at plan time
        op->hc = new (std::nothrow) HeadConsts;
        cudaError_t ce = op->hc ? cudaSuccess : cudaErrorMemoryAllocation;
        if (ce == cudaSuccess) ce = cudaMemcpy(op->hc->w2, d->offsets, 256 * sizeof(float), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(&op->hc->bias, reinterpret_cast<const float*>(d->offsets) + 256, sizeof(float), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(op->hc->scale, d->scale, 256 * sizeof(float), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(op->hc->shift, d->shift, 256 * sizeof(float), cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) {
            delete op->hc;
            delete op;
            return set_error("igemm: reading the head-tail constants failed: %s", cudaGetErrorString(ce));
        }
        for (int i = 0; i < 256; ++i) {
            if (op->hc->scale[i] != 1.0f) {
                delete op->hc;
                delete op;
                return set_error("igemm: HEAD_DOT applies the shift inside the UMMA chain: fold the BN scale into the weights and "
                                 "pass scale = 1 (scale[%d] = %g)", i, (double)op->hc->scale[i]);
            }
        }
    }
    if (getenv("B200_VERBOSE"))
        fprintf(stderr, "igemm plan: mode %d k %d s %d N %d kb %d tile %dx%d share %d resident %d pair %d a_stages %d (x %d B) b_stages %d bufs %d ga %d smem %d\n",
                p.mode, p.ksize, p.stride, p.block_n, p.k_blocks, p.tile_w, p.tile_h, p.a_kb, p.b_resident, op->pair, p.a_stages,
                p.a_bytes, p.b_stages, p.epi_bufs, p.epi_active, op->smem_bytes);

    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        delete op;
        return set_error("igemm: cannot query device: %s", cudaGetErrorString(cudaGetLastError()));
    }
    op->grid = p.num_tiles < sms ? p.num_tiles : sms;
    if (op->pair) {
        const int clusters = p.num_units < sms / 2 ? p.num_units : sms / 2;
        op->grid = 2 * clusters;
    }

    // ---- tensor maps
    int err = 0;
    const int ktot = p.k_blocks * kBlockK;
    {
        const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->n_pad};
        const uint64_t str[1] = {(uint64_t)ktot * 2};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(block_n > 256 ? 256 : p.b_rows)};
        err |= encode_map(&op->tm.b, d->w, 2, dims, str, box, "weights");
    }
    if (!err && !op->gather && d->mode != B200_IGEMM_STEM) {
        const uint64_t C = d->in_c, W = d->in_w, H = d->in_h, LD = x_ld;
        if (d->stride == 1) {
            err |= encode_nhwc(&op->tm.a[0], d->x, C, W, H, d->batch, LD, 1, 1, W, H * W, p.tile_w,
                               share ? p.tile_h + 2 : p.tile_h, "activations");
            if (!err && d->x2) {  // second K segment: x2 sampled at (stride2 * oy, stride2 * ox)
                const uint64_t s2 = d->stride2, LD2 = d->x2_ld ? d->x2_ld : d->in_c2;
                err |= encode_nhwc(&op->tm.a[1], d->x2, d->in_c2, W, H, d->batch, LD2, s2, s2, W * s2, H * s2 * W * s2, p.tile_w,
                                   p.tile_h, "second K segment");
            }
        } else if (stride_w == 1) {
            // stride (2, 1): two row-parity views, full width
            for (int v = 0; v < 2 && !err; ++v) {
                const char* base = reinterpret_cast<const char*>(d->x) + (uint64_t)v * W * LD * 2;
                err |= encode_nhwc(&op->tm.a[v], base, C, W, H / 2, d->batch, LD, 1, 2, W, H * W, p.tile_w, p.tile_h,
                                   "row-strided activations");
            }
        } else {
            // stride 2: four (row parity, column parity) strided views of the NHWC tensor
            const int nviews = (d->ksize == 3) ? 4 : 1;
            for (int v = 0; v < nviews && !err; ++v) {
                const int py = v >> 1, px = v & 1;
                const char* base = reinterpret_cast<const char*>(d->x) + ((uint64_t)py * W + px) * LD * 2;
                err |= encode_nhwc(&op->tm.a[v], base, C, W / 2, H / 2, d->batch, LD, 2, 2, W, H * W, p.tile_w, p.tile_h,
                                   "strided activations");
            }
        }
    }
    if (!err && p.epi_tma) {
        const uint64_t OW = p.out_w, OH = p.out_h, LD = d->out_ld;
        if (d->flags & B200_EPI_SHUFFLE2) {
            for (int v = 0; v < 4 && !err; ++v) {
                const int dy = v >> 1, dx = v & 1;
                char* base = reinterpret_cast<char*>(d->y) + ((uint64_t)dy * 2 * OW + dx) * LD * 2;
                err |= encode_nhwc(&op->tm.out[v], base, d->shuf_c, OW, OH, d->batch, LD, 2, 2, 2 * OW, 4 * OH * OW,
                                   p.tile_w, p.tile_h, "pixel-shuffle output");
            }
        } else {
            err |= encode_nhwc(&op->tm.out[0], d->y, d->out_c, OW, OH, d->batch, LD, 1, 1, OW, OH * OW, p.tile_w,
                               p.tile_h, "output");
        }
    }
    if (!err && (d->flags & B200_EPI_CHAIN)) {
        err |= encode_nhwc(&op->tm.out[1], d->chain_y, d->chain_n, p.out_w, p.out_h, d->batch, d->chain_ld, 1, 1, p.out_w,
                           (uint64_t)p.out_h * p.out_w, p.tile_w, p.tile_h, "chained output");
        const uint64_t dims[2] = {(uint64_t)d->n_pad, (uint64_t)d->chain_n};
        const uint64_t str[1] = {(uint64_t)d->n_pad * 2};
        const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)d->chain_n};
        if (!err) err |= encode_map(&op->tm.b2, d->chain_w, 2, dims, str, box, "chained weights");
    }
    if (!err && p.res_tma) {
        if (d->flags & B200_EPI_RES_UP2)
            err |= encode_nhwc(&op->tm.res, d->residual, d->out_c, p.out_w / 2, p.out_h / 2, d->batch, d->res_ld, 1, 1,
                               p.out_w / 2, (uint64_t)(p.out_h / 2) * (p.out_w / 2), p.tile_w / 2, p.tile_h / 2,
                               "half-resolution residual");
        else
            err |= encode_nhwc(&op->tm.res, d->residual, d->out_c, p.out_w, p.out_h, d->batch, d->res_ld, 1, 1, p.out_w,
                               (uint64_t)p.out_h * p.out_w, p.tile_w, p.tile_h, "residual");
    }
    if (err) {
        delete op;
        return 1;
    }
    // the attribute is per kernel function, not per op: always raise it to the full 227 KB opt-in
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)  // let the gather kernel's L1 have what its launches do not ask for as shared memory
        e = cudaFuncSetAttribute(igemm_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(igemm_kernel<2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) {
        delete op;
        return set_error("igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    }
    *out = op;
    return 0;
}

extern "C" int b200_igemm_launch(const b200_igemm* op, void* stream) {
    if (!op) return set_error("b200_igemm_launch: null op");
    if (op->d2) return dcn2_launch(op->d2, stream);
    // Programmatic dependent launch (also inside stream capture, where it becomes a programmatic graph edge): the
    // kernel's prologue overlaps the tail of the previous launch of the stream; see pdl_wait() in the kernel.
    static const bool pdl = getenv("B200_PDL") != nullptr;  // measured (round 2): 1598 vs 1612 img/s without -- off by default
    cudaLaunchAttribute attr[2];
    int nattr = 0;
    if (pdl) {
        attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    if (op->pair) {
        attr[nattr].id = cudaLaunchAttributeClusterDimension;
        attr[nattr].val.clusterDim.x = 2;
        attr[nattr].val.clusterDim.y = 1;
        attr[nattr].val.clusterDim.z = 1;
        ++nattr;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(op->grid);
    cfg.blockDim = dim3(op->threads);
    cfg.dynamicSmemBytes = op->smem_bytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    cudaError_t e;
    if (op->pair && op->p.a_kb > 1)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<2, 0, true>, op->tm, op->p, NoConsts{});
    else if (op->pair && (op->p.flags & B200_EPI_RES_UP2))
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 1, true>, op->tm, op->p, NoConsts{});
    else if (op->pair)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 0, true>, op->tm, op->p, NoConsts{});
    else if (op->p.flags & B200_EPI_CHAIN)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 4>, op->tm, op->p, NoConsts{});
    else if (op->gather)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<1>, op->tm, op->p, NoConsts{});
    else if ((op->p.flags & B200_EPI_GELU) && op->p.a_kb > 1)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<2, 3>, op->tm, op->p, NoConsts{});
    else if (op->p.flags & B200_EPI_GELU)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 3>, op->tm, op->p, NoConsts{});
    else if (op->p.a_kb > 1)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<2>, op->tm, op->p, NoConsts{});
    else if (op->p.flags & B200_EPI_HEAD_DOT)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 2>, op->tm, op->p, *op->hc);
    else if (op->p.flags & B200_EPI_RES_UP2)
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0, 1>, op->tm, op->p, NoConsts{});
    else
        e = cudaLaunchKernelEx(&cfg, igemm_kernel<0>, op->tm, op->p, NoConsts{});
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error("igemm launch failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

// development aid (not declared in the public header): read and clear the B200_DEBUG & 16 counters
extern "C" int b200_igemm_debug_counters(unsigned long long* out) {
    unsigned long long zero[16] = {0};
    if (cudaMemcpyFromSymbol(out, b200::g_prof, sizeof(zero)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(b200::g_prof, zero, sizeof(zero)) != cudaSuccess;
}

extern "C" int b200_igemm_destroy(b200_igemm* op) {
    if (op && op->d2) dcn2_destroy(op->d2);
    if (op) delete op->hc;
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

extern "C" int b200_dcn_coords(const float* raw, void* coords, int batch, int in_h, int in_w, int x_ld, int stride,
                               void* stream) {
    if (!raw || !coords) return set_error("b200_dcn_coords: null pointer");
    if (batch <= 0 || in_h <= 0 || in_w <= 0 || (stride != 1 && stride != 2) || x_ld <= 0)
        return set_error("b200_dcn_coords: bad shape");
    const int out_h = in_h / stride, out_w = in_w / stride;
    const int64_t total = static_cast<int64_t>(batch) * out_h * out_w;
    int64_t blocks = (total + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dcn_coords_kernel<<<static_cast<int>(blocks), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        raw, reinterpret_cast<uint4*>(coords), batch, out_h, out_w, in_h, in_w, x_ld, stride);
    return check_cuda(cudaGetLastError(), "b200_dcn_coords launch");
}
