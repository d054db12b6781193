// Modulated deformable conv v2 (torchvision.ops.deform_conv2d, model/det/dcn.py:48-57) for layers whose whole output
// width fits one UMMA (Cout = 128 or 256): the sampled A operand goes to TENSOR MEMORY instead of shared memory.
//
// Why: the first DCN kernel (igemm_kernel<1>) is bound by the L1 data pipe -- one 128-byte wavefront
// per clock shared by the LSU and the tensor core's shared-memory operand reads.  Per k-block (128 pixels x 64 channels
// x 1 tap) it carried 512 wavefronts of corner gathers + 128 of record loads + 128 of A stores (st.shared) + 128 of A
// reads (tcgen05.mma) + N of weight reads.  Here
//   * the sampling records of a tile are brought to shared memory by ONE TMA box per tile (zero fill outside the image
//     = weight 0 at a safe address, so the producers carry no validity mask) and read with conflict-free ld.shared:
//     32 wavefronts per k-block instead of 128;
//   * the blended rows go registers -> TMEM with tcgen05.st (16x256b: a row is spread over 4 lanes, so the 8-lanes-per-row
//     gather layout needs only ONE shuffle exchange per pair of rows, 64 wavefront slots per k-block, instead of the
//     128 + 128 of the st.shared / operand-read round trip) and tcgen05.mma reads A from TMEM ([a_tmem] operand form);
//   * K inside a 64-channel chunk is permuted to the order the 16x256b store lays down (ops.pack_dcn_weight applies the
//     same permutation to the weights), so no register transpose is needed.
// Pipe budget per k-block: 1152 -> 736 wavefront slots at N = 256 against 512 cycles of MMA.
//
// Warp roles (24 warps): 0 = TMA loader (weight k-blocks on a ring, one record box per tile, double buffered),
// 1 = MMA issuer, 2 = epilogue DMA (TMA stores), 3 = idle, 4-7 = epilogue math (TMEM -> BN + ReLU -> bf16 -> swizzled
// staging), 8-23 = producers: warp w owns TMEM lane quarter w % 4; the four warps of producer group g = (w - 8) / 4 fill
// the A stages of k-blocks g, g + 4, g + 8, ... (global k-block counter across tiles), 32 rows x 64 channels each.
//
// TMEM columns: [0, nacc * N) accumulators (N = 128: two, so tile i+1's MMAs overlap tile i's epilogue; N = 256: one),
// [256, 512) the A ring, 8 stages x 32 columns (64 bf16 per row).
//
// TMEM row <-> pixel: unit u (0..7) of a warp = rows 4u + sub (sub = lane / 8) of its 32-pixel slice, eight lanes x 16
// bytes per row.  Units 2p and 2p + 1 form a pair: after blending both, lanes l8 < 4 keep the rows of unit 2p and take
// the upper four 16-byte pieces from lanes l8 + 4, lanes l8 >= 4 keep unit 2p + 1 and take the lower pieces from l8 - 4.
// Lane T = 8 sub + l8 then owns TMEM lane T / 4 (+ 8 for the second pair of a 16-row half), which is what
// tcgen05.st.16x256b wants; the epilogue maps TMEM lane rho back to the pixel with the inverse bit shuffle.
#include "common.cuh"
#include "internal.h"
#include "dcn2.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace b200 {

namespace {

constexpr int kProdWarp0 = 8;
constexpr int kProdWarps = 16;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;
constexpr int kAStages = 8;
constexpr int kACol0 = 256;                 // first TMEM column of the A ring
constexpr int kRecBytes = 128 * 9 * 16;     // sampling records of one tile
constexpr int kChunkBytes = 128 * 128;      // one 128-pixel x 64-channel bf16 chunk (epilogue staging buffer)
constexpr int kMaxBStages = 6;
constexpr int kMaxEpiBufs = 4;

struct Smem {
    uint8_t* b;
    uint8_t* epi;
    uint8_t* rec;
    float* ss;
    uint64_t* a_full;
    uint64_t* a_empty;
    uint64_t* b_full;
    uint64_t* b_empty;
    uint64_t* tfull;
    uint64_t* tempty;
    uint64_t* rec_full;
    uint64_t* rec_empty;
    uint64_t* out_ready;
    uint64_t* slot_free;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* raw, const Dcn2Params& p) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    Smem s;
    s.b = base;
    s.epi = s.b + p.b_stages * p.block_n * 128;
    s.rec = s.epi + p.epi_bufs * kChunkBytes;
    s.ss = reinterpret_cast<float*>(s.rec + 2 * kRecBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.ss + 2 * p.block_n);
    s.a_full = bars;
    s.a_empty = s.a_full + kAStages;
    s.b_full = s.a_empty + kAStages;
    s.b_empty = s.b_full + kMaxBStages;
    s.tfull = s.b_empty + kMaxBStages;
    s.tempty = s.tfull + 2;
    s.rec_full = s.tempty + 2;
    s.rec_empty = s.rec_full + 2;
    s.out_ready = s.rec_empty + 2;
    s.slot_free = s.out_ready + kMaxEpiBufs;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.slot_free + kMaxEpiBufs);
    return s;
}

struct Tile {
    int n, y0, x0;
};
__device__ __forceinline__ Tile decode(const Dcn2Params& p, int tile) {
    Tile t;
    const int per_img = p.tiles_x * p.tiles_y;
    t.n = tile / per_img;
    const int m = tile - t.n * per_img;
    const int ty = m / p.tiles_x;
    t.y0 = ty * p.tile_h;
    t.x0 = (m - ty * p.tiles_x) * p.tile_w;
    return t;
}

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint4& lo0, const uint4& lo1, const uint4& hi0,
                                                   const uint4& hi1) {
    // register 4j + 2v + r -> (TMEM lane T/4 + 8v, column 8j + 2(T%4) + r); j = 0,1: lower 16-byte piece, j = 2,3: upper
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(lo0.x), "r"(lo0.y), "r"(lo1.x), "r"(lo1.y), "r"(lo0.z), "r"(lo0.w), "r"(lo1.z), "r"(lo1.w),
        "r"(hi0.x), "r"(hi0.y), "r"(hi1.x), "r"(hi1.y), "r"(hi0.z), "r"(hi0.w), "r"(hi1.z), "r"(hi1.w)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_plain(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    tma_load_4d(smem, m, bar, c0, c1, c2, c3);
}

// position of one producer group in its stream of k-blocks (g, g + 4, g + 8, ... of the CTA's global k-block sequence)
struct Cursor {
    int it;      // tile iteration of this CTA
    int chunk;   // 64-channel chunk
    int tap;     // 0..8
    uint32_t rec_s;               // shared address of this thread's unit-0 record of (tile, tap)
    const __nv_bfloat16* src;     // image base + chunk * 64 + l8 * 8
};

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) dcn2_kernel(const __grid_constant__ Dcn2Tmaps tm, const Dcn2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const Smem sm = carve(smem_raw, p);
    const int b_bytes = p.block_n * 128;
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm.b);
        tma_prefetch_desc(&tm.rec);
        tma_prefetch_desc(&tm.out);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kAStages; ++s) {
                mbar_init(&sm.a_full[s], 4);
                mbar_init(&sm.a_empty[s], 1);
            }
            for (int s = 0; s < kMaxBStages; ++s) {
                mbar_init(&sm.b_full[s], 1);
                mbar_init(&sm.b_empty[s], 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&sm.tfull[a], 1);
                mbar_init(&sm.tempty[a], 4);
                mbar_init(&sm.rec_full[a], 1);
                mbar_init(&sm.rec_empty[a], kProdWarps);
            }
            for (int a = 0; a < kMaxEpiBufs; ++a) {
                mbar_init(&sm.out_ready[a], 4);
                mbar_init(&sm.slot_free[a], 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(sm.tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) {
        sm.ss[i] = __ldg(p.scale + i);
        sm.ss[p.block_n + i] = __ldg(p.shift + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA loader: records + weights
        int s = 0;
        uint32_t ph = 0;
        auto load_rec = [&](int it) {  // record box of tile iteration `it` into buffer it & 1
            const int par = it & 1;
            mbar_wait(&sm.rec_empty[par], ((it >> 1) & 1) ^ 1);
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x));
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.rec_full[par], kRecBytes);
                tma_load_4d_plain(sm.rec + par * kRecBytes, &tm.rec, &sm.rec_full[par], 0, t.x0, t.y0, t.n);
            }
            __syncwarp();
        };
        if (my_tiles > 0) load_rec(0);
        for (int it = 0; it < my_tiles; ++it) {
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                // The next tile's records are requested a few weight k-blocks into this tile: by then every producer warp
                // has left the previous tile (its buffer is free, the wait does not hold the weight ring up), and the
                // producers, at most 12 k-blocks ahead of the MMAs, are still far from needing them.
                if (kb == p.rec_lead && it + 1 < my_tiles) load_rec(it + 1);
                mbar_wait(&sm.b_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&sm.b_full[s], static_cast<uint32_t>(b_bytes));
                    tma_load_2d(sm.b + s * b_bytes, &tm.b, &sm.b_full[s], kb * 64, 0);
                }
                __syncwarp();
                if (++s == p.b_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (converged warp, one lane issues)
        const uint32_t idesc = umma_idesc_bf16(static_cast<uint32_t>(p.block_n));
        const uint32_t b0 = smem_u32(sm.b);
        int sb = 0;
        uint32_t bph = 0;
        uint32_t q = 0;  // global k-block counter: A stage q & 7, phase (q >> 3) & 1
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = (p.nacc == 2) ? (it & 1) : 0;
            const uint32_t tph = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&sm.tempty[acc], tph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.block_n);
            for (int kb = 0; kb < p.k_blocks; ++kb, ++q) {
                const uint32_t sa = q & (kAStages - 1);
                mbar_wait(&sm.a_full[sa], (q >> 3) & 1);
                mbar_wait(&sm.b_full[sb], bph);
                tc_fence_after();
                const uint32_t a_tmem = tmem_base + kACol0 + sa * 32;
                const uint64_t db = umma_desc_sw128(b0 + sb * b_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 K elements = 8 TMEM columns of A, 32 bytes of a swizzled weight row
                        umma_bf16_ts(d_tmem, a_tmem + 8 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&sm.a_empty[sa]);
                    umma_commit(&sm.b_empty[sb]);
                }
                __syncwarp();
                if (++sb == p.b_stages) { sb = 0; bph ^= 1; }
            }
            if (elect_one()) umma_commit(&sm.tfull[acc]);
            __syncwarp();
        }
    } else if (warp == 2) {
        // ===================================================================== epilogue DMA: TMA stores of finished chunks
        const int B = p.epi_bufs, nchunks = p.out_c >> 6;
        int g = 0, slot = 0, rslot = 0;
        uint32_t par = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x));
            for (int c = 0; c < nchunks; ++c, ++g) {
                mbar_wait(&sm.out_ready[slot], par);
                if (elect_one()) {
                    tma_store_4d(&tm.out, sm.epi + slot * kChunkBytes, c * 64, t.x0, t.y0, t.n);
                    bulk_commit();
                    if (g >= 1) {  // the store of chunk g-1 has read its buffer: recycle it
                        bulk_wait_read1();
                        mbar_arrive(&sm.slot_free[rslot]);
                    }
                }
                __syncwarp();
                if (g >= 1 && ++rslot == B) rslot = 0;
                if (++slot == B) { slot = 0; par ^= 1; }
            }
        }
        if (elect_one()) bulk_wait_all();
        __syncwarp();
    } else if (warp >= 4 && warp < kProdWarp0) {
        // ===================================================================== epilogue math (one group of 4 warps)
        const int quarter = warp & 3;
        // TMEM lane rho = 16h + 8v + 2 sub + ab  ->  unit u = 4h + 2v + ab, pixel 4u + sub of the quarter's 32
        const int u = ((lane >> 4) << 2) | (((lane >> 3) & 1) << 1) | (lane & 1);
        const int row = quarter * 32 + 4 * u + ((lane >> 1) & 3);
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const int B = p.epi_bufs, nchunks = p.out_c >> 6;
        const uint32_t ss_s = smem_u32(sm.ss), ss_shift = static_cast<uint32_t>(p.block_n) * 4;
        int slot = 0;
        uint32_t par = 0;
        int g = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = (p.nacc == 2) ? (it & 1) : 0;
            const uint32_t aph = (p.nacc == 2) ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&sm.tfull[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * p.block_n);
            for (int c = 0; c < nchunks; ++c, ++g) {
                const uint32_t mine = smem_u32(sm.epi) + static_cast<uint32_t>(slot) * kChunkBytes + static_cast<uint32_t>(row) * 128;
                if (g >= B) mbar_wait(&sm.slot_free[slot], par ^ 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 64 + h * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int piece = h * 4 + j;
                        const uint32_t sca = ss_s + static_cast<uint32_t>(c * 64 + piece * 8) * 4;
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
                        if (p.flags & B200_EPI_RELU) {
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
                if (c == nchunks - 1) {  // accumulator fully read: hand it back before the staging hand-shake
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.tempty[as]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.out_ready[slot]);
                if (++slot == B) { slot = 0; par ^= 1; }
            }
        }
    } else if (warp >= kProdWarp0) {
        // ===================================================================== producers
        const int pw = warp - kProdWarp0;
        const int qd = warp & 3;   // TMEM lane quarter this warp may touch (hardware rule: warp index % 4)
        const int grp = pw >> 2;   // k-blocks grp, grp + 4, ... of the CTA's sequence
        const int sub = lane >> 3, l8 = lane & 7;
        const bool lowlane = l8 < 4;
        const int cs = (p.in_w > 1) ? p.x_ld : 0;            // element step to the block's right column
        const int rs = (p.in_h > 1) ? p.in_w * p.x_ld : 0;   // element step to the block's bottom row
        const int64_t img_elems = static_cast<int64_t>(p.in_h) * p.in_w * p.x_ld;
        const uint32_t rec_row = smem_u32(sm.rec) + static_cast<uint32_t>((32 * qd + sub) * 144);
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(32 * qd) << 16) + kACol0;
        const int total = my_tiles * p.k_blocks;

        auto cursor_at = [&](int q) {
            Cursor c;
            c.it = q / p.k_blocks;
            const int kb = q - c.it * p.k_blocks;
            c.chunk = kb / 9;
            c.tap = kb - c.chunk * 9;
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + c.it * static_cast<int>(gridDim.x));
            c.rec_s = rec_row + static_cast<uint32_t>((c.it & 1) * kRecBytes + c.tap * 16);
            c.src = p.x + t.n * img_elems + c.chunk * 64 + l8 * 8;
            return c;
        };
        auto advance4 = [&](Cursor& c) {  // four k-blocks further; returns true when that enters a new tile
            c.tap += 4;
            c.rec_s += 64;
            if (c.tap < 9) return false;
            c.tap -= 9;
            c.rec_s -= 144;
            c.src += 64;
            if (++c.chunk < p.cchunks) return false;
            c.chunk = 0;
            ++c.it;
            const Tile t = decode(p, static_cast<int>(blockIdx.x) + c.it * static_cast<int>(gridDim.x));
            c.rec_s = rec_row + static_cast<uint32_t>((c.it & 1) * kRecBytes + c.tap * 16);
            c.src = p.x + t.n * img_elems + l8 * 8;
            return true;
        };

        if (grp < total) {
            Cursor cur = cursor_at(grp);
            mbar_wait(&sm.rec_full[cur.it & 1], (cur.it >> 1) & 1);
            // pipeline state at the top of an iteration: rcd[0], rcd[1] = records of units 0, 1; dat[0] = corners of unit 0
            uint4 rcd[4];
            uint4 dat[2][4];
            rcd[0] = lds128(cur.rec_s);
            rcd[1] = lds128(cur.rec_s + 576);
            {
                const __nv_bfloat16* b0 = cur.src + static_cast<int>(rcd[0].x);
                dat[0][0] = ldg128(b0);
                dat[0][1] = ldg128(b0 + cs);
                dat[0][2] = ldg128(b0 + rs);
                dat[0][3] = ldg128(b0 + rs + cs);
            }
            int i = 0;  // iteration of this group: k-block q = grp + 4 i, A stage grp + 4 (i & 1), use (i >> 1)
            for (int q = grp; q < total; q += 4, ++i) {
                Cursor nxt = cur;
                const bool has_next = q + 4 < total;
                bool new_tile = false;
                if (has_next) {
                    new_tile = advance4(nxt);
                    if (new_tile) mbar_wait(&sm.rec_full[nxt.it & 1], (nxt.it >> 1) & 1);
                }
                const int stage = grp + 4 * (i & 1);
                uint4 res_a = make_uint4(0, 0, 0, 0);
                uint4 lo[2], hi[2];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // record of unit j + 2
                    if (j + 2 < 8) {
                        rcd[(j + 2) & 3] = lds128(cur.rec_s + (j + 2) * 576);
                    } else if (has_next) {
                        rcd[(j + 2) & 3] = lds128(nxt.rec_s + (j + 2 - 8) * 576);
                    }
                    // corners of unit j + 1
                    if (j + 1 < 8 || has_next) {
                        const __nv_bfloat16* b0 = ((j + 1 < 8) ? cur.src : nxt.src) + static_cast<int>(rcd[(j + 1) & 3].x);
                        dat[(j + 1) & 1][0] = ldg128(b0);
                        dat[(j + 1) & 1][1] = ldg128(b0 + cs);
                        dat[(j + 1) & 1][2] = ldg128(b0 + rs);
                        dat[(j + 1) & 1][3] = ldg128(b0 + rs + cs);
                    }
                    // blend unit j: packed bf16x2 FMAs, one bf16 rounding per FMA (same arithmetic as igemm_kernel<1>)
                    const uint32_t w01 = rcd[j & 3].y, w23 = rcd[j & 3].z;
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
                    if ((j & 1) == 0) {
                        res_a = res;
                    } else {
                        // pair exchange: 8 lanes x 16 B per row -> 4 lanes x 32 B per row (what tcgen05.st.16x256b stores)
                        uint4 snd, rcv;
                        snd.x = lowlane ? res.x : res_a.x;
                        snd.y = lowlane ? res.y : res_a.y;
                        snd.z = lowlane ? res.z : res_a.z;
                        snd.w = lowlane ? res.w : res_a.w;
                        rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 4);
                        rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 4);
                        rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 4);
                        rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 4);
                        const int v = (j >> 1) & 1;
                        lo[v].x = lowlane ? res_a.x : rcv.x;
                        lo[v].y = lowlane ? res_a.y : rcv.y;
                        lo[v].z = lowlane ? res_a.z : rcv.z;
                        lo[v].w = lowlane ? res_a.w : rcv.w;
                        hi[v].x = lowlane ? rcv.x : res.x;
                        hi[v].y = lowlane ? rcv.y : res.y;
                        hi[v].z = lowlane ? rcv.z : res.z;
                        hi[v].w = lowlane ? rcv.w : res.w;
                    }
                    if (j == 3 || j == 7) {
                        if (j == 3) {
                            mbar_wait(&sm.a_empty[stage], ((i >> 1) & 1) ^ 1);
                            tc_fence_after();
                        }
                        tmem_st_16x256b_x4(t_lane + (static_cast<uint32_t>(16 * (j >> 2)) << 16) + static_cast<uint32_t>(stage * 32),
                                           lo[0], lo[1], hi[0], hi[1]);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sm.a_full[stage]);
                    if (!has_next || new_tile) mbar_arrive(&sm.rec_empty[cur.it & 1]);  // done with this tile's records
                }
                cur = nxt;
            }
        }
        // groups that never entered a tile's records still owe rec_empty arrivals?  No: every group visits every tile
        // (k_blocks >= 18 > 4), so each warp arrives exactly once per tile.
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------ host side
static int pick_tile_w2(int out_h, int out_w, int max_tw) {
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

int dcn2_supported(const b200_igemm_desc* d) {
    return d->ksize == 3 && (d->n_pad == 128 || d->n_pad == 256) && d->out_c == d->n_pad && (d->in_c % 64) == 0 &&
           (d->flags & ~B200_EPI_RELU) == 0 && (d->block_n == 0 || d->block_n == d->n_pad);
}

int dcn2_create(const b200_igemm_desc* d, Dcn2Plan** out) {
    *out = nullptr;
    if (!dcn2_supported(d))
        return set_error("igemm: DCN_TMEM needs a 3x3 layer with out_c == n_pad in {128, 256}, in_c %% 64 == 0 and no epilogue "
                         "flag besides RELU");
    if (!d->offsets) return set_error("igemm: DCN needs offsets");
    Dcn2Plan* pl = new (std::nothrow) Dcn2Plan;
    if (!pl) return set_error("igemm: out of host memory");
    memset(pl, 0, sizeof(*pl));
    Dcn2Params& p = pl->p;
    p.batch = d->batch;
    p.in_h = d->in_h;
    p.in_w = d->in_w;
    p.in_c = d->in_c;
    p.stride = d->stride;
    p.out_h = d->in_h / d->stride;
    p.out_w = d->in_w / d->stride;
    p.out_c = d->out_c;
    p.flags = d->flags;
    p.block_n = d->n_pad;
    p.x_ld = d->x_ld ? d->x_ld : d->in_c;
    p.nacc = (p.block_n <= 128) ? 2 : 1;
    int max_tw = 16;  // 16 x 8 tiles: the squarest = the smallest gather neighbourhood (measured: 32 x 4 is 3-6 % slower)
    if (const char* e = getenv("B200_CONV_TW")) { if (atoi(e) >= 8 && atoi(e) <= 128) max_tw = atoi(e); }
    p.tile_w = pick_tile_w2(p.out_h, p.out_w, max_tw);
    p.tile_h = 128 / p.tile_w;
    p.tiles_x = (p.out_w + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (p.out_h + p.tile_h - 1) / p.tile_h;
    p.num_tiles = p.batch * p.tiles_x * p.tiles_y;
    p.cchunks = d->in_c / 64;
    p.k_blocks = 9 * p.cchunks;
    p.x = reinterpret_cast<const __nv_bfloat16*>(d->x);
    p.scale = d->scale;
    p.shift = d->shift;
    // shared memory: weight ring + epilogue staging + two record buffers; what the launch does not take stays L1, which the
    // corner gathers live on
    // (measured, batch 16: two 32 KB weight stages feed N = 256 as well as three; four shrink L1 and cost 10 %)
    p.b_stages = (p.block_n == 256) ? 2 : 3;
    p.epi_bufs = 2;
    if (const char* e = getenv("B200_DCN2_BSTAGES")) { if (atoi(e) >= 2 && atoi(e) <= kMaxBStages) p.b_stages = atoi(e); }
    if (const char* e = getenv("B200_DCN2_EPIBUFS")) { if (atoi(e) >= 2 && atoi(e) <= kMaxEpiBufs) p.epi_bufs = atoi(e); }
    p.rec_lead = p.b_stages + 2;
    if (p.rec_lead > p.k_blocks - 1) p.rec_lead = p.k_blocks - 1;
    pl->smem_bytes = 1024 + p.b_stages * p.block_n * 128 + p.epi_bufs * kChunkBytes + 2 * kRecBytes + 8 * p.block_n + 512;
    if (pl->smem_bytes > 227 * 1024) {
        delete pl;
        return set_error("igemm: DCN_TMEM shared memory plan does not fit");
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        delete pl;
        return set_error("igemm: cannot query device: %s", cudaGetErrorString(cudaGetLastError()));
    }
    pl->grid = p.num_tiles < sms ? p.num_tiles : sms;
    int err = 0;
    {
        const uint64_t ktot = (uint64_t)p.k_blocks * 64;
        const uint64_t dims[2] = {ktot, (uint64_t)d->n_pad};
        const uint64_t str[1] = {ktot * 2};
        const uint32_t box[2] = {64u, (uint32_t)p.block_n};
        err |= encode_map_typed(&pl->tm.b, 0, 1, d->w, 2, dims, str, box, "DCN weights");
    }
    if (!err) {
        const uint64_t OW = p.out_w, OH = p.out_h, LD = d->out_ld;
        const uint64_t dims[4] = {(uint64_t)d->out_c, OW, OH, (uint64_t)d->batch};
        const uint64_t str[3] = {LD * 2, OW * LD * 2, OH * OW * LD * 2};
        const uint32_t box[4] = {64u, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1u};
        err |= encode_map_typed(&pl->tm.out, 0, 1, d->y, 4, dims, str, box, "DCN output");
    }
    if (!err) {
        const uint64_t OW = p.out_w, OH = p.out_h;
        const uint64_t dims[4] = {36, OW, OH, (uint64_t)d->batch};
        const uint64_t str[3] = {144, OW * 144, OH * OW * 144};
        const uint32_t box[4] = {36u, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1u};
        err |= encode_map_typed(&pl->tm.rec, 1, 0, d->offsets, 4, dims, str, box, "DCN sampling records");
    }
    if (err) {
        delete pl;
        return 1;
    }
    cudaError_t e = cudaFuncSetAttribute(dcn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
        delete pl;
        return set_error("igemm: cudaFuncSetAttribute(dcn2) failed: %s", cudaGetErrorString(e));
    }
    if (getenv("B200_VERBOSE"))
        fprintf(stderr, "dcn2 plan: N %d kb %d tile %dx%d b_stages %d epi_bufs %d nacc %d smem %d grid %d\n", p.block_n, p.k_blocks,
                p.tile_w, p.tile_h, p.b_stages, p.epi_bufs, p.nacc, pl->smem_bytes, pl->grid);
    *out = pl;
    return 0;
}

int dcn2_launch(const Dcn2Plan* pl, void* stream) {
    dcn2_kernel<<<pl->grid, kThreads, pl->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(pl->tm, pl->p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("dcn2 launch failed: %s", cudaGetErrorString(e));
    return 0;
}

void dcn2_destroy(Dcn2Plan* pl) { delete pl; }

}  // namespace b200
