// GPU replacement of DBPostProcessor.__call__ (src/det/test.py:55-106): threshold, 8-connected component
// labelling with a min-index union-find (row runs found with warp ballots, merged with atomicMin), canonical
// relabel (label = raster rank of the component's first pixel), per-component bounding box / pixel count /
// fixed-point probability sum (warp-segmented run reduction + atomics), score / area filters and the
// closed-form rectangle un-clip.  HBM-bound integer work; everything after the threshold is integer or
// uncontracted IEEE double so it is bit-exact against oracle/postproc_oracle.py.
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kChunkWords = 128;  // 32-pixel words per scan chunk (one block of 256 threads)

struct CompRec {  // per-component accumulator, 32 B
    int minx, miny, maxx, maxy;
    int npix, pad;
    unsigned long long qsum;
};

struct PP {
    const float* prob;
    uint32_t* mask;      // foreground bitmask, one 32-pixel word per (row, word) -- rows are padded to whole words
    uint32_t* rootmask;  // bit set where the pixel is the root (= first pixel in raster order) of its component
    int* L;              // union-find parent per FOREGROUND pixel (global pixel index); background entries are never touched
    int* R;              // canonical rank, valid at root pixels
    int* labels;         // output label map (may be null)
    int* chunk_cnt;
    int* ncomp;
    CompRec* rec;
    b200_box* boxes;
    int* counts;
    int batch, h, w, wpr /*words per row*/, nchunks, maxc, min_area;
    float thresh;
    double box_thresh_q /* box_thresh * 2^24 */, ratio;
};

__device__ __forceinline__ int uf_find(const int* L, int a) {
    int p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}
// find with path compression: re-point the start at the root it found (any ancestor is a valid parent, and parents
// only ever decrease, so a racing writer can at worst store another valid ancestor)
// Path halving on the way up: every visited node is re-pointed at its grandparent.  Only a ROOT's entry is ever the
// target of a union's atomicMin and a node that has a parent never becomes a root again (parents have smaller indices), so a
// plain store of an ancestor into a non-root's entry cannot lose a link.  (Without it a page-high component is a linked list
// of row heads when the rows are merged concurrently, and one find is a thousand dependent L2 round trips.)
__device__ __forceinline__ int uf_find_compress(int* L, int a) {
    const int start = a;
    int p = L[a];
    while (p != a) {
        const int gp = L[p];
        if (gp == p) { a = p; break; }
        L[a] = gp;
        a = gp;
        p = L[a];
    }
    if (a != start && L[start] > a) L[start] = a;
    return a;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find_compress(L, a);
        b = uf_find_compress(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }  // a > b: hang a under b
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;  // somebody re-parented a meanwhile: retry from there
    }
}

// One warp = one image row.  Writes the foreground bitmask and, for foreground pixels only, L[p] = index of the first pixel
// of p's horizontal run -- the WHOLE run, across word boundaries: the start of the run that enters a word from the left is
// found by a scan over the row's word masks (a word of 32 foreground pixels is transparent, any other word either ends the
// run or starts a new one behind its highest background pixel).  The merge kernel then only links runs of adjacent rows;
// linking the words of a run one by one built chains as long as the run (a page-wide component: 32 unions per row, each
// walking the chain of the ones before it).
__global__ void __launch_bounds__(256) pp_init_kernel(const PP p) {
    const int lane = threadIdx.x & 31;
    const uint32_t nrows = static_cast<uint32_t>(p.batch) * p.h;
    const uint32_t below = (1u << lane) - 1u;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += (gridDim.x * blockDim.x) >> 5) {
        const float* pr = p.prob + static_cast<size_t>(row) * p.w;
        const uint32_t g0 = row * p.w;
        int carry = -1;  // x of the first pixel of the run that holds the last pixel of the previous 32-word segment, or -1
        for (int w0 = 0; w0 < p.wpr; w0 += 32) {
            const int nw = min(32, p.wpr - w0);
            // masks of the segment's words: lane j keeps word w0 + j
            uint32_t mymask = 0;
            for (int j0 = 0; j0 < nw; j0 += 8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int x = (w0 + j0 + i) * 32 + lane;
                    v[i] = (j0 + i < nw && x < p.w) ? __ldg(pr + x) : -1.0f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int x = (w0 + j0 + i) * 32 + lane;
                    const uint32_t m = __ballot_sync(0xffffffffu, (j0 + i < nw) && (x < p.w) && (v[i] > p.thresh));
                    if (lane == j0 + i) mymask = m;
                }
            }
            if (lane < nw) p.mask[row * p.wpr + w0 + lane] = mymask;
            // out = x of the first pixel of the run that holds the LAST pixel of this lane's word, -1 if that pixel is background
            const bool full = (mymask == 0xffffffffu);
            int out = -1;
            if (!full && (mymask >> 31)) out = (w0 + lane) * 32 + (32 - __clz(~mymask));
            const uint32_t nonfull = __ballot_sync(0xffffffffu, !full);
            const uint32_t nb = nonfull & below;          // words of this segment before mine that are not all foreground
            const int r = nb ? (31 - __clz(nb)) : 0;      // the nearest of them
            const int o = __shfl_sync(0xffffffffu, out, r);
            if (full) out = nb ? ((o >= 0) ? o : (w0 + r + 1) * 32) : ((carry >= 0) ? carry : w0 * 32);
            int in = __shfl_up_sync(0xffffffffu, out, 1);  // run entering my word from the left (first pixel's x), or -1
            if (lane == 0) in = carry;
            for (int j = 0; j < nw; ++j) {
                const uint32_t m = __shfl_sync(0xffffffffu, mymask, j);
                const int inj = __shfl_sync(0xffffffffu, in, j);
                if ((m >> lane) & 1u) {
                    // background pixels strictly below this lane; the run starts right above the highest of them
                    const uint32_t z = ~m & below;
                    const int x0 = (w0 + j) * 32;
                    const int start = z ? (x0 + 32 - __clz(z)) : ((inj >= 0) ? inj : x0);
                    p.L[g0 + x0 + lane] = static_cast<int>(g0) + start;
                }
            }
            carry = __shfl_sync(0xffffffffu, out, nw - 1);
        }
    }
}

// Merge runs with the row above (8-connectivity; a horizontal run is one set since pp_init).  Which unions are needed is
// decided from six bitmask words per warp, skipping the ones a horizontally adjacent pixel is guaranteed to perform.
// One warp walks a 32-pixel column of kMergeRows rows DOWNWARDS: when row y is linked, row y-1 of the same strip already hangs
// under the strip's top, so a find is a hop or two.  (Rows linked concurrently each grab their still-unlinked upper
// neighbour and a tall component becomes a linked list of row heads; the first find over it was hundreds of dependent L2
// round trips -- a 170 us tail on a page-sized component.)  Lists now form only across strips (h / 32 entries).
constexpr int kMergeRows = 32;
__global__ void __launch_bounds__(256) pp_merge_kernel(const PP p) {
    const int lane = threadIdx.x & 31;
    const uint32_t bit = 1u << lane;
    const uint32_t strips = static_cast<uint32_t>((p.h + kMergeRows - 1) / kMergeRows);
    const uint32_t items = static_cast<uint32_t>(p.batch) * strips * p.wpr;
    for (uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t t = it / p.wpr;
        const int wx = static_cast<int>(it - t * p.wpr);
        const uint32_t n = t / strips;
        const int y0 = static_cast<int>(t - n * strips) * kMergeRows;
        const int y1 = min(y0 + kMergeRows, p.h);
        const bool has_l = wx > 0, has_r = wx + 1 < p.wpr;
        uint32_t wd = (n * p.h + y0) * p.wpr + wx;
        uint32_t g = (n * p.h + y0) * p.w + wx * 32 + lane;
        uint32_t u = 0, ul = 0, ur = 0;  // masks of the row above
        if (y0 > 0) {
            u = p.mask[wd - p.wpr];
            ul = has_l ? p.mask[wd - p.wpr - 1] : 0u;
            ur = has_r ? p.mask[wd - p.wpr + 1] : 0u;
        }
        for (int y = y0; y < y1; ++y, wd += p.wpr, g += p.w) {
            const uint32_t m = p.mask[wd];
            const uint32_t ml = has_l ? p.mask[wd - 1] : 0u, mr = has_r ? p.mask[wd + 1] : 0u;
            if (m) {
                const uint32_t left = (m << 1) | (ml >> 31);    // bit i: pixel i-1 of this row is foreground
                const uint32_t right = (m >> 1) | (mr << 31);   // bit i: pixel i+1
                const uint32_t upl = (u << 1) | (ul >> 31), upr = (u >> 1) | (ur << 31);
                const uint32_t do_up = m & u & ~(left & upl);   // the left pair has not linked these two runs yet
                const uint32_t do_ul = m & ~u & upl & ~left;
                const uint32_t do_ur = m & ~u & upr & ~right;
                if (do_up & bit) uf_union(p.L, static_cast<int>(g), static_cast<int>(g - p.w));
                if (do_ul & bit) uf_union(p.L, static_cast<int>(g), static_cast<int>(g - p.w - 1));
                if (do_ur & bit) uf_union(p.L, static_cast<int>(g), static_cast<int>(g - p.w + 1));
            }
            u = m; ul = ml; ur = mr;
        }
    }
}

// Flatten to roots, publish the per-word root bitmask and count roots per chunk.
// grid = batch * nchunks blocks of 256 threads; a chunk = 128 consecutive words of one image (raster order).
__global__ void __launch_bounds__(256) pp_flatten_count_kernel(const PP p) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int n = blockIdx.x / p.nchunks, ck = blockIdx.x % p.nchunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wimg = p.h * p.wpr;
    int cnt = 0;
    // (row, word-in-row) of this warp's words advance by additions: the kernels of this file are bound by instruction issue,
    // and a division by a run-time divisor is ~25 of them
    const int wl0 = ck * kChunkWords + warp;
    int fy = wl0 / p.wpr, fx = wl0 - fy * p.wpr;
    for (int i = warp; i < kChunkWords; i += 8, fx += 8) {
        const int wl = ck * kChunkWords + i;
        if (wl >= wimg) break;
        while (fx >= p.wpr) { fx -= p.wpr; ++fy; }
        const uint32_t wd = static_cast<uint32_t>(n) * wimg + wl;
        const uint32_t m = p.mask[wd];
        uint32_t rm = 0;
        if (m) {
            const uint32_t g = (static_cast<uint32_t>(n) * p.h + fy) * p.w + fx * 32 + lane;
            bool root = false;
            if ((m >> lane) & 1u) {
                const int r = uf_find(p.L, static_cast<int>(g));
                p.L[g] = r;  // only ever shortens paths; roots are stable after the merge kernel
                root = (r == static_cast<int>(g));
            }
            rm = __ballot_sync(0xffffffffu, root);
        }
        if (lane == 0) {
            p.rootmask[wd] = rm;
            cnt += __popc(rm);
        }
    }
    if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) p.chunk_cnt[blockIdx.x] = s_cnt;
}

// Exclusive scan of the chunk counts of one image (one block per image) + record reset.
__global__ void __launch_bounds__(256) pp_scan_kernel(const PP p) {
    __shared__ int s_part[256];
    __shared__ int s_carry;
    const int n = blockIdx.x;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < p.nchunks; c0 += 256) {
        const int c = c0 + threadIdx.x;
        const int v = (c < p.nchunks) ? p.chunk_cnt[n * p.nchunks + c] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {  // Hillis-Steele inclusive scan
            const int t = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        const int carry = s_carry;
        if (c < p.nchunks) p.chunk_cnt[n * p.nchunks + c] = carry + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + s_part[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.ncomp[n] = s_carry;
    for (int i = threadIdx.x; i < p.maxc; i += 256) {
        CompRec r;
        r.minx = 0x7fffffff; r.miny = 0x7fffffff; r.maxx = -1; r.maxy = -1; r.npix = 0; r.pad = 0; r.qsum = 0ull;
        p.rec[static_cast<int64_t>(n) * p.maxc + i] = r;
    }
}

// Rank roots: R[root] = 1 + (#roots before it in raster order), from the root bitmasks alone.  Same grid as above.
__global__ void __launch_bounds__(256) pp_rank_kernel(const PP p) {
    __shared__ int s_pre[kChunkWords];
    const int n = blockIdx.x / p.nchunks, ck = blockIdx.x % p.nchunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wimg = p.h * p.wpr;
    const uint32_t wbase = static_cast<uint32_t>(n) * wimg + static_cast<uint32_t>(ck) * kChunkWords;
    // inclusive scan of the 128 per-word root counts (threads 0..127, Hillis-Steele in shared memory)
    int c = 0;
    if (threadIdx.x < kChunkWords && ck * kChunkWords + static_cast<int>(threadIdx.x) < wimg) c = __popc(p.rootmask[wbase + threadIdx.x]);
    if (threadIdx.x < kChunkWords) s_pre[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < kChunkWords; o <<= 1) {
        int t = 0;
        if (threadIdx.x < kChunkWords && static_cast<int>(threadIdx.x) >= o) t = s_pre[threadIdx.x - o];
        __syncthreads();
        if (threadIdx.x < kChunkWords) s_pre[threadIdx.x] += t;
        __syncthreads();
    }
    const int base = p.chunk_cnt[blockIdx.x];
    for (int i = warp; i < kChunkWords; i += 8) {
        const int wl = ck * kChunkWords + i;
        if (wl >= wimg) break;
        const uint32_t rm = p.rootmask[wbase + i];
        if ((rm >> lane) & 1u) {
            const uint32_t g = ((wbase + i) / p.wpr) * p.w + (wl % p.wpr) * 32 + lane;
            const int before = (i ? s_pre[i - 1] : 0) + __popc(rm & ((1u << lane) - 1u));
            p.R[g] = base + before + 1;
        }
    }
}

// Write the canonical label map and accumulate per-component statistics.  One warp = one 32-pixel word; empty words
// cost one 4-byte load (plus the zero labels, when a label map is requested).  Blocks walk contiguous word ranges and
// keep ONE component record in shared memory (the first one they meet): a page-sized component then costs shared-memory
// atomics plus one global flush per block instead of six same-address global atomics per run.
__global__ void __launch_bounds__(256) pp_stats_kernel(const PP p) {
    __shared__ int s_key;  // image * maxc + (label - 1) of the cached record, -1 = none yet
    __shared__ int s_minx, s_miny, s_maxx, s_maxy, s_npix;
    __shared__ unsigned long long s_qsum;
    if (threadIdx.x == 0) {
        s_key = -1; s_minx = 0x7fffffff; s_miny = 0x7fffffff; s_maxx = -1; s_maxy = -1; s_npix = 0; s_qsum = 0ull;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwords = static_cast<uint32_t>(p.batch) * p.h * p.wpr;
    const uint32_t per_block = (nwords + gridDim.x - 1) / gridDim.x;
    const uint32_t w_begin = blockIdx.x * per_block;
    const uint32_t w_end = (w_begin + per_block < nwords) ? w_begin + per_block : nwords;
    // runs of the block's cached component are summed in registers first (every lane its own run heads) and reach shared
    // memory once per warp: on a page-sized component the eight warps of a block were serialising six same-address
    // shared-memory atomics per word
    int a_minx = 0x7fffffff, a_miny = 0x7fffffff, a_maxx = -1, a_maxy = -1, a_npix = 0;
    unsigned long long a_qsum = 0ull;
    uint32_t row = (w_begin + warp) / p.wpr;
    int wx = static_cast<int>(w_begin + warp - row * p.wpr) - 8;
    int n = static_cast<int>(row / p.h);
    int y = static_cast<int>(row - static_cast<uint32_t>(n) * p.h);
    for (uint32_t wd = w_begin + warp; wd < w_end; wd += 8) {
        wx += 8;  // (row, word-in-row, image) by additions, see pp_flatten_count_kernel
        while (wx >= p.wpr) {
            wx -= p.wpr;
            ++row;
            if (++y == p.h) { y = 0; ++n; }
        }
        const uint32_t m = p.mask[wd];
        if (m == 0 && !p.labels) continue;
        const int x = wx * 32 + lane;
        const uint32_t g = row * p.w + x;
        if (m == 0) {
            if (x < p.w) p.labels[g] = 0;
            continue;
        }
        int lab = 0;
        uint32_t q = 0;
        const bool fg = (m >> lane) & 1u;
        const bool run_head = fg && (lane == 0 || !((m >> (lane - 1)) & 1u));
        if (fg) {
            // the label costs two dependent loads: every pixel needs it for the label map, the statistics only at run heads
            if (p.labels || run_head) lab = p.R[p.L[g]];
            q = __float2uint_rn(p.prob[g] * 16777216.0f);
        }
        if (p.labels && x < p.w) p.labels[g] = lab;
        const uint32_t zabove = ~m & ~((2u << lane) - 1u);  // background lanes strictly above this lane
        const int end = zabove ? (__ffs(zabove) - 1) : 32;   // exclusive end of this lane's run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, q, o);
            if (lane + o < end) q += t;
        }
        const bool head = run_head && (lab != 0);
        if (head && lab <= p.maxc) {
            const int key = n * p.maxc + (lab - 1);
            int cur = s_key;
            if (cur == -1) cur = atomicCAS(&s_key, -1, key), cur = (cur == -1) ? key : cur;
            if (cur == key) {
                a_minx = min(a_minx, x);
                a_maxx = max(a_maxx, x + (end - lane) - 1);
                a_miny = min(a_miny, y);
                a_maxy = max(a_maxy, y);
                a_npix += end - lane;
                a_qsum += static_cast<unsigned long long>(q);
            } else {
                CompRec* r = p.rec + key;
                atomicMin(&r->minx, x);
                atomicMax(&r->maxx, x + (end - lane) - 1);
                atomicMin(&r->miny, y);
                atomicMax(&r->maxy, y);
                atomicAdd(&r->npix, end - lane);
                atomicAdd(&r->qsum, static_cast<unsigned long long>(q));
            }
        }
    }
    a_minx = __reduce_min_sync(0xffffffffu, a_minx);
    a_miny = __reduce_min_sync(0xffffffffu, a_miny);
    a_maxx = __reduce_max_sync(0xffffffffu, a_maxx);
    a_maxy = __reduce_max_sync(0xffffffffu, a_maxy);
    a_npix = __reduce_add_sync(0xffffffffu, a_npix);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a_qsum += __shfl_down_sync(0xffffffffu, a_qsum, o);
    if (lane == 0 && a_npix > 0) {
        atomicMin(&s_minx, a_minx);
        atomicMax(&s_maxx, a_maxx);
        atomicMin(&s_miny, a_miny);
        atomicMax(&s_maxy, a_maxy);
        atomicAdd(&s_npix, a_npix);
        atomicAdd(&s_qsum, a_qsum);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_key >= 0) {
        CompRec* r = p.rec + s_key;
        atomicMin(&r->minx, s_minx);
        atomicMax(&r->maxx, s_maxx);
        atomicMin(&r->miny, s_miny);
        atomicMax(&r->maxy, s_maxy);
        atomicAdd(&r->npix, s_npix);
        atomicAdd(&r->qsum, s_qsum);
    }
}

__device__ __forceinline__ int round_away(double v) {  // Clipper's Round(): half away from zero
    return static_cast<int>(v < 0.0 ? (v - 0.5) : (v + 0.5));
}

// Filters (src/det/test.py:77,81) + rectangle un-clip (test.py:37-43) + label-ordered compaction.
__global__ void __launch_bounds__(256) pp_boxes_kernel(const PP p) {
    __shared__ int s_part[256];
    __shared__ int s_base;
    const int n = blockIdx.x;
    const int nc = min(p.ncomp[n], p.maxc);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nc; c0 += 256) {
        const int c = c0 + threadIdx.x;
        bool keep = false;
        b200_box bx;
        if (c < nc) {
            const CompRec r = p.rec[static_cast<int64_t>(n) * p.maxc + c];
            const int bw = r.maxx - r.minx, bh = r.maxy - r.miny;  // pixel-centre extents (cv2 contour convention)
            const long long area = static_cast<long long>(bw) * bh;
            const bool score_ok =
                static_cast<double>(r.qsum) >= __dmul_rn(p.box_thresh_q, static_cast<double>(r.npix));
            keep = score_ok && bw >= 1 && bh >= 1 && area >= p.min_area;
            if (keep) {
                const double d = __ddiv_rn(__dmul_rn(static_cast<double>(area), p.ratio),
                                           static_cast<double>(2 * (bw + bh)));
                bx.x0 = round_away(static_cast<double>(r.minx) - d);
                bx.y0 = round_away(static_cast<double>(r.miny) - d);
                bx.x1 = round_away(static_cast<double>(r.maxx) + d);
                bx.y1 = round_away(static_cast<double>(r.maxy) + d);
                bx.bx0 = r.minx; bx.by0 = r.miny; bx.bx1 = r.maxx; bx.by1 = r.maxy;
                bx.label = c + 1;
                bx.npix = r.npix;
                bx.qsum = r.qsum;
            }
        }
        s_part[threadIdx.x] = keep ? 1 : 0;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int t = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        const int base = s_base;
        if (keep) p.boxes[static_cast<int64_t>(n) * p.maxc + base + s_part[threadIdx.x] - 1] = bx;
        __syncthreads();
        if (threadIdx.x == 255) s_base = base + s_part[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.counts[n] = s_base;
}

static int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_postproc_workspace_bytes(int batch, int h, int w, int max_candidates) {
    if (batch <= 0 || h <= 0 || w <= 0 || max_candidates <= 0) return -1;
    const int64_t npx = static_cast<int64_t>(batch) * h * w;
    const int64_t wimg = static_cast<int64_t>(h) * ((w + 31) / 32);
    const int64_t nchunks = (wimg + kChunkWords - 1) / kChunkWords;
    return 2 * align256(npx * 4) + 2 * align256(batch * wimg * 4) + align256(batch * nchunks * 4) + align256(batch * 4) +
           align256(static_cast<int64_t>(batch) * max_candidates * sizeof(CompRec));
}

extern "C" int b200_postproc(const float* prob, int32_t* labels, b200_box* boxes, int32_t* counts, int32_t* ncomp,
                             void* workspace, int64_t workspace_bytes, int batch, int h, int w, double thresh,
                             double box_thresh, double unclip_ratio, int min_area, int max_candidates, void* stream) {
    if (!prob || !boxes || !counts || !workspace) return set_error("b200_postproc: null pointer");
    const int64_t need = b200_postproc_workspace_bytes(batch, h, w, max_candidates);
    if (need < 0) return set_error("b200_postproc: bad shape");
    if (workspace_bytes < need)
        return set_error("b200_postproc: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    const int64_t npx = static_cast<int64_t>(batch) * h * w;
    if (npx >= 0x7fffffffLL) return set_error("b200_postproc: batch*h*w must be < 2^31");
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error("b200_postproc: workspace must be 256-byte aligned");
    PP p;
    p.prob = prob;
    p.labels = labels;
    p.boxes = boxes;
    p.counts = counts;
    p.batch = batch; p.h = h; p.w = w;
    p.wpr = (w + 31) / 32;
    const int64_t wimg = static_cast<int64_t>(h) * p.wpr;
    p.nchunks = static_cast<int>((wimg + kChunkWords - 1) / kChunkWords);
    p.maxc = max_candidates;
    p.min_area = min_area;
    p.thresh = static_cast<float>(thresh);
    p.box_thresh_q = box_thresh * 16777216.0;
    p.ratio = unclip_ratio;
    char* ws = reinterpret_cast<char*>(workspace);
    p.L = reinterpret_cast<int*>(ws); ws += align256(npx * 4);
    p.R = reinterpret_cast<int*>(ws); ws += align256(npx * 4);
    p.mask = reinterpret_cast<uint32_t*>(ws); ws += align256(batch * wimg * 4);
    p.rootmask = reinterpret_cast<uint32_t*>(ws); ws += align256(batch * wimg * 4);
    p.chunk_cnt = reinterpret_cast<int*>(ws); ws += align256(static_cast<int64_t>(batch) * p.nchunks * 4);
    int* ncomp_ws = reinterpret_cast<int*>(ws); ws += align256(batch * 4);
    p.rec = reinterpret_cast<CompRec*>(ws);
    p.ncomp = ncomp ? ncomp : ncomp_ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t nwords = static_cast<int64_t>(batch) * wimg;
    const int cap = 148 * 8;  // 8 resident 256-thread blocks per SM
    const int wblocks = static_cast<int>(((nwords + 7) / 8 < cap) ? (nwords + 7) / 8 : cap);
    const int64_t nrows = static_cast<int64_t>(batch) * h;
    const int rblocks = static_cast<int>(((nrows + 7) / 8 < cap) ? (nrows + 7) / 8 : cap);
    pp_init_kernel<<<rblocks, 256, 0, st>>>(p);
    const int64_t mitems = static_cast<int64_t>(batch) * ((h + kMergeRows - 1) / kMergeRows) * p.wpr;
    const int mblocks = static_cast<int>(((mitems + 7) / 8 < cap) ? (mitems + 7) / 8 : cap);
    pp_merge_kernel<<<mblocks, 256, 0, st>>>(p);
    pp_flatten_count_kernel<<<batch * p.nchunks, 256, 0, st>>>(p);
    pp_scan_kernel<<<batch, 256, 0, st>>>(p);
    pp_rank_kernel<<<batch * p.nchunks, 256, 0, st>>>(p);
    pp_stats_kernel<<<wblocks, 256, 0, st>>>(p);
    pp_boxes_kernel<<<batch, 256, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "b200_postproc launch");
}
