// GPU replacement of DBPostProcessor.__call__ (src/det/test.py:55-106): threshold, 8-connected component
// labelling with a min-index union-find (row runs found with warp ballots, merged with atomicMin), canonical
// relabel (label = raster rank of the component's first pixel), per-component bounding box / pixel count /
// fixed-point probability sum (warp-segmented run reduction + atomics), score / area filters and the
// closed-form rectangle un-clip.  HBM-bound integer work; everything after the threshold is integer or
// uncontracted IEEE double so it is bit-exact against oracle/postproc_oracle.py.
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kChunk = 4096;  // pixels per scan chunk (256 threads x 16)

struct CompRec {  // per-component accumulator, 32 B
    int minx, miny, maxx, maxy;
    int npix, pad;
    unsigned long long qsum;
};

struct PP {
    const float* prob;
    int* L;       // union-find parent per pixel (global pixel index), -1 = background
    int* R;       // canonical rank, valid at root pixels
    int* labels;  // output label map (may be null)
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
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
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

// One warp = 32 consecutive pixels of a row.  L[p] = index of the first pixel of p's run inside this word.
__global__ void pp_init_kernel(const PP p) {
    const int lane = threadIdx.x & 31;
    const int64_t nwords = static_cast<int64_t>(p.batch) * p.h * p.wpr;
    for (int64_t wd = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; wd < nwords;
         wd += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const int wx = static_cast<int>(wd % p.wpr);
        const int64_t row = wd / p.wpr;  // n*h + y
        const int x = wx * 32 + lane;
        const int64_t g = row * p.w + x;
        const bool fg = (x < p.w) && (p.prob[g] > p.thresh);
        const uint32_t m = __ballot_sync(0xffffffffu, fg);
        if (fg) {
            // zeros strictly below this lane; the run starts right above the highest of them
            const uint32_t z = ~m & ((1u << lane) - 1u);
            const int start = z ? (32 - __clz(z)) : 0;
            p.L[g] = static_cast<int>(g) - (lane - start);
        } else if (x < p.w) {
            p.L[g] = -1;
        }
    }
}

// Merge runs across word boundaries and with the row above (8-connectivity), skipping unions that a
// horizontally adjacent pixel is guaranteed to perform.
__global__ void pp_merge_kernel(const PP p) {
    const int64_t total = static_cast<int64_t>(p.batch) * p.h * p.w;
    for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (p.L[g] < 0) continue;
        const int x = static_cast<int>(g % p.w);
        const int y = static_cast<int>((g / p.w) % p.h);
        const bool left = (x > 0) && (p.L[g - 1] >= 0);
        if (left && (x & 31) == 0) uf_union(p.L, static_cast<int>(g), static_cast<int>(g - 1));
        if (y == 0) continue;
        const int64_t u = g - p.w;
        const bool up = p.L[u] >= 0;
        const bool ul = (x > 0) && (p.L[u - 1] >= 0);
        if (up) {
            if (!(left && ul)) uf_union(p.L, static_cast<int>(g), static_cast<int>(u));
        } else {
            if (ul && !left) uf_union(p.L, static_cast<int>(g), static_cast<int>(u - 1));
            const bool ur = (x + 1 < p.w) && (p.L[u + 1] >= 0);
            const bool right = (x + 1 < p.w) && (p.L[g + 1] >= 0);
            if (ur && !right) uf_union(p.L, static_cast<int>(g), static_cast<int>(u + 1));
        }
    }
}

// Flatten to roots and count roots per chunk.  grid = batch * nchunks blocks of 256 threads.
__global__ void __launch_bounds__(256) pp_flatten_count_kernel(const PP p) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int n = blockIdx.x / p.nchunks, ck = blockIdx.x % p.nchunks;
    const int npx = p.h * p.w;
    const int64_t base = static_cast<int64_t>(n) * npx;
    int cnt = 0;
    for (int i = threadIdx.x; i < kChunk; i += 256) {
        const int l = ck * kChunk + i;
        if (l >= npx) break;
        const int64_t g = base + l;
        if (p.L[g] >= 0) {
            const int r = uf_find(p.L, static_cast<int>(g));
            p.L[g] = r;  // only ever shortens paths; roots are stable after the merge kernel
            cnt += (r == static_cast<int>(g));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
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

// Rank roots: R[root] = 1 + (#roots before it in raster order).  Same grid as flatten_count.
__global__ void __launch_bounds__(256) pp_rank_kernel(const PP p) {
    __shared__ int s_warp[8];
    const int n = blockIdx.x / p.nchunks, ck = blockIdx.x % p.nchunks;
    const int npx = p.h * p.w;
    const int64_t base = static_cast<int64_t>(n) * npx;
    int running = p.chunk_cnt[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = 0; i0 < kChunk; i0 += 256) {  // raster order: pass i0 covers 256 consecutive pixels
        const int l = ck * kChunk + i0 + threadIdx.x;
        const int64_t g = base + l;
        const bool root = (l < npx) && (p.L[g] == static_cast<int>(g));
        const uint32_t m = __ballot_sync(0xffffffffu, root);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) {
            const int c = s_warp[wq];
            before += (wq < warp) ? c : 0;
            tot += c;
        }
        if (root) p.R[g] = running + before + __popc(m & ((1u << lane) - 1u)) + 1;
        running += tot;
        __syncthreads();
    }
}

// Write the canonical label map and accumulate per-component statistics.  One warp = one 32-pixel word.
__global__ void pp_stats_kernel(const PP p) {
    const int lane = threadIdx.x & 31;
    const int64_t nwords = static_cast<int64_t>(p.batch) * p.h * p.wpr;
    for (int64_t wd = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; wd < nwords;
         wd += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const int wx = static_cast<int>(wd % p.wpr);
        const int64_t row = wd / p.wpr;
        const int y = static_cast<int>(row % p.h);
        const int n = static_cast<int>(row / p.h);
        const int x = wx * 32 + lane;
        const int64_t g = row * p.w + x;
        int lab = 0;
        uint32_t q = 0;
        if (x < p.w) {
            const int root = p.L[g];
            if (root >= 0) {
                lab = p.R[root];
                q = __float2uint_rn(p.prob[g] * 16777216.0f);
            }
            if (p.labels) p.labels[g] = lab;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, lab != 0);
        if (m == 0) continue;
        const uint32_t zabove = ~m & ~((2u << lane) - 1u);  // background lanes strictly above this lane
        const int end = zabove ? (__ffs(zabove) - 1) : 32;   // exclusive end of this lane's run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, q, o);
            if (lane + o < end) q += t;
        }
        const bool head = (lab != 0) && (lane == 0 || !((m >> (lane - 1)) & 1u));
        if (head && lab <= p.maxc) {
            CompRec* r = p.rec + static_cast<int64_t>(n) * p.maxc + (lab - 1);
            atomicMin(&r->minx, x);
            atomicMax(&r->maxx, x + (end - lane) - 1);
            atomicMin(&r->miny, y);
            atomicMax(&r->maxy, y);
            atomicAdd(&r->npix, end - lane);
            atomicAdd(&r->qsum, static_cast<unsigned long long>(q));
        }
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
    const int64_t nchunks = (static_cast<int64_t>(h) * w + kChunk - 1) / kChunk;
    return 2 * align256(npx * 4) + align256(batch * nchunks * 4) + align256(batch * 4) +
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
    p.nchunks = static_cast<int>((static_cast<int64_t>(h) * w + kChunk - 1) / kChunk);
    p.maxc = max_candidates;
    p.min_area = min_area;
    p.thresh = static_cast<float>(thresh);
    p.box_thresh_q = box_thresh * 16777216.0;
    p.ratio = unclip_ratio;
    char* ws = reinterpret_cast<char*>(workspace);
    p.L = reinterpret_cast<int*>(ws); ws += align256(npx * 4);
    p.R = reinterpret_cast<int*>(ws); ws += align256(npx * 4);
    p.chunk_cnt = reinterpret_cast<int*>(ws); ws += align256(static_cast<int64_t>(batch) * p.nchunks * 4);
    int* ncomp_ws = reinterpret_cast<int*>(ws); ws += align256(batch * 4);
    p.rec = reinterpret_cast<CompRec*>(ws);
    p.ncomp = ncomp ? ncomp : ncomp_ws;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t nwords = static_cast<int64_t>(batch) * h * p.wpr;
    const int wblocks = static_cast<int>(((nwords + 7) / 8 < 148 * 32) ? (nwords + 7) / 8 : 148 * 32);
    const int pblocks = static_cast<int>(((npx + 255) / 256 < 148 * 32) ? (npx + 255) / 256 : 148 * 32);
    pp_init_kernel<<<wblocks, 256, 0, st>>>(p);
    pp_merge_kernel<<<pblocks, 256, 0, st>>>(p);
    pp_flatten_count_kernel<<<batch * p.nchunks, 256, 0, st>>>(p);
    pp_scan_kernel<<<batch, 256, 0, st>>>(p);
    pp_rank_kernel<<<batch * p.nchunks, 256, 0, st>>>(p);
    pp_stats_kernel<<<wblocks, 256, 0, st>>>(p);
    pp_boxes_kernel<<<batch, 256, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "b200_postproc launch");
}
