// Detection -> recognition hand-over on the GPU: detector box records -> rescale to the
// original page -> crop_image -> preprocess_for_recognition, one launch for a whole batch of pages, so the boxes never
// travel to the host between the detector and the recogniser.
//
// Reference (file:line under the reference tree), restated in oracle/recprep_oracle.py and pinned to the reference's own
// functions by tests/golden/recprep_reference.npz:
//   box rescale                     src/pipeline/pipeline2.py:323-328   (float64 quotient truncated toward zero)
//   crop_image                      src/det/test.py:123-130             (cv2.boundingRect, clamps, slice)
//   preprocess_for_recognition      src/pipeline/pipeline2.py:92-128    (cv2.resize 8-bit linear, white pad, normalise)
//   blank tensor for an empty crop  src/pipeline/pipeline2.py:153-157
//
// Byte / integer work, HBM-bound: one block per crop slot.  The block computes the crop rectangle once, the y
// coefficients once per row (threads 0..th-1 -> shared memory), the x coefficients once per thread (fp64 set-up
// amortised over the th rows), and a 768-entry normalisation table ((v / 255 - mean) / std with IEEE divisions, exactly
// the reference's fp32 arithmetic) once; the inner loop is 12 byte loads (L1-resident crop rows), integer blending,
// three table look-ups and three row-coalesced fp32 stores per output pixel.
#include "imgproc.cuh"
#include "internal.h"

namespace b200 {

constexpr int kRecMaxImages = B200_REC_MAX_IMAGES;
constexpr int kRecMaxTh = 64;
constexpr int kRecThreads = 128;  // 48 registers: 10 blocks per SM, a 16-page batch (~1000 slots) is resident in one wave

struct RecImages {
    const uint8_t* pixels[kRecMaxImages];
    double scale_h[kRecMaxImages], scale_w[kRecMaxImages];
    int32_t h[kRecMaxImages], w[kRecMaxImages], pitch[kRecMaxImages];
    int32_t n;
};

struct RecParams {
    const b200_box* boxes;         // kFromBoxes: [n_images][max_candidates]
    const int32_t* counts;         // kFromBoxes: [n_images]
    const b200_rec_rect* rects_in; // !kFromBoxes: [n] explicit rectangles
    int32_t n_rects;
    int32_t max_candidates;
    float* out;                    // [capacity][3][th][tw]
    b200_rec_rect* rects;          // [capacity]
    int32_t* total;                // [1]
    int32_t capacity, th, tw;
    float mean[3], stdv[3];
};

// cv2.boundingRect of the rescaled rectangle + the clamps of crop_image; bw / bh <= 0 means an empty crop.
__device__ __forceinline__ void rect_from_box(const b200_box& r, double scale_h, double scale_w, int img_h, int img_w, int& x,
                                              int& y, int& bw, int& bh) {
    // box[:, 0] = box[:, 0] / scale_w on an int64 array: IEEE double quotient, C cast (toward zero), then astype(int32)
    const int tx0 = static_cast<int>(static_cast<long long>(__ddiv_rn(static_cast<double>(r.x0), scale_w)));
    const int tx1 = static_cast<int>(static_cast<long long>(__ddiv_rn(static_cast<double>(r.x1), scale_w)));
    const int ty0 = static_cast<int>(static_cast<long long>(__ddiv_rn(static_cast<double>(r.y0), scale_h)));
    const int ty1 = static_cast<int>(static_cast<long long>(__ddiv_rn(static_cast<double>(r.y1), scale_h)));
    const int xmin = min(tx0, tx1), xmax = max(tx0, tx1), ymin = min(ty0, ty1), ymax = max(ty0, ty1);
    x = max(0, xmin);  // a negative corner shifts the crop, it does not shrink it (test.py:126-129)
    y = max(0, ymin);
    bw = min(xmax - xmin + 1, img_w - x);
    bh = min(ymax - ymin + 1, img_h - y);
}

template <bool kFromBoxes>
__global__ void __launch_bounds__(kRecThreads) rec_prep_kernel(const __grid_constant__ RecImages imgs,
                                                                const __grid_constant__ RecParams p) {
    __shared__ float s_norm[768];
    __shared__ int s_prefix[kRecMaxImages + 1];
    __shared__ int s_y0[kRecMaxTh], s_y1[kRecMaxTh], s_b0[kRecMaxTh], s_b1[kRecMaxTh];
    const int tid = threadIdx.x;

    for (int i = tid; i < 768; i += kRecThreads) {
        const int c = i >> 8;
        // torch: x.float() / 255.0, then (x - mean) / std, all fp32 with IEEE rounding
        s_norm[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), p.mean[c]), p.stdv[c]);
    }
    int n_slots;
    if (kFromBoxes) {
        if (tid < imgs.n) s_prefix[tid + 1] = min(max(__ldg(p.counts + tid), 0), p.max_candidates);  // loads in parallel
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int i = 0; i < imgs.n; ++i) {
                const int c = s_prefix[i + 1];
                s_prefix[i] = acc;
                acc += c;
            }
            s_prefix[imgs.n] = acc;
            if (blockIdx.x == 0) *p.total = acc;
        }
        __syncthreads();
        n_slots = min(s_prefix[imgs.n], p.capacity);
    } else {
        if (blockIdx.x == 0 && tid == 0) *p.total = p.n_rects;
        __syncthreads();
        n_slots = min(p.n_rects, p.capacity);
    }

    const int th = p.th, tw = p.tw;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        int im, x, y, bw, bh;
        if (kFromBoxes) {
            im = 0;
            while (im + 1 < imgs.n && slot >= s_prefix[im + 1]) ++im;
            const b200_box r = p.boxes[static_cast<int64_t>(im) * p.max_candidates + (slot - s_prefix[im])];
            rect_from_box(r, imgs.scale_h[im], imgs.scale_w[im], imgs.h[im], imgs.w[im], x, y, bw, bh);
        } else {
            const b200_rec_rect r = p.rects_in[slot];
            im = min(max(r.image, 0), imgs.n - 1);
            x = r.x; y = r.y; bw = r.w; bh = r.h;
            if (x < 0 || y < 0 || x + bw > imgs.w[im] || y + bh > imgs.h[im]) bw = bh = 0;  // not a slice of the page: empty
        }
        int status = B200_REC_OK, new_w = 0;
        if (bw <= 0 || bh <= 0) {
            status = B200_REC_EMPTY;
        } else {
            // new_w = int(w * (target_h / h)): the quotient is rounded before the product (pipeline2.py:101-102)
            new_w = static_cast<int>(__dmul_rn(static_cast<double>(bw), __ddiv_rn(static_cast<double>(th), static_cast<double>(bh))));
            if (new_w == 0) status = B200_REC_DEGENERATE;
        }
        if (tid == 0) p.rects[slot] = b200_rec_rect{im, x, y, bw, bh, status};
        float* o = p.out + static_cast<int64_t>(slot) * 3 * th * tw;
        const int64_t plane = static_cast<int64_t>(th) * tw;

        if (status != B200_REC_OK) {  // torch.zeros(3, th, tw) for an empty crop; the same for a crop cv2.resize refuses
            for (int64_t i = tid; i < 3 * plane; i += kRecThreads) o[i] = 0.0f;
            continue;
        }
        const int dst_w = new_w > tw ? tw : new_w;  // wider than the target: both axes scaled (pipeline2.py:104-105)
        const double scale_x = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst_w), static_cast<double>(bw)));
        const double scale_y = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(th), static_cast<double>(bh)));
        __syncthreads();  // previous slot's readers of s_y* are done
        if (tid < th) resize_coef(tid, scale_y, bh, false, s_y0[tid], s_y1[tid], s_b0[tid], s_b1[tid]);
        __syncthreads();
        const uint8_t* base = imgs.pixels[im] + static_cast<int64_t>(y) * imgs.pitch[im] + static_cast<int64_t>(x) * 3;
        const int pitch = imgs.pitch[im];
        const float pad0 = s_norm[255], pad1 = s_norm[256 + 255], pad2 = s_norm[512 + 255];  // white right padding
        for (int dx = tid; dx < tw; dx += kRecThreads) {
            if (dx >= dst_w) {
                for (int dy = 0; dy < th; ++dy) {
                    float* q = o + static_cast<int64_t>(dy) * tw + dx;
                    q[0] = pad0; q[plane] = pad1; q[2 * plane] = pad2;
                }
                continue;
            }
            int x0, x1, a0, a1;
            resize_coef(dx, scale_x, bw, true, x0, x1, a0, a1);
            x0 *= 3; x1 *= 3;
#pragma unroll 4
            for (int dy = 0; dy < th; ++dy) {
                const uint8_t* r0 = base + static_cast<int64_t>(s_y0[dy]) * pitch;
                const uint8_t* r1 = base + static_cast<int64_t>(s_y1[dy]) * pitch;
                const int b0 = s_b0[dy], b1 = s_b1[dy];
                float* q = o + static_cast<int64_t>(dy) * tw + dx;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int v = resize_blend(__ldg(r0 + x0 + c), __ldg(r0 + x1 + c), __ldg(r1 + x0 + c), __ldg(r1 + x1 + c), a0,
                                               a1, b0, b1);
                    q[c * plane] = s_norm[c * 256 + v];
                }
            }
        }
    }
}

static int fill_images(const b200_image* images, int n_images, RecImages& im, const char* who) {
    if (!images || n_images <= 0 || n_images > kRecMaxImages) return set_error("%s: 1..%d images per call", who, kRecMaxImages);
    im.n = n_images;
    for (int i = 0; i < n_images; ++i) {
        const b200_image& s = images[i];
        if (!s.pixels || s.h <= 0 || s.w <= 0 || s.pitch < 3 * s.w) return set_error("%s: image %d: bad shape or null pointer", who, i);
        if (!(s.scale_h > 0.0) || !(s.scale_w > 0.0)) return set_error("%s: image %d: scales must be positive", who, i);
        im.pixels[i] = static_cast<const uint8_t*>(s.pixels);
        im.h[i] = s.h; im.w[i] = s.w; im.pitch[i] = s.pitch;
        im.scale_h[i] = s.scale_h; im.scale_w[i] = s.scale_w;
    }
    return 0;
}

static int fill_params(RecParams& p, float* out, b200_rec_rect* rects, int32_t* total, int capacity, int th, int tw,
                       const float* mean3, const float* std3, const char* who) {
    if (!out || !rects || !total || !mean3 || !std3) return set_error("%s: null pointer", who);
    if (capacity <= 0 || th <= 0 || th > kRecMaxTh || tw <= 0 || tw > 4096) return set_error("%s: need capacity > 0, 1 <= th <= %d, 1 <= tw <= 4096", who, kRecMaxTh);
    if (std3[0] == 0.0f || std3[1] == 0.0f || std3[2] == 0.0f) return set_error("%s: zero std", who);
    p.out = out; p.rects = rects; p.total = total; p.capacity = capacity; p.th = th; p.tw = tw;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.stdv[c] = std3[c]; }
    return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rec_prep_boxes(const b200_image* images, int n_images, const b200_box* boxes, const int32_t* counts,
                                   int max_candidates, float* out, b200_rec_rect* rects, int32_t* total, int capacity, int th,
                                   int tw, const float* mean3, const float* std3, void* stream) {
    RecImages im{};
    RecParams p{};
    if (fill_images(images, n_images, im, "b200_rec_prep_boxes")) return 1;
    if (fill_params(p, out, rects, total, capacity, th, tw, mean3, std3, "b200_rec_prep_boxes")) return 1;
    if (!boxes || !counts || max_candidates <= 0) return set_error("b200_rec_prep_boxes: null boxes / counts");
    p.boxes = boxes; p.counts = counts; p.max_candidates = max_candidates;
    const int grid = capacity < 148 * 8 ? capacity : 148 * 8;  // the slot count lives on the device: blocks past it exit
    rec_prep_kernel<true><<<grid, kRecThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(im, p);
    return check_cuda(cudaGetLastError(), "b200_rec_prep_boxes launch");
}

extern "C" int b200_rec_prep_rects(const b200_image* images, int n_images, const b200_rec_rect* rects_in, int n, float* out,
                                   b200_rec_rect* rects, int32_t* total, int capacity, int th, int tw, const float* mean3,
                                   const float* std3, void* stream) {
    RecImages im{};
    RecParams p{};
    if (fill_images(images, n_images, im, "b200_rec_prep_rects")) return 1;
    if (fill_params(p, out, rects, total, capacity, th, tw, mean3, std3, "b200_rec_prep_rects")) return 1;
    if (!rects_in || n < 0) return set_error("b200_rec_prep_rects: null rectangles");
    p.rects_in = rects_in; p.n_rects = n;
    const int slots = n < capacity ? n : capacity;
    const int grid = slots < 1 ? 1 : (slots < 148 * 8 ? slots : 148 * 8);
    rec_prep_kernel<false><<<grid, kRecThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(im, p);
    return check_cuda(cudaGetLastError(), "b200_rec_prep_rects launch");
}
