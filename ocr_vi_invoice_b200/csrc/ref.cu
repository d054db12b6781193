// Plain CUDA-core cross-check kernels: one thread per (output pixel, output channel), fp32
// accumulation over bf16 operands in the same K order and with the same epilogue flags as the
// tcgen05 path.  Slow on purpose; the GPU tests use them to localise a tensor-core / TMA bug to a
// layer, the engine never calls them.
#include "common.cuh"
#include "internal.h"

namespace b200 {

struct RefParams {
    const __nv_bfloat16* x;
    const __nv_bfloat16* w;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    void* y;
    const float* offsets;
    int batch, in_h, in_w, in_c, out_h, out_w, out_c, w_ld, ksize, stride, flags, out_ld, res_ld, shuf_c;
};

__device__ __forceinline__ void ref_store(const RefParams& p, float acc, int n, int oy, int ox, int co) {
    float v = fmaf(acc, p.scale[co], p.shift[co]);
    int64_t opix;
    int ch = co;
    if (p.flags & B200_EPI_SHUFFLE2) {
        const int sub = co / p.shuf_c;
        ch = co - sub * p.shuf_c;
        opix = (static_cast<int64_t>(n) * (2 * p.out_h) + (2 * oy + (sub >> 1))) * (2 * p.out_w) + (2 * ox + (sub & 1));
    } else {
        opix = (static_cast<int64_t>(n) * p.out_h + oy) * p.out_w + ox;
    }
    if (p.flags & (B200_EPI_RES | B200_EPI_RES_UP2)) {
        const int64_t rpix = (p.flags & B200_EPI_RES_UP2)
                                 ? (static_cast<int64_t>(n) * (p.out_h >> 1) + (oy >> 1)) * (p.out_w >> 1) + (ox >> 1)
                                 : opix;
        v += __bfloat162float(p.residual[rpix * p.res_ld + ch]);
    }
    if (p.flags & B200_EPI_RELU) v = fmaxf(v, 0.0f);
    if (p.flags & B200_EPI_OUT_F32)
        reinterpret_cast<float*>(p.y)[opix * p.out_ld + ch] = v;
    else
        reinterpret_cast<__nv_bfloat16*>(p.y)[opix * p.out_ld + ch] = __float2bfloat16_rn(v);
}

__global__ void ref_conv_kernel(const RefParams p) {
    const int64_t total = static_cast<int64_t>(p.batch) * p.out_h * p.out_w * p.out_c;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % p.out_c);
        int64_t r = i / p.out_c;
        const int ox = static_cast<int>(r % p.out_w);
        r /= p.out_w;
        const int oy = static_cast<int>(r % p.out_h);
        const int n = static_cast<int>(r / p.out_h);
        const int pad = (p.ksize - 1) / 2;
        float acc = 0.0f;
        for (int ky = 0; ky < p.ksize; ++ky) {
            const int iy = oy * p.stride - pad + ky;
            if (iy < 0 || iy >= p.in_h) continue;
            for (int kx = 0; kx < p.ksize; ++kx) {
                const int ix = ox * p.stride - pad + kx;
                if (ix < 0 || ix >= p.in_w) continue;
                const __nv_bfloat16* xs = p.x + ((static_cast<int64_t>(n) * p.in_h + iy) * p.in_w + ix) * p.in_c;
                // K order (64-channel chunk, tap, channel in chunk), as the tensor-core path
                const int taps = p.ksize * p.ksize, tap = ky * p.ksize + kx;
                const __nv_bfloat16* ws = p.w + static_cast<int64_t>(co) * p.w_ld;
                for (int c = 0; c < p.in_c; ++c)
                    acc = fmaf(__bfloat162float(xs[c]), __bfloat162float(ws[((c >> 6) * taps + tap) * 64 + (c & 63)]), acc);
            }
        }
        ref_store(p, acc, n, oy, ox, co);
    }
}

// torchvision deform_conv2d semantics (model/det/dcn.py:48-57): the sampled, mask-modulated value is
// rounded to bf16 before the product, as the tensor-core path stores it in the A tile.
__global__ void ref_dcn_kernel(const RefParams p) {
    const int64_t total = static_cast<int64_t>(p.batch) * p.out_h * p.out_w * p.out_c;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % p.out_c);
        int64_t r = i / p.out_c;
        const int ox = static_cast<int>(r % p.out_w);
        r /= p.out_w;
        const int oy = static_cast<int>(r % p.out_h);
        const int n = static_cast<int>(r / p.out_h);
        const float* off = p.offsets + ((static_cast<int64_t>(n) * p.out_h + oy) * p.out_w + ox) * 32;
        const __nv_bfloat16* img = p.x + static_cast<int64_t>(n) * p.in_h * p.in_w * p.in_c;
        float acc = 0.0f;
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
            const float m = 1.0f / (1.0f + expf(-off[18 + tap]));
            const float h = static_cast<float>(oy * p.stride - 1 + ky) + off[2 * tap];
            const float w = static_cast<float>(ox * p.stride - 1 + kx) + off[2 * tap + 1];
            if (!(h > -1.0f && w > -1.0f && h < static_cast<float>(p.in_h) && w < static_cast<float>(p.in_w))) continue;
            const float hf = floorf(h), wf = floorf(w);
            const int hl = static_cast<int>(hf), wl = static_cast<int>(wf), hh = hl + 1, wh = wl + 1;
            const float lh = h - hf, lw = w - wf, uh = 1.0f - lh, uw = 1.0f - lw;
            const bool t_ok = hl >= 0, b_ok = hh <= p.in_h - 1, l_ok = wl >= 0, r_ok = wh <= p.in_w - 1;
            const float w1 = (t_ok && l_ok) ? m * (uh * uw) : 0.0f;
            const float w2 = (t_ok && r_ok) ? m * (uh * lw) : 0.0f;
            const float w3 = (b_ok && l_ok) ? m * (lh * uw) : 0.0f;
            const float w4 = (b_ok && r_ok) ? m * (lh * lw) : 0.0f;
            const __nv_bfloat16* p1 = img + (static_cast<int64_t>(t_ok && l_ok ? hl : 0) * p.in_w + (t_ok && l_ok ? wl : 0)) * p.in_c;
            const __nv_bfloat16* p2 = img + (static_cast<int64_t>(t_ok && r_ok ? hl : 0) * p.in_w + (t_ok && r_ok ? wh : 0)) * p.in_c;
            const __nv_bfloat16* p3 = img + (static_cast<int64_t>(b_ok && l_ok ? hh : 0) * p.in_w + (b_ok && l_ok ? wl : 0)) * p.in_c;
            const __nv_bfloat16* p4 = img + (static_cast<int64_t>(b_ok && r_ok ? hh : 0) * p.in_w + (b_ok && r_ok ? wh : 0)) * p.in_c;
            const __nv_bfloat16* ws = p.w + static_cast<int64_t>(co) * p.w_ld;
            for (int c = 0; c < p.in_c; ++c) {
                const float s = w1 * __bfloat162float(p1[c]) + w2 * __bfloat162float(p2[c]) +
                                w3 * __bfloat162float(p3[c]) + w4 * __bfloat162float(p4[c]);
                acc = fmaf(__bfloat162float(__float2bfloat16_rn(s)), __bfloat162float(ws[((c >> 6) * 9 + tap) * 64 + (c & 63)]), acc);
            }
        }
        ref_store(p, acc, n, oy, ox, co);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ref_conv2d(const void* x, const void* w, const float* scale, const float* shift,
                               const void* residual, void* y, int batch, int in_h, int in_w, int in_c, int out_c,
                               int w_ld, int ksize, int stride, int flags, int out_ld, int res_ld, int shuf_c,
                               void* stream) {
    if (!x || !w || !scale || !shift || !y) return set_error("b200_ref_conv2d: null pointer");
    RefParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.w = reinterpret_cast<const __nv_bfloat16*>(w);
    p.scale = scale; p.shift = shift;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.y = y;
    p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.in_c = in_c;
    p.out_h = in_h / stride; p.out_w = in_w / stride; p.out_c = out_c; p.w_ld = w_ld;
    p.ksize = ksize; p.stride = stride; p.flags = flags; p.out_ld = out_ld; p.res_ld = res_ld; p.shuf_c = shuf_c;
    ref_conv_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_cuda(cudaGetLastError(), "b200_ref_conv2d launch");
}

extern "C" int b200_ref_dcn2d(const void* x, const float* offsets, const void* w, const float* scale,
                              const float* shift, void* y, int batch, int in_h, int in_w, int in_c, int out_c,
                              int stride, int flags, int out_ld, void* stream) {
    if (!x || !w || !scale || !shift || !y || !offsets) return set_error("b200_ref_dcn2d: null pointer");
    RefParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.w = reinterpret_cast<const __nv_bfloat16*>(w);
    p.scale = scale; p.shift = shift; p.y = y; p.offsets = offsets;
    p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.in_c = in_c;
    p.out_h = in_h / stride; p.out_w = in_w / stride; p.out_c = out_c; p.w_ld = 9 * in_c;
    p.ksize = 3; p.stride = stride; p.flags = flags; p.out_ld = out_ld;
    ref_dcn_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_cuda(cudaGetLastError(), "b200_ref_dcn2d launch");
}
