// Bandwidth-bound kernels of the DBNet++ path: stem packing, max-pool, the fused adaptive-scale-fusion
// (ScaleFeatureSelection) and the DB head tail.  All activations NHWC bf16, 16-byte vector accesses.
#include "common.cuh"
#include "imgproc.cuh"
#include "internal.h"

namespace b200 {

// ------------------------------------------------------------------------------------------ stem pack
// fp32 NCHW image -> zero-padded 2x2 space-to-depth bf16 buffer [b][H/2+3][W/2+3][16];
// channel (dy*2+dx)*3 + c holds img[c][2*br+dy][2*bc+dx] of block (br, bc) = (row-2, col-2); 12..15 = 0.
// With this layout the 7x7 s2 p3 stem conv (torchvision resnet.py:197) is a 4x4 s1 conv whose A rows are
// contiguous 128-byte runs (4 blocks x 16 channels).
// Idx: uint32_t whenever the packed pixel count fits (a 64-bit division by a run-time divisor is ~100 instructions, and the
// flat index is split by three of them)
template <typename Idx>
__global__ void stem_pack_kernel(const float* __restrict__ img, uint4* __restrict__ out, int batch, int h, int w) {
    const int ph = h / 2 + 3, pw = w / 2 + 3;
    const Idx total = static_cast<Idx>(batch) * ph * pw;
    for (Idx i = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<Idx>(gridDim.x) * blockDim.x) {
        const int px = static_cast<int>(i % pw);
        Idx r = i / pw;
        const int py = static_cast<int>(r % ph);
        const int n = static_cast<int>(r / ph);
        const int br = py - 2, bc = px - 2;
        float v[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) v[j] = 0.0f;
        if (br >= 0 && br < h / 2 && bc >= 0 && bc < w / 2) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* src = img + ((static_cast<int64_t>(n) * 3 + c) * h + 2 * br) * w + 2 * bc;
                const float2 r0 = __ldg(reinterpret_cast<const float2*>(src));
                const float2 r1 = __ldg(reinterpret_cast<const float2*>(src + w));
                v[0 * 3 + c] = r0.x;
                v[1 * 3 + c] = r0.y;
                v[2 * 3 + c] = r1.x;
                v[3 * 3 + c] = r1.y;
            }
        }
        uint4 lo, hi;
        lo.x = pack_bf16(v[0], v[1]);
        lo.y = pack_bf16(v[2], v[3]);
        lo.z = pack_bf16(v[4], v[5]);
        lo.w = pack_bf16(v[6], v[7]);
        hi.x = pack_bf16(v[8], v[9]);
        hi.y = pack_bf16(v[10], v[11]);
        hi.z = 0u;
        hi.w = 0u;
        out[2 * i] = lo;
        out[2 * i + 1] = hi;
    }
}

// Same packing straight from the decoded image: uint8 HWC RGB -> (v / 255 - mean) / std (the reference's host-side
// normalisation, src/pipeline/pipeline2.py:312-314) -> bf16 space-to-depth.  4x less H2D traffic than fp32 NCHW.
template <typename Idx>
__global__ void stem_pack_u8_kernel(const uint8_t* __restrict__ img, uint4* __restrict__ out, int batch, int h, int w,
                                    float m0, float m1, float m2, float i0, float i1, float i2) {
    const int ph = h / 2 + 3, pw = w / 2 + 3;
    const Idx total = static_cast<Idx>(batch) * ph * pw;
    const float mean[3] = {m0, m1, m2}, inv[3] = {i0, i1, i2};
    for (Idx i = blockIdx.x * static_cast<Idx>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<Idx>(gridDim.x) * blockDim.x) {
        const int px = static_cast<int>(i % pw);
        Idx r = i / pw;
        const int py = static_cast<int>(r % ph);
        const int n = static_cast<int>(r / ph);
        const int br = py - 2, bc = px - 2;
        float v[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) v[j] = 0.0f;
        if (br >= 0 && br < h / 2 && bc >= 0 && bc < w / 2) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                // two horizontally adjacent RGB pixels = 6 consecutive bytes at an even byte offset
                const uint8_t* src = img + ((static_cast<int64_t>(n) * h + 2 * br + dy) * w + 2 * bc) * 3;
                const uint16_t a = *reinterpret_cast<const uint16_t*>(src), b = *reinterpret_cast<const uint16_t*>(src + 2),
                               c = *reinterpret_cast<const uint16_t*>(src + 4);
                const float px0[3] = {static_cast<float>(a & 0xff), static_cast<float>(a >> 8), static_cast<float>(b & 0xff)};
                const float px1[3] = {static_cast<float>(b >> 8), static_cast<float>(c & 0xff), static_cast<float>(c >> 8)};
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    v[(dy * 2 + 0) * 3 + ch] = (px0[ch] / 255.0f - mean[ch]) * inv[ch];
                    v[(dy * 2 + 1) * 3 + ch] = (px1[ch] / 255.0f - mean[ch]) * inv[ch];
                }
            }
        }
        uint4 lo, hi;
        lo.x = pack_bf16(v[0], v[1]);
        lo.y = pack_bf16(v[2], v[3]);
        lo.z = pack_bf16(v[4], v[5]);
        lo.w = pack_bf16(v[6], v[7]);
        hi.x = pack_bf16(v[8], v[9]);
        hi.y = pack_bf16(v[10], v[11]);
        hi.z = 0u;
        hi.w = 0u;
        out[2 * i] = lo;
        out[2 * i + 1] = hi;
    }
}

// ------------------------------------------------------------------------------------------ max-pool
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&m);
}

// MaxPool2d(kernel 3, stride 2, padding 1), NHWC bf16; one thread per (output pixel, 8 channels).
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int batch, int h,
                               int w, int c) {
    const int oh = h / 2, ow = w / 2, c8 = c / 8;
    const int64_t total = static_cast<int64_t>(batch) * oh * ow * c8;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cc = static_cast<int>(i % c8);
        int64_t r = i / c8;
        const int ox = static_cast<int>(r % ow);
        r /= ow;
        const int oy = static_cast<int>(r % oh);
        const int n = static_cast<int>(r / oh);
        const uint32_t ninf = 0xFF80FF80u;  // (-inf, -inf) bf16
        uint4 m = make_uint4(ninf, ninf, ninf, ninf);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= w) continue;
                const uint4 v = ldg128(x + ((static_cast<int64_t>(n) * h + iy) * w + ix) * c + cc * 8);
                m.x = bf16x2_max(m.x, v.x);
                m.y = bf16x2_max(m.y, v.y);
                m.z = bf16x2_max(m.z, v.z);
                m.w = bf16x2_max(m.w, v.w);
            }
        }
        *reinterpret_cast<uint4*>(y + ((static_cast<int64_t>(n) * oh + oy) * ow + ox) * c + cc * 8) = m;
    }
}

// Strip version: a thread owns 8 channels of one output column and walks kRows output rows downwards.  Input row 2oy+1 is
// the last row of output oy AND the first of output oy+1, so its 3-column maximum is carried in registers: (2 kRows + 1) x 3
// loads per kRows outputs instead of 9 each, and no 64-bit index arithmetic per output (the generic kernel spent 57 % of its
// issue slots on div / mod of the flat index).  Block = (256 / c8) output columns x c8 channel groups.
template <int kRows>
__global__ void __launch_bounds__(256) maxpool_strip_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            int h, int w, int c, int c8_shift) {
    const int oh = h >> 1, ow = w >> 1;
    const int cc = threadIdx.x & ((1 << c8_shift) - 1);
    const int ox = blockIdx.x * (256 >> c8_shift) + (threadIdx.x >> c8_shift);
    const int oy0 = blockIdx.y * kRows, n = blockIdx.z;
    if (ox >= ow) return;
    const uint32_t ninf = 0xFF80FF80u;  // (-inf, -inf) bf16
    const __nv_bfloat16* xin = x + static_cast<int64_t>(n) * h * w * c + cc * 8;
    const int ix0 = 2 * ox - 1;
    const bool left = ix0 >= 0;  // columns 2ox and 2ox+1 always exist (w even)
    auto row_max = [&](int iy) -> uint4 {
        const __nv_bfloat16* r = xin + (static_cast<int64_t>(iy) * w + ix0) * c;
        uint4 m = ldg128(r + c);
        const uint4 b = ldg128(r + 2 * c);
        m.x = bf16x2_max(m.x, b.x); m.y = bf16x2_max(m.y, b.y); m.z = bf16x2_max(m.z, b.z); m.w = bf16x2_max(m.w, b.w);
        if (left) {
            const uint4 a = ldg128(r);
            m.x = bf16x2_max(m.x, a.x); m.y = bf16x2_max(m.y, a.y); m.z = bf16x2_max(m.z, a.z); m.w = bf16x2_max(m.w, a.w);
        }
        return m;
    };
    uint4 carry = (oy0 > 0) ? row_max(2 * oy0 - 1) : make_uint4(ninf, ninf, ninf, ninf);
    __nv_bfloat16* yout = y + ((static_cast<int64_t>(n) * oh + oy0) * ow + ox) * c + cc * 8;
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
        const int oy = oy0 + j;
        if (oy >= oh) break;
        const uint4 a = row_max(2 * oy), b = row_max(2 * oy + 1);  // both exist (h even)
        uint4 m;
        m.x = bf16x2_max(bf16x2_max(carry.x, a.x), b.x);
        m.y = bf16x2_max(bf16x2_max(carry.y, a.y), b.y);
        m.z = bf16x2_max(bf16x2_max(carry.z, a.z), b.z);
        m.w = bf16x2_max(bf16x2_max(carry.w, a.w), b.w);
        *reinterpret_cast<uint4*>(yout + static_cast<int64_t>(j) * ow * c) = m;
        carry = b;
    }
}

// ------------------------------------------------------------------------------------------ ASF
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16lo(u.x); f[1] = bf16hi(u.x);
    f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
    f[4] = bf16lo(u.z); f[5] = bf16hi(u.z);
    f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

// bilinear, align_corners=True (ATen upsample_bilinear2d: src = dst * (in-1)/(out-1), computed in fp32)
__device__ __forceinline__ void bilinear_ac(const __nv_bfloat16* __restrict__ src, int n, int hin, int win, float sy,
                                            float sx, int y, int x, int lane, float (&f)[8]) {
    const float fy = sy * static_cast<float>(y), fx = sx * static_cast<float>(x);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + ((y0 < hin - 1) ? 1 : 0), x1 = x0 + ((x0 < win - 1) ? 1 : 0);
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const __nv_bfloat16* base = src + static_cast<int64_t>(n) * hin * win * 256 + lane * 8;
    float a[8], b[8], c[8], d[8];
    unpack8(ldg128(base + (static_cast<int64_t>(y0) * win + x0) * 256), a);
    unpack8(ldg128(base + (static_cast<int64_t>(y0) * win + x1) * 256), b);
    unpack8(ldg128(base + (static_cast<int64_t>(y1) * win + x0) * 256), c);
    unpack8(ldg128(base + (static_cast<int64_t>(y1) * win + x1) * 256), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = hy * (hx * a[j] + lx * b[j]) + ly * (hx * c[j] + lx * d[j]);
}

// ScaleFeatureSelection.forward (model/det/neck.py:57-79), one warp per output pixel, lane = 8 channels of
// each of the 4 levels.  The 1024-channel concat is never materialised.
__global__ void __launch_bounds__(256) asf_kernel(const __nv_bfloat16* __restrict__ p2, const __nv_bfloat16* __restrict__ p3,
                                                  const __nv_bfloat16* __restrict__ p4, const __nv_bfloat16* __restrict__ p5,
                                                  const float* __restrict__ w_att, const float* __restrict__ b_att,
                                                  __nv_bfloat16* __restrict__ out, int batch, int h2, int w2, int h3, int w3,
                                                  int h4, int w4, int h5, int w5) {
    __shared__ float sw[4 * 4 * 8 * 32];  // [logit i][level l][j][lane]  (conflict-free per-lane reads)
    for (int t = threadIdx.x; t < 4096; t += blockDim.x) {
        const int i = t >> 10, k = t & 1023;  // w_att[i][k], k = l*256 + lane*8 + j
        const int l = k >> 8, ln = (k >> 3) & 31, j = k & 7;
        sw[((i * 4 + l) * 8 + j) * 32 + ln] = w_att[t];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float sy3 = (h2 > 1) ? static_cast<float>(h3 - 1) / static_cast<float>(h2 - 1) : 0.0f;
    const float sx3 = (w2 > 1) ? static_cast<float>(w3 - 1) / static_cast<float>(w2 - 1) : 0.0f;
    const float sy4 = (h2 > 1) ? static_cast<float>(h4 - 1) / static_cast<float>(h2 - 1) : 0.0f;
    const float sx4 = (w2 > 1) ? static_cast<float>(w4 - 1) / static_cast<float>(w2 - 1) : 0.0f;
    const float sy5 = (h2 > 1) ? static_cast<float>(h5 - 1) / static_cast<float>(h2 - 1) : 0.0f;
    const float sx5 = (w2 > 1) ? static_cast<float>(w5 - 1) / static_cast<float>(w2 - 1) : 0.0f;
    const float b0 = b_att[0], b1 = b_att[1], b2 = b_att[2], b3 = b_att[3];
    // A block walks 8 x 16 pixel tiles (warp = tile row, 16 pixels along x each) so the bilinear taps of p3..p5, which
    // neighbouring pixels share, are fetched from L2 once per tile and re-read from L1.
    const int tiles_x = (w2 + 15) / 16, tiles_y = (h2 + 7) / 8;
    const int64_t ntiles = static_cast<int64_t>(batch) * tiles_x * tiles_y;
    const int wrow = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    for (int px = 0; px < 16; ++px) {
        const int tx = static_cast<int>(tile % tiles_x);
        const int64_t r = tile / tiles_x;
        const int tyi = static_cast<int>(r % tiles_y);
        const int n = static_cast<int>(r / tiles_y);
        const int x = tx * 16 + px, y = tyi * 8 + wrow;
        if (x >= w2 || y >= h2) continue;
        const int64_t pix = (static_cast<int64_t>(n) * h2 + y) * w2 + x;
        float f[4][8];
        unpack8(ldg128(p2 + pix * 256 + lane * 8), f[0]);
        bilinear_ac(p3, n, h3, w3, sy3, sx3, y, x, lane, f[1]);
        bilinear_ac(p4, n, h4, w4, sy4, sx4, y, x, lane, f[2]);
        bilinear_ac(p5, n, h5, w5, sy5, sx5, y, x, lane, f[3]);
        float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int l = 0; l < 4; ++l)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i] = fmaf(sw[((i * 4 + l) * 8 + j) * 32 + lane], f[l][j], a[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        }
        a[0] += b0; a[1] += b1; a[2] += b2; a[3] += b3;
        const float mx = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
        const float e0 = expf(a[0] - mx), e1 = expf(a[1] - mx), e2 = expf(a[2] - mx), e3 = expf(a[3] - mx);
        const float inv = 1.0f / (e0 + e1 + e2 + e3);
        const float s0 = e0 * inv, s1 = e1 * inv, s2 = e2 * inv, s3 = e3 * inv;
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = f[0][j] * s0 + f[1][j] * s1 + f[2][j] * s2 + f[3][j] * s3;
        uint4 u;
        u.x = pack_bf16(o8[0], o8[1]);
        u.y = pack_bf16(o8[2], o8[3]);
        u.z = pack_bf16(o8[4], o8[5]);
        u.w = pack_bf16(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(out + pix * 256 + lane * 8) = u;
    }
}

// ------------------------------------------------------------------------------------------ ASF, split form
// The attention logits are linear in the (bilinearly upsampled) features, and bilinear upsampling commutes with the
// per-pixel 1x1 conv:  conv_atten(cat(up(p_l))) = sum_l up(W_l p_l).  So z_l = W_l p_l (4 channels) is computed at
// each level's NATIVE resolution by four tiny GEMMs, and this kernel only (a) interpolates the z maps, (b) takes the
// softmax, (c) blends the 13 feature taps with pre-multiplied weights s_l * w_corner.  No dot products, no shuffles.
struct AsfLevel {
    const __nv_bfloat16* p;
    const float* z;  // [pixel][zld] fp32, first 4 entries = logit contributions
    int h, w;
    float sy, sx;
};
struct AsfParams {
    AsfLevel lv[4];
    const float* b_att;
    __nv_bfloat16* out;
    int batch, zld;
};

// One warp = 4 consecutive pixels of a row; lane = (pixel, channel group cg); the lane owns channels k*64 + cg*8 + j
// (k < 4, j < 8), so the 8 lanes of a pixel read 128 contiguous bytes per load instruction.
__global__ void __launch_bounds__(256, 3) asf_combine_kernel(const AsfParams P) {
    const int lane = threadIdx.x & 31;
    const int pq = lane >> 3, cg = lane & 7;
    const int h2 = P.lv[0].h, w2 = P.lv[0].w;
    const int quads_x = (w2 + 3) >> 2;
    const int64_t nquads = static_cast<int64_t>(P.batch) * h2 * quads_x;
    const float b0 = P.b_att[0], b1 = P.b_att[1], b2 = P.b_att[2], b3 = P.b_att[3];
    for (int64_t qd = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; qd < nquads;
         qd += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const int qx = static_cast<int>(qd % quads_x);
        const int64_t r = qd / quads_x;
        const int y = static_cast<int>(r % h2);
        const int n = static_cast<int>(r / h2);
        const int x = qx * 4 + pq;
        if (x >= w2) continue;
        const int64_t pix = (static_cast<int64_t>(n) * h2 + y) * w2 + x;
        // ---- logits: z2 at the pixel + bilinear(align_corners) taps of z3..z5
        float4 lg = __ldg(reinterpret_cast<const float4*>(P.lv[0].z + pix * P.zld));
        int cidx[3][4];  // pixel index inside the level (batch * h * w < 2^31)
        float cw[3][4];
#pragma unroll
        for (int l = 1; l < 4; ++l) {
            const AsfLevel& L = P.lv[l];
            const float fy = L.sy * static_cast<float>(y), fx = L.sx * static_cast<float>(x);
            const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
            const int y1 = y0 + ((y0 < L.h - 1) ? 1 : 0), x1 = x0 + ((x0 < L.w - 1) ? 1 : 0);
            const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            const int base = n * L.h * L.w;
            cidx[l - 1][0] = base + y0 * L.w + x0;  cw[l - 1][0] = hy * hx;
            cidx[l - 1][1] = base + y0 * L.w + x1;  cw[l - 1][1] = hy * lx;
            cidx[l - 1][2] = base + y1 * L.w + x0;  cw[l - 1][2] = ly * hx;
            cidx[l - 1][3] = base + y1 * L.w + x1;  cw[l - 1][3] = ly * lx;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 zc = __ldg(reinterpret_cast<const float4*>(L.z + static_cast<int64_t>(cidx[l - 1][c]) * P.zld));
                lg.x = fmaf(cw[l - 1][c], zc.x, lg.x);
                lg.y = fmaf(cw[l - 1][c], zc.y, lg.y);
                lg.z = fmaf(cw[l - 1][c], zc.z, lg.z);
                lg.w = fmaf(cw[l - 1][c], zc.w, lg.w);
            }
        }
        const float a0 = lg.x + b0, a1 = lg.y + b1, a2 = lg.z + b2, a3 = lg.w + b3;
        const float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        const float e0 = expf(a0 - mx), e1 = expf(a1 - mx), e2 = expf(a2 - mx), e3 = expf(a3 - mx);
        const float inv = 1.0f / (e0 + e1 + e2 + e3);
        const float sl[4] = {e0 * inv, e1 * inv, e2 * inv, e3 * inv};
        // ---- blend: out = s0 * p2 + sum_l s_l * sum_corner w_corner * p_l[corner]
        float acc[4][8];
        {
            const __nv_bfloat16* src = P.lv[0].p + pix * 256 + cg * 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
                unpack8(ldg128(src + k * 64), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[k][j] = sl[0] * f[j];
            }
        }
#pragma unroll
        for (int l = 1; l < 4; ++l) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float wgt = sl[l] * cw[l - 1][c];
                const __nv_bfloat16* src = P.lv[l].p + static_cast<int64_t>(cidx[l - 1][c]) * 256 + cg * 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float f[8];
                    unpack8(ldg128(src + k * 64), f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(wgt, f[j], acc[k][j]);
                }
            }
        }
        __nv_bfloat16* dst = P.out + pix * 256 + cg * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint4 u;
            u.x = pack_bf16(acc[k][0], acc[k][1]);
            u.y = pack_bf16(acc[k][2], acc[k][3]);
            u.z = pack_bf16(acc[k][4], acc[k][5]);
            u.w = pack_bf16(acc[k][6], acc[k][7]);
            *reinterpret_cast<uint4*>(dst + k * 64) = u;
        }
    }
}

// ------------------------------------------------------------------------------------------ image resize
// cv2.resize(img, (dst_w, dst_h)) for 8-bit images with the default INTER_LINEAR (the reference's resize_image_for_det,
// src/pipeline/pipeline2.py:34-40, and resize_image, src/det/test.py:108-121), bit for bit: OpenCV's fixed-point path
// (imgproc/resize.cpp: 11-bit coefficients, horizontal pass in int, vertical pass ((b * (S >> 4)) >> 16), + 2 >> 2).
// x: fx = float((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx, clamped at both borders (fx = 0);
// y: NOT clamped -- the two source rows are clipped into the image instead, so a border row is blended with itself by
// two separately truncated products.
__global__ void resize_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sh, int sw, int dh, int dw,
                                 int src_pitch, int dst_pitch, double scale_x, double scale_y) {
    const int64_t total = static_cast<int64_t>(dh) * dw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int dx = static_cast<int>(i % dw), dy = static_cast<int>(i / dw);
        int x0, x1, a0, a1, y0, y1, b0, b1;
        resize_coef(dx, scale_x, sw, true, x0, x1, a0, a1);
        resize_coef(dy, scale_y, sh, false, y0, y1, b0, b1);
        const uint8_t* r0 = src + static_cast<int64_t>(y0) * src_pitch;
        const uint8_t* r1 = src + static_cast<int64_t>(y1) * src_pitch;
        uint8_t* o = dst + static_cast<int64_t>(dy) * dst_pitch + dx * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int s0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
            const int s1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
            const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
            o[c] = static_cast<uint8_t>(min(max(v, 0), 255));
        }
    }
}

// ------------------------------------------------------------------------------------------ DB head tail
struct HeadParams {
    const __nv_bfloat16* t_bin;
    const __nv_bfloat16* t_thr;
    const float* w2_bin;
    const float* w2_thr;
    float b2_bin, b2_thr, k;
    float* binary;
    float* thresh;
    float* thresh_binary;
    float* bin_logits;
    float* thresh_logits;
    int batch, h, w, ld;
};

constexpr int kHeadPix = 128;       // pixels (threads) per block
constexpr int kHeadRowBytes = 144;  // 128 B of channels + 16 B pad: conflict-free 16-byte row reads

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Last ConvTranspose2d(64 -> 1, k2 s2) of both branches + sigmoid + DB step (model/det/head.py:15,23,28-48).
// Persistent blocks: while one 128-pixel tile (2 x 128 x 64 channels) is reduced from shared memory -- one thread per
// input pixel, its 2x2 output patch of all requested maps, 8-byte coalesced stores -- the next tile streams in with
// cp.async (coalesced 16-byte copies, no register staging).
__global__ void __launch_bounds__(kHeadPix) db_head_kernel(const HeadParams p) {
    extern __shared__ __align__(16) uint8_t sb_raw[];  // [2 buffers][2 branches][kHeadPix * kHeadRowBytes]
    __shared__ float sw[2][256];
    for (int t = threadIdx.x; t < 256; t += kHeadPix) {
        sw[0][t] = p.w2_bin[t];
        sw[1][t] = p.w2_thr[t];
    }
    const int64_t total = static_cast<int64_t>(p.batch) * p.h * p.w;
    const int64_t ntiles = (total + kHeadPix - 1) / kHeadPix;
    constexpr int kBranch = kHeadPix * kHeadRowBytes;
    auto fetch = [&](int64_t tile, int buf) {
        const int64_t pix0 = tile * kHeadPix;
        const int npx = static_cast<int>(min(static_cast<int64_t>(kHeadPix), total - pix0));
        uint8_t* dst = sb_raw + buf * 2 * kBranch;
        for (int t = threadIdx.x; t < kHeadPix * 8; t += kHeadPix) {  // 128 pixels x 8 chunks of 16 B per branch
            const int px = t >> 3, ch = t & 7;
            if (px < npx) {
                cp_async16(dst + px * kHeadRowBytes + ch * 16, p.t_bin + (pix0 + px) * p.ld + ch * 8);
                cp_async16(dst + kBranch + px * kHeadRowBytes + ch * 16, p.t_thr + (pix0 + px) * p.ld + ch * 8);
            }
        }
        cp_async_commit();
    };
    int buf = 0;
    if (static_cast<int64_t>(blockIdx.x) < ntiles) fetch(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int64_t next = tile + gridDim.x;
        if (next < ntiles) {
            fetch(next, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int64_t pix0 = tile * kHeadPix;
        const int px = threadIdx.x;
        if (pix0 + px < total) {
            const uint8_t* src = sb_raw + buf * 2 * kBranch;
            float acc[2][4];
#pragma unroll
            for (int br = 0; br < 2; ++br) {
                acc[br][0] = acc[br][1] = acc[br][2] = acc[br][3] = 0.0f;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    float f[8];
                    unpack8(*reinterpret_cast<const uint4*>(src + br * kBranch + px * kHeadRowBytes + ch * 16), f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 wv = *reinterpret_cast<const float4*>(&sw[br][(ch * 8 + j) * 4]);
                        acc[br][0] = fmaf(f[j], wv.x, acc[br][0]);
                        acc[br][1] = fmaf(f[j], wv.y, acc[br][1]);
                        acc[br][2] = fmaf(f[j], wv.z, acc[br][2]);
                        acc[br][3] = fmaf(f[j], wv.w, acc[br][3]);
                    }
                }
            }
            const int64_t pix = pix0 + px;
            const int x = static_cast<int>(pix % p.w);
            const int64_t r = pix / p.w;
            const int y = static_cast<int>(r % p.h);
            const int n = static_cast<int>(r / p.h);
            const int ow = 2 * p.w;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const float bl0 = acc[0][dy * 2 + 0] + p.b2_bin, bl1 = acc[0][dy * 2 + 1] + p.b2_bin;
                const float tl0 = acc[1][dy * 2 + 0] + p.b2_thr, tl1 = acc[1][dy * 2 + 1] + p.b2_thr;
                const float bp0 = 1.0f / (1.0f + expf(-bl0)), bp1 = 1.0f / (1.0f + expf(-bl1));
                const float tp0 = 1.0f / (1.0f + expf(-tl0)), tp1 = 1.0f / (1.0f + expf(-tl1));
                const int64_t o = (static_cast<int64_t>(n) * (2 * p.h) + (2 * y + dy)) * ow + 2 * x;
                if (p.binary) *reinterpret_cast<float2*>(p.binary + o) = make_float2(bp0, bp1);
                if (p.thresh) *reinterpret_cast<float2*>(p.thresh + o) = make_float2(tp0, tp1);
                if (p.thresh_binary) {
                    const float tb0 = 1.0f / (1.0f + expf(-p.k * (bp0 - tp0))), tb1 = 1.0f / (1.0f + expf(-p.k * (bp1 - tp1)));
                    *reinterpret_cast<float2*>(p.thresh_binary + o) = make_float2(tb0, tb1);
                }
                if (p.bin_logits) *reinterpret_cast<float2*>(p.bin_logits + o) = make_float2(bl0, bl1);
                if (p.thresh_logits) *reinterpret_cast<float2*>(p.thresh_logits + o) = make_float2(tl0, tl1);
            }
        }
        __syncthreads();  // everyone is done with `buf` before the next iteration's fetch overwrites it
    }
}

// sigmoid + DB step from the logit maps written by the B200_EPI_HEAD_DOT epilogue; same expressions as db_head_kernel.
__global__ void db_maps_kernel(const float4* __restrict__ bl, const float4* __restrict__ tl, float k, float4* binary,
                               float4* thresh, float4* thresh_binary, float4* bl_out, float4* tl_out, int64_t n4) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 b = __ldg(bl + i), t = __ldg(tl + i);
        const float bv[4] = {b.x, b.y, b.z, b.w}, tv[4] = {t.x, t.y, t.z, t.w};
        float bp[4], tp[4], tb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bp[j] = 1.0f / (1.0f + expf(-bv[j]));
            tp[j] = 1.0f / (1.0f + expf(-tv[j]));
            tb[j] = 1.0f / (1.0f + expf(-k * (bp[j] - tp[j])));
        }
        if (binary) binary[i] = make_float4(bp[0], bp[1], bp[2], bp[3]);
        if (thresh) thresh[i] = make_float4(tp[0], tp[1], tp[2], tp[3]);
        if (thresh_binary) thresh_binary[i] = make_float4(tb[0], tb[1], tb[2], tb[3]);
        if (bl_out) bl_out[i] = b;
        if (tl_out) tl_out[i] = t;
    }
}

static int grid_for(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_stem_pack(const float* img, void* packed, int batch, int h, int w, void* stream) {
    if (!img || !packed) return set_error("b200_stem_pack: null pointer");
    if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return set_error("b200_stem_pack: H and W must be even");
    const int64_t total = static_cast<int64_t>(batch) * (h / 2 + 3) * (w / 2 + 3);
    if (total < (int64_t(1) << 30))
        stem_pack_kernel<uint32_t><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            img, reinterpret_cast<uint4*>(packed), batch, h, w);
    else
        stem_pack_kernel<int64_t><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            img, reinterpret_cast<uint4*>(packed), batch, h, w);
    return check_cuda(cudaGetLastError(), "b200_stem_pack launch");
}

extern "C" int b200_stem_pack_u8(const void* img_hwc, void* packed, int batch, int h, int w, const float* mean3,
                                 const float* std3, void* stream) {
    if (!img_hwc || !packed || !mean3 || !std3) return set_error("b200_stem_pack_u8: null pointer");
    if (batch <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return set_error("b200_stem_pack_u8: H and W must be even");
    if (std3[0] == 0.0f || std3[1] == 0.0f || std3[2] == 0.0f) return set_error("b200_stem_pack_u8: zero std");
    const int64_t total = static_cast<int64_t>(batch) * (h / 2 + 3) * (w / 2 + 3);
    if (total < (int64_t(1) << 30))
        stem_pack_u8_kernel<uint32_t><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const uint8_t*>(img_hwc), reinterpret_cast<uint4*>(packed), batch, h, w, mean3[0], mean3[1],
            mean3[2], 1.0f / std3[0], 1.0f / std3[1], 1.0f / std3[2]);
    else
        stem_pack_u8_kernel<int64_t><<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const uint8_t*>(img_hwc), reinterpret_cast<uint4*>(packed), batch, h, w, mean3[0], mean3[1],
            mean3[2], 1.0f / std3[0], 1.0f / std3[1], 1.0f / std3[2]);
    return check_cuda(cudaGetLastError(), "b200_stem_pack_u8 launch");
}

extern "C" int b200_resize_u8_linear(const void* src_hwc, int src_h, int src_w, int src_pitch, void* dst_hwc, int dst_h,
                                     int dst_w, int dst_pitch, void* stream) {
    if (!src_hwc || !dst_hwc) return set_error("b200_resize_u8_linear: null pointer");
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || src_pitch < 3 * src_w || dst_pitch < 3 * dst_w)
        return set_error("b200_resize_u8_linear: bad shape");
    // cv::resize: inv_scale = dsize / (double)ssize; scale = 1. / inv_scale
    const double scale_x = 1.0 / (dst_w / static_cast<double>(src_w)), scale_y = 1.0 / (dst_h / static_cast<double>(src_h));
    const int64_t total = static_cast<int64_t>(dst_h) * dst_w;
    resize_u8_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint8_t*>(src_hwc), reinterpret_cast<uint8_t*>(dst_hwc), src_h, src_w, dst_h, dst_w, src_pitch,
        dst_pitch, scale_x, scale_y);
    return check_cuda(cudaGetLastError(), "b200_resize_u8_linear launch");
}

extern "C" int b200_db_maps(const float* bin_logits, const float* thresh_logits, float k, float* binary, float* thresh,
                            float* thresh_binary, float* bin_logits_out, float* thresh_logits_out, int64_t n, void* stream) {
    if (!bin_logits || !thresh_logits) return set_error("b200_db_maps: null pointer");
    if (n <= 0 || (n & 3)) return set_error("b200_db_maps: n must be a positive multiple of 4");
    db_maps_kernel<<<grid_for(n / 4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(bin_logits), reinterpret_cast<const float4*>(thresh_logits), k,
        reinterpret_cast<float4*>(binary), reinterpret_cast<float4*>(thresh), reinterpret_cast<float4*>(thresh_binary),
        reinterpret_cast<float4*>(bin_logits_out), reinterpret_cast<float4*>(thresh_logits_out), n / 4);
    return check_cuda(cudaGetLastError(), "b200_db_maps launch");
}

extern "C" int b200_maxpool3x3s2(const void* x, void* y, int batch, int h, int w, int c, void* stream) {
    if (!x || !y) return set_error("b200_maxpool3x3s2: null pointer");
    if (batch <= 0 || (h & 1) || (w & 1) || (c % 8)) return set_error("b200_maxpool3x3s2: need even H, W and C %% 8 == 0");
    const int c8 = c / 8;
    if (c8 <= 256 && (c8 & (c8 - 1)) == 0 && batch <= 65535 && !getenv("B200_MAXPOOL_GENERIC")) {  // strip kernel
        constexpr int kRows = 4;
        int shift = 0;
        while ((1 << shift) < c8) ++shift;
        const int px = 256 >> shift;
        const dim3 grid((w / 2 + px - 1) / px, (h / 2 + kRows - 1) / kRows, batch);
        if (grid.y <= 65535) {
            maxpool_strip_kernel<kRows><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
                reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), h, w, c, shift);
            return check_cuda(cudaGetLastError(), "b200_maxpool3x3s2 launch");
        }
    }
    const int64_t total = static_cast<int64_t>(batch) * (h / 2) * (w / 2) * (c / 8);
    maxpool_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), batch, h, w, c);
    return check_cuda(cudaGetLastError(), "b200_maxpool3x3s2 launch");
}

extern "C" int b200_asf_fwd(const void* p2, const void* p3, const void* p4, const void* p5, const float* w_att,
                            const float* b_att, void* out, int batch, int h2, int w2, int h3, int w3, int h4, int w4,
                            int h5, int w5, void* stream) {
    if (!p2 || !p3 || !p4 || !p5 || !w_att || !b_att || !out) return set_error("b200_asf_fwd: null pointer");
    if (batch <= 0 || h2 <= 0 || w2 <= 0) return set_error("b200_asf_fwd: bad shape");
    const int64_t ntiles = static_cast<int64_t>(batch) * ((h2 + 7) / 8) * ((w2 + 15) / 16);
    asf_kernel<<<grid_for(ntiles, 1), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(p2), reinterpret_cast<const __nv_bfloat16*>(p3),
        reinterpret_cast<const __nv_bfloat16*>(p4), reinterpret_cast<const __nv_bfloat16*>(p5), w_att, b_att,
        reinterpret_cast<__nv_bfloat16*>(out), batch, h2, w2, h3, w3, h4, w4, h5, w5);
    return check_cuda(cudaGetLastError(), "b200_asf_fwd launch");
}

extern "C" int b200_asf_combine(const void* p2, const void* p3, const void* p4, const void* p5, const float* z2,
                                const float* z3, const float* z4, const float* z5, int zld, const float* b_att, void* out,
                                int batch, int h2, int w2, int h3, int w3, int h4, int w4, int h5, int w5, void* stream) {
    if (!p2 || !p3 || !p4 || !p5 || !z2 || !z3 || !z4 || !z5 || !b_att || !out) return set_error("b200_asf_combine: null pointer");
    if (batch <= 0 || h2 <= 0 || w2 <= 0 || zld < 4 || (zld % 4)) return set_error("b200_asf_combine: bad shape");
    AsfParams P;
    const void* ps[4] = {p2, p3, p4, p5};
    const float* zs[4] = {z2, z3, z4, z5};
    const int hs[4] = {h2, h3, h4, h5}, ws[4] = {w2, w3, w4, w5};
    for (int l = 0; l < 4; ++l) {
        P.lv[l].p = reinterpret_cast<const __nv_bfloat16*>(ps[l]);
        P.lv[l].z = zs[l];
        P.lv[l].h = hs[l];
        P.lv[l].w = ws[l];
        P.lv[l].sy = (h2 > 1) ? static_cast<float>(hs[l] - 1) / static_cast<float>(h2 - 1) : 0.0f;
        P.lv[l].sx = (w2 > 1) ? static_cast<float>(ws[l] - 1) / static_cast<float>(w2 - 1) : 0.0f;
    }
    P.b_att = b_att;
    P.out = reinterpret_cast<__nv_bfloat16*>(out);
    P.batch = batch;
    P.zld = zld;
    const int64_t nquads = static_cast<int64_t>(batch) * h2 * ((w2 + 3) / 4);
    asf_combine_kernel<<<grid_for(nquads, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P);
    return check_cuda(cudaGetLastError(), "b200_asf_combine launch");
}

extern "C" int b200_db_head_fwd(const void* t_bin, const void* t_thr, const float* w2_bin, float b2_bin,
                                const float* w2_thr, float b2_thr, float k, float* binary, float* thresh,
                                float* thresh_binary, float* bin_logits, float* thresh_logits, int batch, int h, int w,
                                int ld, void* stream) {
    if (!t_bin || !t_thr || !w2_bin || !w2_thr) return set_error("b200_db_head_fwd: null pointer");
    if (batch <= 0 || h <= 0 || w <= 0 || ld < 64 || (ld % 8)) return set_error("b200_db_head_fwd: bad shape");
    HeadParams p;
    p.t_bin = reinterpret_cast<const __nv_bfloat16*>(t_bin);
    p.t_thr = reinterpret_cast<const __nv_bfloat16*>(t_thr);
    p.w2_bin = w2_bin; p.w2_thr = w2_thr; p.b2_bin = b2_bin; p.b2_thr = b2_thr; p.k = k;
    p.binary = binary; p.thresh = thresh; p.thresh_binary = thresh_binary;
    p.bin_logits = bin_logits; p.thresh_logits = thresh_logits;
    p.batch = batch; p.h = h; p.w = w; p.ld = ld;
    const int64_t total = static_cast<int64_t>(batch) * h * w;
    const int64_t ntiles = (total + kHeadPix - 1) / kHeadPix;
    const int smem = 4 * kHeadPix * kHeadRowBytes;  // 2 buffers x 2 branches = 72 KB -> 3 blocks per SM
    cudaError_t e = cudaFuncSetAttribute(db_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(e, "b200_db_head_fwd attribute");
    const int blocks = static_cast<int>(ntiles < 148 * 3 ? ntiles : 148 * 3);
    db_head_kernel<<<blocks, kHeadPix, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return check_cuda(cudaGetLastError(), "b200_db_head_fwd launch");
}
