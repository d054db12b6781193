// cv2.resize (8-bit, INTER_LINEAR) coefficient set-up shared by the whole-image resize (elementwise.cu) and the
// recognition crop kernel (recprep.cu).  OpenCV imgproc/resize.cpp restated: see oracle/preprocess_oracle.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// destination index d -> the two source indices (clipped into the image) and their 11-bit weights.
// clamp_frac = true is the x rule (fraction forced to 0 at both borders), false the y rule (rows clipped only).
__device__ __forceinline__ void resize_coef(int d, double scale, int src, bool clamp_frac, int& s0, int& s1, int& a0, int& a1) {
    // explicit round-to-nearest multiply and add: no FMA contraction (OpenCV's host code rounds the product first)
    float f = static_cast<float>(__dadd_rn(__dmul_rn(d + 0.5, scale), -0.5));
    int s = static_cast<int>(floorf(f));
    f -= static_cast<float>(s);
    if (clamp_frac) {
        if (s < 0) { f = 0.0f; s = 0; }
        if (s >= src - 1) { f = 0.0f; s = src - 1; }
    }
    a0 = __float2int_rn((1.0f - f) * 2048.0f);  // saturate_cast<short>(float) = round half to even
    a1 = __float2int_rn(f * 2048.0f);
    s0 = min(max(s, 0), src - 1);
    s1 = min(max(s + 1, 0), src - 1);
}

// one output byte from the four source bytes: horizontal pass in int, vertical pass with OpenCV's truncations
__device__ __forceinline__ int resize_blend(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
    const int s0 = p00 * a0 + p01 * a1;
    const int s1 = p10 * a0 + p11 * a1;
    const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
    return min(max(v, 0), 255);
}

}  // namespace b200
