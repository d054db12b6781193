// Greedy CTC decode on the GPU (the tail of the recognition path).
//
// Reference: SVTRv2.decode_probs, model/rec2/svtrv2.py:550-569 -- `log_probs.argmax(dim=-1)` over (T, B, C), then a
// Python loop per sequence that keeps frame t's class p iff p != blank and p != the class of frame t-1.  Restated in
// oracle/ctc_oracle.py and pinned to the reference's own method by tests/golden/ctc_reference.npz.  Index work:
// bit-exact, including torch.argmax's conventions (first index of the maximum; a NaN is the maximum).
//
// HBM-bound: the (T, B, C) fp32 tensor is read exactly once (lane-strided, coalesced rows); one block per sequence, one
// warp per frame for the argmax, then a ballot / prefix compaction of the kept frames.  Only B * (T + 1) int32 go back.
#include "internal.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b200 {

constexpr int kCtcThreads = 256;

struct Best {
    float v;
    int i;
    bool nan;
};

// does a beat b under torch.argmax's order?  NaN above everything, then larger value, then smaller index
__device__ __forceinline__ bool beats(const Best& a, const Best& b) {
    if (a.nan != b.nan) return a.nan;
    if (a.nan) return a.i < b.i;
    return a.v > b.v || (a.v == b.v && a.i < b.i);
}

__global__ void __launch_bounds__(kCtcThreads) ctc_greedy_kernel(const float* __restrict__ lp, int T, int B, int C, int64_t stride_t,
                                                                 int64_t stride_b, int blank, int32_t* __restrict__ tokens,
                                                                 int32_t* __restrict__ lengths) {
    extern __shared__ int s_ids[];  // [T]
    __shared__ int s_warp_count[kCtcThreads / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kCtcThreads / 32;

    for (int t = warp; t < T; t += kWarps) {
        const float* row = lp + t * stride_t + b * stride_b;
        Best best{-INFINITY, 0x7fffffff, false};
        for (int c0 = 0; c0 < C; c0 += 256) {  // eight independent loads in flight per lane, then the compare chain
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c0 + k * 32 + lane;
                v[k] = c < C ? __ldg(row + c) : -INFINITY;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c0 + k * 32 + lane;
                const Best cand{v[k], c < C ? c : 0x7fffffff, isnan(v[k])};
                if (beats(cand, best)) best = cand;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Best o;
            o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
            o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
            o.nan = __shfl_xor_sync(0xffffffffu, static_cast<int>(best.nan), off) != 0;
            if (beats(o, best)) best = o;
        }
        if (lane == 0) s_ids[t] = best.i;  // C >= 1: some lane saw index 0, so best.i is a real class (all -inf -> 0)
    }
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();

    int32_t* out = tokens + static_cast<int64_t>(b) * T;
    for (int t0 = 0; t0 < T; t0 += kCtcThreads) {
        const int t = t0 + threadIdx.x;
        int id = -1;
        bool keep = false;
        if (t < T) {
            id = s_ids[t];
            keep = id != blank && (t == 0 || id != s_ids[t - 1]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp_count[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp_count[w];
        if (keep) out[before + __popc(m & ((1u << lane) - 1u))] = id;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += s_warp_count[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int n = s_base;
    for (int t = n + threadIdx.x; t < T; t += kCtcThreads) out[t] = -1;
    if (threadIdx.x == 0) lengths[b] = n;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_ctc_greedy(const float* log_probs, int T, int B, int C, int64_t stride_t, int64_t stride_b, int blank_id,
                               int32_t* tokens, int32_t* lengths, void* stream) {
    if (!log_probs || !tokens || !lengths) return set_error("b200_ctc_greedy: null pointer");
    if (T <= 0 || B <= 0 || C <= 0 || T > 8192) return set_error("b200_ctc_greedy: need 1 <= T <= 8192, B >= 1, C >= 1");
    if (stride_t < 0 || stride_b < 0) return set_error("b200_ctc_greedy: negative stride");
    ctc_greedy_kernel<<<B, kCtcThreads, T * sizeof(int), reinterpret_cast<cudaStream_t>(stream)>>>(
        log_probs, T, B, C, stride_t, stride_b, blank_id, tokens, lengths);
    return check_cuda(cudaGetLastError(), "b200_ctc_greedy launch");
}
