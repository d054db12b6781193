// Development probe: shared-memory throughput of 128-bit stores for the address patterns igemm.cu uses.
#include "common.cuh"
#include <cstdio>
using namespace b200;
__global__ void __launch_bounds__(512, 1) sts_kernel(int iters, int pattern, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t off;
    const int sub = lane >> 3, l8 = lane & 7;
    switch (pattern) {
        case 0: off = warp * 512 + lane * 16; break;                                        // contiguous
        case 1: { const int r = warp * 4 + sub; off = r * 128 + ((l8 ^ (r & 7)) << 4); } break;  // DCN producer (swizzled row per quarter-warp)
        case 2: { const int r = (warp & 3) * 32 + lane; off = r * 128 + (((warp >> 2) ^ (r & 7)) << 4); } break;  // epilogue (lane = row)
        case 3: { const int r = warp * 4 + sub; off = r * 128 + (l8 << 4); } break;         // row per quarter-warp, no swizzle
        case 4: { const int r = warp * 4 + (lane & 3); const int c = lane >> 2; off = r * 128 + ((c ^ (r & 7)) << 4); } break; // rows interleaved
        default: { const int r = warp * 4 + sub; off = r * 128 + ((((l8 + 2 * sub) & 7) ^ (r & 7)) << 4); } break;  // rotated chunks per row
    }
    const uint4 v = make_uint4(lane, warp, 3, 4);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        sts128(base + off, v);
        sts128(base + 16384 + off, v);
        sts128(base + 32768 + off, v);
        sts128(base + 49152 + off, v);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(sts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int iters = 2000;
    for (int p = 0; p < 6; ++p) {
        sts_kernel<<<148, 512, 80 * 1024>>>(iters, p, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("pattern %d: %.2f cycles per warp-store (16 warps), %.1f B/clk  (%s)\n", p, (double)h / (iters * 4.0 * 16), 512.0 * iters * 4 * 16 / h, cudaGetErrorString(e));
    }
    return 0;
}
