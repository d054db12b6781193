// Development probe (not part of the product): K-major UMMA operands WITHOUT swizzle (descriptor layout type 0), the
// compact form for a single K = 16 step: core matrix = 8 rows x 16 bytes (128 contiguous bytes); element (r, k) of an
// operand lives at (r % 8) * 16 + (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 2.
// A[i][k] = (i + 3 k) % 7, B[n][k] = (5 n + k) % 11 (exact in bf16), M = 128, N = 64, K = 16: D[i][n] = sum_k A[i][k] B[n][k].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ocr_vi_invoice_b200/csrc noswizzle_probe.cu -o noswizzle_probe
#include "common.cuh"
#include <cstdio>

using namespace b200;

__device__ __forceinline__ uint64_t desc_k_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1, layout type 0 = no swizzle
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(float* out, int a_lbo, int a_sbo, int b_lbo, int b_sbo) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = base;
    uint8_t* b_s = base + 8192;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    __syncthreads();
    {
        const int i = threadIdx.x;
        for (int k = 0; k < 16; ++k)
            *reinterpret_cast<__nv_bfloat16*>(a_s + (i % 8) * 16 + (i / 8) * a_sbo + (k / 8) * a_lbo + (k % 8) * 2) =
                __float2bfloat16(static_cast<float>((i + 3 * k) % 7));
        if (i < 64)
            for (int k = 0; k < 16; ++k)
                *reinterpret_cast<__nv_bfloat16*>(b_s + (i % 8) * 16 + (i / 8) * b_sbo + (k / 8) * b_lbo + (k % 8) * 2) =
                    __float2bfloat16(static_cast<float>((5 * i + k) % 11));
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            umma_bf16(tmem, desc_k_noswz(smem_u32(a_s), a_lbo, a_sbo), desc_k_noswz(smem_u32(b_s), b_lbo, b_sbo), umma_idesc_bf16(64u), 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    static float h[128 * 64];
    int rc = 0;
    // (A LBO, A SBO, B LBO, B SBO): K halves adjacent (LBO 128, SBO 256) or row groups adjacent (SBO 128, LBO 8 KB-ish)
    const int cfg[2][4] = {{128, 256, 128, 256}, {2048, 128, 1024, 128}};
    for (auto& c : cfg) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 32 * 1024>>>(d, c[0], c[1], c[2], c[3]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 128; ++i)
            for (int n = 0; n < 64; ++n) {
                float want = 0;
                for (int k = 0; k < 16; ++k) want += static_cast<float>(((i + 3 * k) % 7) * ((5 * n + k) % 11));
                if (h[i * 64 + n] != want) {
                    if (bad < 6) printf("  D[%d][%d] = %g, expected %g\n", i, n, h[i * 64 + n], want);
                    ++bad;
                }
            }
        printf("no-swizzle K-major, A (LBO %d, SBO %d) B (LBO %d, SBO %d): %s (%d mismatches)\n", c[0], c[1], c[2], c[3],
               bad ? "UNEXPECTED" : "as expected", bad);
        rc |= bad ? 2 : 0;
    }
    return rc;
}
