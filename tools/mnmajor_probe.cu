// Development probe (not part of the product): the B operand of tcgen05.mma in MN-major form with the 128-byte swizzle,
// i.e. shared-memory rows = K index (here: source pixels), 64 N elements (channels) contiguous per row -- the layout a TMA
// box of an NHWC tensor lands in.  Descriptor: LBO = bytes between consecutive 64-element N groups, SBO = bytes between
// groups of 8 K rows (1024 when rows are packed), instruction descriptor bit 16 (b_major) = 1.
// A[i][k] = (k == i % 16) (K-major, known-good path), B[n][k] = (7 n + 3 k) % 251, N = 128, K = 16: D[i][n] must be B[n][i % 16].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ocr_vi_invoice_b200/csrc mnmajor_probe.cu -o mnmajor_probe
#include "common.cuh"
#include <cstdio>

using namespace b200;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(float* out, int k_rows_per_group_stride /*SBO*/, int lbo) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = base;           // A: 128 rows x 128 B (K-major SW128), only K = 0..15 used
    uint8_t* b_s = base + 16384;   // B: n-group g at g * lbo; inside: row k (128 B) at (k / 8) * SBO + (k % 8) * 128
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    __syncthreads();
    {
        const int i = threadIdx.x;  // A row
        for (int k = 0; k < 16; ++k) {
            const int piece = (k >> 3) ^ (i & 7);
            reinterpret_cast<__nv_bfloat16*>(a_s + i * 128 + piece * 16)[k & 7] = __float2bfloat16((k == (i & 15)) ? 1.0f : 0.0f);
        }
        const int n = threadIdx.x;  // B column (channel)
        for (int k = 0; k < 16; ++k) {
            const int g = n >> 6, nn = n & 63;
            const int piece = (nn >> 3) ^ (k & 7);
            uint8_t* row = b_s + g * lbo + (k >> 3) * k_rows_per_group_stride + (k & 7) * 128;
            reinterpret_cast<__nv_bfloat16*>(row + piece * 16)[nn & 7] = __float2bfloat16(static_cast<float>((7 * n + 3 * k) % 251));
        }
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
            const uint32_t idesc = umma_idesc_bf16(128u) | (1u << 16);  // b_major = MN
            umma_bf16(tmem, umma_desc_sw128(smem_u32(a_s)), desc_mn_sw128(smem_u32(b_s), lbo, k_rows_per_group_stride), idesc, 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 128 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    static float h[128 * 128];
    int rc = 0;
    const int cfg[3][2] = {{1024, 2048}, {1024, 4096}, {2048, 1024}};  // (SBO, LBO): packed; padded N groups; interleaved
    for (auto& c : cfg) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 40 * 1024>>>(d, c[0], c[1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 128; ++i)
            for (int n = 0; n < 128; ++n) {
                const float want = static_cast<float>((7 * n + 3 * (i & 15)) % 251);
                if (h[i * 128 + n] != want) {
                    if (bad < 6) printf("  D[%d][%d] = %g, expected %g\n", i, n, h[i * 128 + n], want);
                    ++bad;
                }
            }
        printf("MN-major B, SBO %d LBO %d: %s (%d mismatches)\n", c[0], c[1], bad ? "UNEXPECTED" : "as expected", bad);
        rc |= bad ? 2 : 0;
    }
    return rc;
}
