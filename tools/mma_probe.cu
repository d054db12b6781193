// Development probe (not part of the product): cycles per tcgen05.mma (M = 128, K = 16, bf16, operands in shared
// memory) as a function of N and of how many independent TMEM accumulators the instruction stream alternates between.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ocr_vi_invoice_b200/csrc mma_probe.cu -o mma_probe
#include "common.cuh"
#include <cstdio>

using namespace b200;

__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int n, int naccum, int kstep_mode, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint64_t da = umma_desc_sw128(smem_u32(base)), db = umma_desc_sw128(smem_u32(base + 16384));
        const uint32_t idesc = umma_idesc_bf16(static_cast<uint32_t>(n));
        const long long t0 = clock64();
        long long t_issue = 0;
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int idx = i * 4 + k;
                    const uint32_t d = tmem + static_cast<uint32_t>((idx % naccum) * n);
                    const int ks = kstep_mode ? k : 0;
                    umma_bf16(d, da + 2 * ks, db + 2 * ks, idesc, 1u);
                }
            }
            t_issue = clock64() - t0;
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && elect_one()) { out[0] = t1 - t0; out[1] = t_issue; }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int n : {32, 64, 128, 256}) {
        for (int na : {1, 2, 4}) {
            if (na * n > 512) continue;
            for (int km : {0, 1}) {
                mma_kernel<<<148, 128, 64 * 1024>>>(iters, n, na, km, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2] = {0, 0};
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N %3d accumulators %d kstep %d: %6.1f cycles/MMA total, %6.1f issue  (%s)\n", n, na, km,
                       (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
