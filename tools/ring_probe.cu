// Development probe (not part of the product): cycles per iteration of the producer/consumer mbarrier ring used by
// igemm_kernel, with nothing but the hand-shake in the loop.  Variants isolate what each instruction costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ocr_vi_invoice_b200/csrc ring_probe.cu -o ring_probe
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

using namespace b200;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// mode bits: 1 = consumer releases with tcgen05.commit (else mbarrier.arrive), 2 = tcgen05.fence::after_thread_sync after
// the consumer wait, 4 = producer adds expect_tx(0), 8 = lane-0-only loops (divergent) instead of converged + elect,
// 16 = consumer issues 4 dummy-free fences (nothing), 32 = spin with test_wait instead of try_wait
__global__ void __launch_bounds__(128, 1) ring_kernel(int iters, int stages, int mode, long long* out) {
    __shared__ __align__(8) uint64_t full[16], empty[16];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(&tmem_slot, 32); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const long long t0 = clock64();
    if (warp == 0 && (mode & 32)) {
        int s = 0; uint32_t ph = 0;
        bool ready = mbar_test_wait(&empty[0], 1);
        for (int i = 0; i < iters; ++i) {
            if (!ready) mbar_wait(&empty[s], ph ^ 1);
            int s2 = s + 1; uint32_t ph2 = ph;
            if (s2 == stages) { s2 = 0; ph2 ^= 1; }
            ready = mbar_test_wait(&empty[s2], ph2 ^ 1);
            if (elect_one()) mbar_arrive(&full[s]);
            __syncwarp();
            s = s2; ph = ph2;
        }
    } else if (warp == 1 && (mode & 16)) {
        int s = 0; uint32_t ph = 0;
        bool ready = mbar_test_wait(&full[0], 0);
        for (int i = 0; i < iters; ++i) {
            if (!ready) mbar_wait(&full[s], ph);
            int s2 = s + 1; uint32_t ph2 = ph;
            if (s2 == stages) { s2 = 0; ph2 ^= 1; }
            ready = mbar_test_wait(&full[s2], ph2);
            if (elect_one()) { if (mode & 1) umma_commit(&empty[s]); else mbar_arrive(&empty[s]); }
            __syncwarp();
            s = s2; ph = ph2;
        }
    } else if (warp == 0) {
        if (!(mode & 8) || lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < iters; ++i) {
                mbar_wait(&empty[s], ph ^ 1);
                if (mode & 8) {
                    if (mode & 4) mbar_arrive_expect_tx(&full[s], 0); else mbar_arrive(&full[s]);
                } else {
                    if (elect_one()) { if (mode & 4) mbar_arrive_expect_tx(&full[s], 0); else mbar_arrive(&full[s]); }
                    __syncwarp();
                }
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (!(mode & 8) || lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < iters; ++i) {
                mbar_wait(&full[s], ph);
                if (mode & 2) tc_fence_after();
                if (mode & 8) {
                    if (mode & 1) umma_commit(&empty[s]); else mbar_arrive(&empty[s]);
                } else {
                    if (elect_one()) { if (mode & 1) umma_commit(&empty[s]); else mbar_arrive(&empty[s]); }
                    __syncwarp();
                }
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_slot, 32); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 20000;
    for (int stages : {2, 4, 6}) {
        for (int mode : {0, 1, 16, 32, 48, 49}) {
            ring_kernel<<<148, 128>>>(iters, stages, mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("stages %d mode %2d: %7.1f cycles/iter  (%s)\n", stages, mode, (double)h / iters, cudaGetErrorString(e));
        }
    }
    return 0;
}
