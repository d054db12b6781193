// Development probe (not part of the product): the two TMEM facts the DCN kernel with the A operand in tensor memory
// relies on.
//   1. tcgen05.st.16x256b.x4: which (TMEM lane, column) each (thread, register) lands in, read back with
//      tcgen05.ld.32x32b (thread = lane, register = column).  Expected: thread T, register 4j + 2v + r ->
//      lane base + T/4 + 8v, column 8j + 2(T%4) + r.
//   2. tcgen05.mma kind::f16 with A read from TMEM ([a_tmem] operand): row i of A lives in lane i, K element 2c (2c+1)
//      in the low (high) half of column c.  Checked with A[i][k] = (k == i % 16), B[n][k] = 16 n + k (exact in bf16),
//      so D[i][n] must be 16 n + i % 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../ocr_vi_invoice_b200/csrc tmem_probe.cu -o tmem_probe
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

using namespace b200;

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out_st, float* out_mma) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // B[n][k] = 16 n + k for n < 16, k < 16, K-major SW128 rows of 128 bytes: 16-byte piece c of row r sits at piece c ^ (r & 7)
    for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    __syncthreads();
    if (threadIdx.x < 16) {
        const int n = threadIdx.x;
        for (int k = 0; k < 16; ++k) {
            const int piece = (k >> 3) ^ (n & 7);
            reinterpret_cast<__nv_bfloat16*>(base + n * 128 + piece * 16)[k & 7] = __float2bfloat16(static_cast<float>(16 * n + k));
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

    // ---- 1. store layout: value = (thread << 8) | register, columns 64..95
    for (int h = 0; h < 2; ++h) {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = 0x80000000u | (static_cast<uint32_t>(h) << 16) | (static_cast<uint32_t>(lane) << 8) | i;
        tmem_st_16x256b_x4(tmem + lane_base + (static_cast<uint32_t>(16 * h) << 16) + 64, r);
    }
    tmem_st_wait();
    {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + 64, v);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out_st[(warp * 32 + lane) * 32 + c] = v[c];
    }
    // ---- 2. A from TMEM: row = lane, columns 128..135 hold K = 16 bf16 (two per column)
    {
        const int i = warp * 32 + lane;
        uint32_t a[8];
        for (int c = 0; c < 8; ++c) {
            const float lo = (2 * c == (i & 15)) ? 1.0f : 0.0f, hi = (2 * c + 1 == (i & 15)) ? 1.0f : 0.0f;
            a[c] = pack_bf16(lo, hi);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem + lane_base + 128),
                     "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                     : "memory");
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            umma_bf16_ts(tmem + 0, tmem + 128, umma_desc_sw128(smem_u32(base)), umma_idesc_bf16(16u), 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + 0, v);
        tmem_ld_wait();
        for (int c = 0; c < 16; ++c) out_mma[(warp * 32 + lane) * 16 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    uint32_t* d_st;
    float* d_mma;
    cudaMalloc(&d_st, 128 * 32 * 4);
    cudaMalloc(&d_mma, 128 * 16 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    probe<<<1, 128, 8192>>>(d_st, d_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
    static uint32_t st[128 * 32];
    static float mm[128 * 16];
    cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
    cudaMemcpy(mm, d_mma, sizeof(mm), cudaMemcpyDeviceToHost);
    int bad_st = 0, bad_mma = 0;
    for (int row = 0; row < 128; ++row) {
        for (int c = 0; c < 32; ++c) {
            const uint32_t v = st[row * 32 + c];
            const int h = (v >> 16) & 1, T = (v >> 8) & 31, reg = v & 255;
            const int j = reg >> 2, vv = (reg >> 1) & 1, r = reg & 1;
            const int exp_row = (row & ~31) + 16 * h + T / 4 + 8 * vv, exp_col = 8 * j + 2 * (T % 4) + r;
            if (!(v >> 31) || exp_row != row || exp_col != c) {
                if (bad_st < 12) printf("st: lane %d col %d holds h %d thread %d reg %d (expected lane %d col %d)\n", row, c, h, T, reg, exp_row, exp_col);
                ++bad_st;
            }
        }
    }
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 16; ++n) {
            const float want = static_cast<float>(16 * n + (i & 15));
            if (mm[i * 16 + n] != want) {
                if (bad_mma < 12) printf("mma: D[%d][%d] = %g, expected %g\n", i, n, mm[i * 16 + n], want);
                ++bad_mma;
            }
        }
    printf("tcgen05.st 16x256b.x4 layout: %s (%d mismatches)\n", bad_st ? "UNEXPECTED" : "as expected", bad_st);
    printf("tcgen05.mma A-from-TMEM layout: %s (%d mismatches)\n", bad_mma ? "UNEXPECTED" : "as expected", bad_mma);
    return (bad_st || bad_mma) ? 2 : 0;
}
