// Status / error plumbing of the C-ABI (include/dbnet_b200.h).  Nothing here throws or aborts:
// the reference only ever guards with try/except (src/det/test.py:87-95), so every failure comes
// back as a non-zero status plus a message.
#include "internal.h"

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

namespace b200 {

static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check_cuda(int e, const char* what) {
    if (e == cudaSuccess) return 0;
    return set_error("%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(e)));
}

}  // namespace b200

extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" const char* b200_last_error(void) { return b200::g_err; }

extern "C" int b200_device_check(int dev) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return b200::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    if (prop.major != 10) return b200::set_error("device %d is sm_%d%d, this library is sm_100a only", dev, prop.major, prop.minor);
    return 0;
}
