// Host-side helpers shared by the translation units of libdbnet_b200.so.
#pragma once
#include "../../include/dbnet_b200.h"

namespace b200 {
// printf-style; stores the message for b200_last_error() (thread local) and returns 1.
int set_error(const char* fmt, ...);
// returns 0 when `e` is cudaSuccess, else records "<what>: <cuda error string>" and returns 1.
int check_cuda(int e, const char* what);
}  // namespace b200
