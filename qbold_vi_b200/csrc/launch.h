// Host-side helpers shared by the translation units of libqbold.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qbold.h"

namespace qb {

constexpr int kThreads = 256;

int fail(int code, const char* fmt, ...);          // records qbold_last_error(), returns code
int after_launch(const char* kernel_name);         // counts the launch, maps cudaGetLastError()
int cuda_check(cudaError_t e, const char* what);   // 0 or QBOLD_ECUDA
int sm_count();                                    // SMs of the current device (cached per device)

}  // namespace qb
