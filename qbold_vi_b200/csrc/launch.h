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
// Device work counter for kernels whose per-voxel cost is uneven (masked volumes): one of a ring of 1024 counters on
// the current device, zeroed on `stream` ahead of the launch that will use it.  nullptr on failure.
unsigned long long* next_work_counter(cudaStream_t stream);

}  // namespace qb
