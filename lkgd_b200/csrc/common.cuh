// Shared host-side helpers: error plumbing, launch counter, SM count, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/lkgd_b200.h"

namespace lkgd {

int set_cuda_error(cudaError_t e);          // records the message, returns LKGD_ECUDA
int launch_epilogue();                      // counts the launch, checks cudaGetLastError()
int sm_count();
// bf16 tiled tensor map, SWIZZLE_128B, zero OOB fill. dims/box innermost first; strides in bytes for dims 1..rank-1.
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
              const uint32_t* box);

// One-time per-DEVICE set-up (cudaFuncSetAttribute is per device): true the first time it is called for the calling
// thread's current device with this flag word (one bit per device ordinal).
struct DeviceOnce {
  unsigned long long bits = 0;
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return !(__atomic_fetch_or(&bits, bit, __ATOMIC_ACQ_REL) & bit);
  }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace lkgd
