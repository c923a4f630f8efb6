#include "common.cuh"

#include <atomic>
#include <mutex>
#include <stdio.h>

namespace lkgd {

static char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return LKGD_ECUDA;
}

int launch_epilogue() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? LKGD_OK : set_cuda_error(e);
}

int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
              const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled entry point not available");
    return LKGD_ECUDA;
  }
  if (!aligned16(base)) return LKGD_EALIGN;
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides[i];
    if (s[i] % 16) return LKGD_EALIGN;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)",
             (int)r, rank, (unsigned long long)d[0], (unsigned long long)d[1], b[0], b[1]);
    return LKGD_ECUDA;
  }
  return LKGD_OK;
}

}  // namespace lkgd

using namespace lkgd;

extern "C" int lkgd_abi_version(void) { return LKGD_ABI_VERSION; }

extern "C" const char* lkgd_strerror(int code) {
  switch (code) {
    case LKGD_OK: return "ok";
    case LKGD_ESHAPE: return "unsupported or inconsistent shape";
    case LKGD_EALIGN: return "pointer or pitch not 16-byte aligned";
    case LKGD_EARCH: return "device is not sm_100";
    case LKGD_EWS: return "workspace too small";
    case LKGD_ECUDA: return "CUDA error";
    default: return "unknown error";
  }
}

extern "C" const char* lkgd_last_cuda_error(void) { return g_err; }

extern "C" int lkgd_device_check(int dev) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_cuda_error(e);
  return major == 10 ? LKGD_OK : LKGD_EARCH;
}

extern "C" uint64_t lkgd_launch_count(void) { return g_launches.load(); }
