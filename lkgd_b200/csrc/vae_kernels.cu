// HBM-bound pieces of the VAE either side of the denoise loop: the row softmax of the single-head,
// 512-wide mid-block attention (computed as Q K^T GEMM -> softmax -> P V GEMM: a 512-wide head does not fit the TMEM
// budget of the flash kernel) and the decoder's last layer, the (3,1,1) `time_conv_out` over frames fused with the
// channels-last -> NCHW unpack.  See include/lkgd_b200.h for the contract.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// One 256-thread block per row; the row (N <= 256 * 4 * SM_V floats) is read ONCE into registers.
constexpr int SM_V = 16;
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, long long ldx, int N, float scale_log2,
                                                           __nv_bfloat16* __restrict__ out, long long ldo) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int n4 = N >> 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4 v[SM_V];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SM_V; ++i) {
    const int j = tid + i * 256;
    if (j < n4) {
      v[i] = __ldg(xr + j);
      mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float off = mx * scale_log2;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SM_V; ++i) {
    const int j = tid + i * 256;
    if (j < n4) {
      v[i].x = exp2f(fmaf(v[i].x, scale_log2, -off));
      v[i].y = exp2f(fmaf(v[i].y, scale_log2, -off));
      v[i].z = exp2f(fmaf(v[i].z, scale_log2, -off));
      v[i].w = exp2f(fmaf(v[i].w, scale_log2, -off));
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.f / sum;
  uint2* orow = reinterpret_cast<uint2*>(out + row * ldo);
#pragma unroll
  for (int i = 0; i < SM_V; ++i) {
    const int j = tid + i * 256;
    if (j < n4) orow[j] = make_uint2(pack_bf16x2(v[i].x * inv, v[i].y * inv), pack_bf16x2(v[i].z * inv, v[i].w * inv));
  }
}

// out[b, f, co, p] = bias[co] + sum_{kt, ci} w[co, ci, kt] * x[b, f + kt - 1, p, ci]   (zero beyond the clip's ends)
template <int C>
__global__ void time_conv_out_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ out, int F, long long HW,
                                     long long total) {
  __shared__ float sw[C * C * 3 + C];
  for (int i = threadIdx.x; i < C * C * 3; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < C) sw[C * C * 3 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (b, f, p)
  if (idx >= total) return;
  const long long p = idx % HW;
  const int f = (int)((idx / HW) % F);
  float acc[C];
#pragma unroll
  for (int co = 0; co < C; ++co) acc[co] = sw[C * C * 3 + co];
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
    const int fi = f + kt - 1;
    if (fi < 0 || fi >= F) continue;
    const float* xr = x + (idx + (long long)(kt - 1) * HW) * ldx;
    float xin[C];
    if (C == 4 || (C == 3 && ldx % 4 == 0)) {       // one 16-byte load per tap when the row pitch allows
      const float4 t = __ldg(reinterpret_cast<const float4*>(xr));
      xin[0] = t.x; xin[1] = t.y; xin[2] = t.z;
      if (C == 4) xin[C - 1] = t.w;
    } else {
#pragma unroll
      for (int ci = 0; ci < C; ++ci) xin[ci] = __ldg(xr + ci);
    }
#pragma unroll
    for (int co = 0; co < C; ++co)
#pragma unroll
      for (int ci = 0; ci < C; ++ci) acc[co] = fmaf(sw[(co * C + ci) * 3 + kt], xin[ci], acc[co]);
  }
  const long long img = idx / HW;                                              // b * F + f
#pragma unroll
  for (int co = 0; co < C; ++co) out[(img * C + co) * HW + p] = acc[co];
}

}  // namespace lkgd

using namespace lkgd;

extern "C" int lkgd_softmax_rows(const float* x, int64_t ldx, int64_t M, int32_t N, float scale, void* out, int64_t ldo,
                                 void* stream) {
  if (M <= 0 || M > 0x7fffffffLL || N <= 0 || N % 4 || N > 256 * 4 * SM_V || ldx < N || ldo < N) return LKGD_ESHAPE;
  if (!aligned16(x) || ldx % 4 || (reinterpret_cast<uintptr_t>(out) & 7) || ldo % 4) return LKGD_EALIGN;
  softmax_rows_kernel<<<(unsigned)M, 256, 0, ST(stream)>>>(x, ldx, N, scale * 1.4426950408889634f,
                                                           reinterpret_cast<__nv_bfloat16*>(out), ldo);
  return launch_epilogue();
}

extern "C" int lkgd_time_conv_out(const float* x, int32_t ldx, const float* weight, const float* bias, float* out, int32_t NB,
                                  int32_t F, int64_t HW, int32_t C, void* stream) {
  if (NB <= 0 || F <= 0 || HW <= 0 || ldx < C) return LKGD_ESHAPE;
  if (!aligned16(x)) return LKGD_EALIGN;
  const long long total = (long long)NB * F * HW;
  const unsigned grid = (unsigned)((total + 255) / 256);
  switch (C) {
    case 3: time_conv_out_kernel<3><<<grid, 256, 0, ST(stream)>>>(x, ldx, weight, bias, out, F, HW, total); break;
    case 4: time_conv_out_kernel<4><<<grid, 256, 0, ST(stream)>>>(x, ldx, weight, bias, out, F, HW, total); break;
    case 1: time_conv_out_kernel<1><<<grid, 256, 0, ST(stream)>>>(x, ldx, weight, bias, out, F, HW, total); break;
    default: return LKGD_ESHAPE;
  }
  return launch_epilogue();
}
