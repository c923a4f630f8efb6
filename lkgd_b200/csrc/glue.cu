// Bandwidth-bound glue of the denoise step: layout packing, upsample / concat, residual injection, the fp32
// conditioning helpers and the fused CFG + Euler-Karras update.  See include/lkgd_b200.h for the contract.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == 1) return silu_f(x);
  if (act == 3) return x > 0.f ? x : 0.1f * x;
  return x;
}

// one warp per output column n; rows m in groups of 4; 128-bit loads when K % 4 == 0 and the pitches allow
__global__ void small_linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                    const float* __restrict__ b, float* __restrict__ y, int ldy, int M, int N, int K,
                                    int act_in, int act_out, int vec) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* w = W + (size_t)n * K;
  for (int m0 = 0; m0 < M; m0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
      const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll 4
      for (int k = lane; k < K / 4; k += 32) {
        const float4 wv = __ldg(w4 + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (m0 + i < M) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)(m0 + i) * ldx) + k);
            acc[i] = fmaf(act_apply(xv.x, act_in), wv.x, acc[i]);
            acc[i] = fmaf(act_apply(xv.y, act_in), wv.y, acc[i]);
            acc[i] = fmaf(act_apply(xv.z, act_in), wv.z, acc[i]);
            acc[i] = fmaf(act_apply(xv.w, act_in), wv.w, acc[i]);
          }
      }
    } else if ((K & 3) == 0) {
      // operands that are not 16-byte aligned (views into a flat parameter buffer): scalar loads, but the SAME
      // lane / accumulation order as the vector path - the result must not depend on where a tensor happens to live
      for (int k = lane; k < K / 4; k += 32) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float wv = __ldg(w + 4 * k + e);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (m0 + i < M) acc[i] = fmaf(act_apply(__ldg(x + (size_t)(m0 + i) * ldx + 4 * k + e), act_in), wv, acc[i]);
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (m0 + i < M) acc[i] = fmaf(act_apply(__ldg(x + (size_t)(m0 + i) * ldx + k), act_in), wv, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m0 + i < M) y[(size_t)(m0 + i) * ldy + n] = act_apply(acc[i] + (b ? b[n] : 0.f), act_out);
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int M, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * half) return;
  const int m = idx / half, k = idx % half;
  const float freq = expf(-logf(10000.0f) * (float)k / (float)half);
  const float arg = t[m] * freq;
  out[(size_t)m * dim + k] = cosf(arg);          // flip_sin_to_cos=True: [cos | sin]
  out[(size_t)m * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[(size_t)m * dim + dim - 1] = 0.f;
}

// out[n,f,h,w,c] (bf16, Cpad channels) from NCHW-per-frame fp32 sources; thread = one pixel, 8 channels per store
__global__ void pack_input_kernel(const float* __restrict__ s0, int N0, int C0, float scale0,
                                  const float* __restrict__ scale0_dev, const float* __restrict__ s1, int N1, int C1,
                                  __nv_bfloat16* __restrict__ out, int N, int F, int HW, int Cpad) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * F * HW;
  if (pix >= total) return;
  if (scale0_dev != nullptr) scale0 = __ldg(scale0_dev);     // per-step scalar of a replayed CUDA graph
  const int p = (int)(pix % HW);
  const int f = (int)((pix / HW) % F);
  const int n = (int)(pix / ((long long)HW * F));
  __nv_bfloat16* o = out + pix * Cpad;
  for (int c8 = 0; c8 < Cpad; c8 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c8 + i;
      float val = 0.f;
      if (c < C0) val = s0[(((size_t)(n % N0) * F + f) * C0 + c) * HW + p] * scale0;
      else if (c < C0 + C1) val = s1[(((size_t)(n % N1) * F + f) * C1 + (c - C0)) * HW + p];
      v[i] = val;
    }
    *reinterpret_cast<uint4*>(o + c8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

__global__ void unpack_output_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, long long NF,
                                     int C, int HW) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NF * C * HW) return;
  const int p = (int)(idx % HW);
  const int c = (int)((idx / HW) % C);
  const long long nf = idx / ((long long)HW * C);
  dst[idx] = src[(nf * HW + p) * ld + c];
}

// 32x32 smem-tiled transposes between [N, C, HW] fp32 and [N, HW, C] bf16
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * HW + p) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(src[((size_t)n * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

__device__ __forceinline__ uint4 load8_as_bf16(const void* base, long long vec_index, int f32) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(base) + 2 * vec_index;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
  return __ldg(reinterpret_cast<const uint4*>(base) + vec_index);
}

__global__ void upsample2x_kernel(const void* __restrict__ src, int src_f32, uint4* __restrict__ dst, int N, int H,
                                  int W, int vecs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over output vectors
  const long long total = (long long)N * 4 * H * W * vecs;
  if (idx >= total) return;
  const int v = (int)(idx % vecs);
  long long pix = idx / vecs;
  const int wo = (int)(pix % (2 * W));
  const int ho = (int)((pix / (2 * W)) % (2 * H));
  const int n = (int)(pix / ((long long)4 * H * W));
  dst[idx] = load8_as_bf16(src, (((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * vecs + v, src_f32);
}

__global__ void concat_kernel(const void* __restrict__ a, int va, const void* __restrict__ b, int vb, int src_f32,
                              uint4* __restrict__ dst, long long M) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vt = va + vb;
  if (idx >= M * vt) return;
  const long long m = idx / vt;
  const int v = (int)(idx % vt);
  dst[idx] = v < va ? load8_as_bf16(a, m * va + v, src_f32) : load8_as_bf16(b, m * vb + (v - va), src_f32);
}

__device__ __forceinline__ void load8_f(const void* base, long long vec_index, int f32, float (&f)[8]) {
  if (f32) {
    const float4* p = reinterpret_cast<const float4*>(base) + 2 * vec_index;
    const float4 a = p[0], b = p[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack_bf16x8(reinterpret_cast<const uint4*>(base)[vec_index], f);
  }
}

__global__ void axpby_kernel(const void* __restrict__ x, int x_f32, float alpha, void* __restrict__ y, int y_f32,
                             float beta, long long nvec) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nvec) return;
  float fx[8], fy[8];
  load8_f(x, idx, x_f32, fx);
  load8_f(y, idx, y_f32, fy);
#pragma unroll
  for (int i = 0; i < 8; ++i) fy[i] = alpha * fx[i] + beta * fy[i];
  if (y_f32) {
    float4* p = reinterpret_cast<float4*>(y) + 2 * idx;
    p[0] = make_float4(fy[0], fy[1], fy[2], fy[3]);
    p[1] = make_float4(fy[4], fy[5], fy[6], fy[7]);
  } else {
    reinterpret_cast<uint4*>(y)[idx] = make_uint4(pack_bf16x2(fy[0], fy[1]), pack_bf16x2(fy[2], fy[3]),
                                                  pack_bf16x2(fy[4], fy[5]), pack_bf16x2(fy[6], fy[7]));
  }
}

__global__ void cast_bf16_kernel(const void* __restrict__ src, uint4* __restrict__ dst, long long nvec) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < nvec) dst[idx] = load8_as_bf16(src, idx, 1);
}

// thread = one (s, f, c, p) element of the fp32 latent; pred is channels-last [2S*F*HW, ld]
__global__ void scale_f32_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y, long long n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) y[idx] = alpha * x[idx];
}

// x_next may alias x (in-place update of a CUDA graph's static latent buffer): every thread reads its x before it writes.
// pred_c != nullptr (ld > 0): the conditional half lives at its own base (row s, not S + s) - the CFG-pair split, where one
// of the two bases is the PARTNER GPU's buffer mapped into this address space (read over NVLink, no all-gather)
__global__ void cfg_euler_kernel(const float* __restrict__ pred, const float* __restrict__ pred_c, int ld, int cfg,
                                 const float* __restrict__ guidance,
                                 const float* x, float* x_next, float* __restrict__ v_out, float* __restrict__ x0_out,
                                 int S, int F, int C, int HW, float sigma, float sigma_next,
                                 const float* __restrict__ sigmas_dev) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * F * C * HW;
  if (idx >= total) return;
  if (sigmas_dev != nullptr) {                               // per-step scalars of a replayed CUDA graph
    sigma = __ldg(sigmas_dev);
    sigma_next = __ldg(sigmas_dev + 1);
  }
  const int p = (int)(idx % HW);
  const int c = (int)((idx / HW) % C);
  const int f = (int)((idx / ((long long)HW * C)) % F);
  const int s = (int)(idx / ((long long)HW * C * F));
  float v, cnd = 0.f;
  if (ld > 0) {   // channels-last prediction rows
    v = pred[(((size_t)s * F + f) * HW + p) * ld + c];
    if (cfg) cnd = pred_c != nullptr ? pred_c[(((size_t)s * F + f) * HW + p) * ld + c]
                                     : pred[(((size_t)(S + s) * F + f) * HW + p) * ld + c];
  } else {        // ld == 0: prediction in the latent's own [.,F,C,H,W] layout
    v = pred[idx];
    if (cfg) cnd = pred[total + idx];
  }
  if (cfg) v = v + guidance[f] * (cnd - v);
  if (v_out) v_out[idx] = v;
  const float xs = x[idx];
  const float s2 = sigma * sigma + 1.0f;
  const float x0 = v * (-sigma / sqrtf(s2)) + xs / s2;
  if (x0_out) x0_out[idx] = x0;
  const float deriv = (xs - x0) / sigma;
  x_next[idx] = xs + deriv * (sigma_next - sigma);
}

// Bidirectional ("direct fusion") Euler step of the reference's trans pipelines
// (pipeline/pipeline_stable_video_diffusion_trans_controlnet.py:639-667): the batch holds a forward half and a
// time-reversed half; their denoised predictions x0 are blended frame by frame, x0 = w[f] x0_fwd[f] + (1 - w[f])
// x0_bwd[F-1-f] with w = linspace(1, 0, F), the backward half takes the flipped blend, then both halves take the Euler
// step.  One thread per (sample, frame, channel, pixel) of the FORWARD half computes both outputs.
__global__ void fusion_euler_kernel(const float* __restrict__ v, const float* __restrict__ x,
                                    const float* __restrict__ w, float* __restrict__ x_next, int S, int F, int CHW,
                                    float sigma, float sigma_next) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long half = (long long)S * F * CHW;
  if (idx >= half) return;
  const int e = (int)(idx % CHW);
  const int f = (int)((idx / CHW) % F);
  const int s = (int)(idx / ((long long)CHW * F));
  const long long ib = half + ((long long)s * F + (F - 1 - f)) * CHW + e;   // backward half, mirrored frame
  const float s2 = sigma * sigma + 1.0f;
  const float k = -sigma / sqrtf(s2);
  const float xf = x[idx], xb = x[ib];
  const float x0f = v[idx] * k + xf / s2;
  const float x0b = v[ib] * k + xb / s2;
  const float wf = w[f];
  const float x0 = x0f * wf + x0b * (1.0f - wf);
  const float dt = sigma_next - sigma;
  x_next[idx] = xf + (xf - x0) / sigma * dt;
  x_next[ib] = xb + (xb - x0) / sigma * dt;
}

__global__ void axpy_f32_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y, long long n) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) y[idx] = fmaf(alpha, x[idx], y[idx]);
}

// mode 0: (re, im) -> (mag, pha) = (|z|, atan2(im, re));  mode 1: (mag, pha) -> (re, im)
__global__ void polar_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o0,
                             float* __restrict__ o1, int n, int mode) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (mode == 0) {
    o0[idx] = hypotf(a[idx], b[idx]);
    o1[idx] = atan2f(b[idx] + 0.0f, a[idx]);   // -0 + 0 = +0: a zero imaginary part gives angle 0 / +pi like torch.angle
  } else {
    o0[idx] = a[idx] * cosf(b[idx]);
    o1[idx] = a[idx] * sinf(b[idx]);
  }
}

}  // namespace lkgd

using namespace lkgd;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int lkgd_small_linear(const float* x, int32_t ldx, const float* W, const float* b, float* y, int32_t ldy,
                                 int32_t M, int32_t N, int32_t K, int32_t act_in, int32_t act_out, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || M > 4096) return LKGD_ESHAPE;
  const int vec = (K % 4 == 0) && (ldx % 4 == 0) && aligned16(x) && aligned16(W);
  small_linear_kernel<<<blocks_for(N, 8), 256, 0, ST(stream)>>>(x, ldx, W, b, y, ldy, M, N, K, act_in, act_out, vec);
  return launch_epilogue();
}

extern "C" int lkgd_timestep_embedding(const float* t, int32_t M, int32_t dim, float* out, void* stream) {
  if (M <= 0 || dim < 2) return LKGD_ESHAPE;
  timestep_embedding_kernel<<<blocks_for((long long)M * (dim / 2), 128), 128, 0, ST(stream)>>>(t, M, dim, out);
  return launch_epilogue();
}

extern "C" int lkgd_pack_input(const float* src0, int32_t N0, int32_t C0, float scale0, const float* scale0_dev,
                               const float* src1, int32_t N1, int32_t C1, void* out, int32_t N, int32_t F, int32_t H,
                               int32_t W, int32_t Cpad, void* stream) {
  if (src1 == nullptr) { C1 = 0; N1 = 1; }
  if (N <= 0 || F <= 0 || Cpad % 8 || C0 + C1 > Cpad || N0 <= 0 || N1 <= 0) return LKGD_ESHAPE;
  if (!aligned16(out)) return LKGD_EALIGN;
  const long long total = (long long)N * F * H * W;
  pack_input_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(src0, N0, C0, scale0, scale0_dev, src1, N1, C1,
                                                                   reinterpret_cast<__nv_bfloat16*>(out), N, F,
                                                                   H * W, Cpad);
  return launch_epilogue();
}

extern "C" int lkgd_unpack_output(const float* src, int32_t ld, float* dst, int32_t NF, int32_t C, int32_t H,
                                  int32_t W, void* stream) {
  if (NF <= 0 || C <= 0 || C > ld) return LKGD_ESHAPE;
  const long long total = (long long)NF * C * H * W;
  unpack_output_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(src, ld, dst, NF, C, H * W);
  return launch_epilogue();
}

extern "C" int lkgd_nchw_to_nhwc(const float* src, void* dst, int32_t N, int32_t C, int32_t H, int32_t W,
                                 void* stream) {
  if (N <= 0 || C <= 0 || N > 65535) return LKGD_ESHAPE;
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, ST(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), C, H * W);
  return launch_epilogue();
}

extern "C" int lkgd_nhwc_to_nchw(const void* src, float* dst, int32_t N, int32_t C, int32_t H, int32_t W,
                                 void* stream) {
  if (N <= 0 || C <= 0 || N > 65535) return LKGD_ESHAPE;
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, ST(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, C, H * W);
  return launch_epilogue();
}

extern "C" int lkgd_upsample2x(const void* src, int32_t src_f32, void* dst, int32_t N, int32_t H, int32_t W,
                               int32_t C, void* stream) {
  if (C % 8 || N <= 0) return LKGD_ESHAPE;
  if (!aligned16(src) || !aligned16(dst)) return LKGD_EALIGN;
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(src, src_f32, reinterpret_cast<uint4*>(dst), N,
                                                                   H, W, C / 8);
  return launch_epilogue();
}

extern "C" int lkgd_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n <= 0 || n % 8) return LKGD_ESHAPE;
  if (!aligned16(src) || !aligned16(dst)) return LKGD_EALIGN;
  cast_bf16_kernel<<<blocks_for(n / 8, 256), 256, 0, ST(stream)>>>(src, reinterpret_cast<uint4*>(dst), n / 8);
  return launch_epilogue();
}

extern "C" int lkgd_concat_channels(const void* a, int32_t Ca, const void* b, int32_t Cb, int32_t src_f32, void* dst,
                                    int64_t M, void* stream) {
  if (Ca % 8 || Cb % 8 || M <= 0) return LKGD_ESHAPE;
  if (!aligned16(a) || !aligned16(b) || !aligned16(dst)) return LKGD_EALIGN;
  const long long total = (long long)M * ((Ca + Cb) / 8);
  concat_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(a, Ca / 8, b, Cb / 8, src_f32,
                                                               reinterpret_cast<uint4*>(dst), M);
  return launch_epilogue();
}

extern "C" int lkgd_axpby(const void* x, int32_t x_f32, float alpha, void* y, int32_t y_f32, float beta, int64_t n,
                          void* stream) {
  if (n <= 0 || n % 8) return LKGD_ESHAPE;
  if (!aligned16(x) || !aligned16(y)) return LKGD_EALIGN;
  axpby_kernel<<<blocks_for(n / 8, 256), 256, 0, ST(stream)>>>(x, x_f32, alpha, y, y_f32, beta, n / 8);
  return launch_epilogue();
}

extern "C" int lkgd_cfg_euler_step(const float* pred, int32_t ld, int32_t cfg, const float* guidance, const float* x,
                                   float* x_next, float* v_out, float* x0_out, int32_t S, int32_t F, int32_t C,
                                   int32_t H, int32_t W, float sigma, float sigma_next, const float* sigmas_dev,
                                   void* stream) {
  if (S <= 0 || F <= 0 || C <= 0 || (ld != 0 && C > ld) || (sigmas_dev == nullptr && sigma <= 0.f)) return LKGD_ESHAPE;
  const long long total = (long long)S * F * C * H * W;
  cfg_euler_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(pred, nullptr, ld, cfg, guidance, x, x_next, v_out,
                                                                  x0_out, S, F, C, H * W, sigma, sigma_next, sigmas_dev);
  return launch_epilogue();
}

extern "C" int lkgd_cfg_euler_step_pair(const float* pred_uncond, const float* pred_cond, int32_t ld,
                                        const float* guidance, const float* x, float* x_next, float* v_out,
                                        float* x0_out, int32_t S, int32_t F, int32_t C, int32_t H, int32_t W, float sigma,
                                        float sigma_next, const float* sigmas_dev, void* stream) {
  if (pred_uncond == nullptr || pred_cond == nullptr || guidance == nullptr) return LKGD_ESHAPE;
  if (S <= 0 || F <= 0 || C <= 0 || ld <= 0 || C > ld || (sigmas_dev == nullptr && sigma <= 0.f)) return LKGD_ESHAPE;
  const long long total = (long long)S * F * C * H * W;
  cfg_euler_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(pred_uncond, pred_cond, ld, 1, guidance, x, x_next,
                                                                  v_out, x0_out, S, F, C, H * W, sigma, sigma_next,
                                                                  sigmas_dev);
  return launch_epilogue();
}

extern "C" int lkgd_fusion_euler_step(const float* v, const float* x, const float* weights, float* x_next, int32_t S,
                                      int32_t F, int32_t C, int32_t H, int32_t W, float sigma, float sigma_next,
                                      void* stream) {
  if (S <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0 || sigma <= 0.f) return LKGD_ESHAPE;
  const long long half = (long long)S * F * C * H * W;
  fusion_euler_kernel<<<blocks_for(half, 256), 256, 0, ST(stream)>>>(v, x, weights, x_next, S, F, C * H * W, sigma,
                                                                     sigma_next);
  return launch_epilogue();
}

// ViT patch unfold: one thread per (patch, 8 output columns)
__global__ void patchify_kernel(const float* __restrict__ x, int C, int H, int W, int P, __nv_bfloat16* __restrict__ out,
                                int Kpad, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k8 = Kpad / 8;
  const int kc = (int)(idx % k8) * 8;
  const long long patch = idx / k8;
  const int gw = W / P, gh = H / P;
  const int px0 = (int)(patch % gw), py0 = (int)((patch / gw) % gh);
  const long long n = patch / ((long long)gw * gh);
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kc + i;
    float val = 0.f;
    if (k < C * P * P) {
      const int c = k / (P * P), r = k % (P * P), py = r / P, px = r % P;
      val = __ldg(x + ((n * C + c) * H + (py0 * P + py)) * (long long)W + px0 * P + px);
    }
    v[i] = val;
  }
  *reinterpret_cast<uint4*>(out + patch * Kpad + kc) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

extern "C" int lkgd_patchify(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int32_t P, void* out, int32_t Kpad,
                             void* stream) {
  if (N <= 0 || C <= 0 || P <= 0 || H % P || W % P || Kpad % 8 || Kpad < C * P * P) return LKGD_ESHAPE;
  if (!aligned16(out)) return LKGD_EALIGN;
  const long long total = (long long)N * (H / P) * (W / P) * (Kpad / 8);
  patchify_kernel<<<blocks_for(total, 256), 256, 0, ST(stream)>>>(x, C, H, W, P, reinterpret_cast<__nv_bfloat16*>(out),
                                                                  Kpad, total);
  return launch_epilogue();
}

// out[m, :] = srcs[g(m)][m, :]  (bf16 rows, 16-byte vectors): the per-row choice between attention results computed against
// different contexts - temporal cross-attention with more than one key under the diffusers 0.27.2 context order, where
// row m of the temporal batch attends to context g(m) = ((m / (HW F)) HW + m % HW) % B.
struct SelectSrcs { const uint4* p[8]; };
__device__ __forceinline__ int sel_index(int mode, long long m, int HW, int F, int B) {
  switch (mode) {
    case LKGD_RV_FRAME: return (int)(m / HW);
    case LKGD_RV_FRAMEPOS: return (int)((m / HW) % F);
    case LKGD_RV_BATCH: return (int)(m / ((long long)HW * F));
    case LKGD_RV_TCTX_0272: return (int)(((m / ((long long)HW * F)) * HW + (m % HW)) % B);
    default: return 0;
  }
}
__global__ void select_rows_kernel(SelectSrcs srcs, int n_src, uint4* __restrict__ out, long long M, int C8, int mode,
                                   int HW, int F, int B) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * C8) return;
  const long long m = idx / C8;
  int g = sel_index(mode, m, HW, F, B);
  g = g < n_src ? g : n_src - 1;
  out[idx] = __ldg(srcs.p[g] + idx);
}

extern "C" int lkgd_select_rows(const void* const* srcs, int32_t n_src, void* out, int64_t M, int32_t C, int32_t rv_mode,
                                int32_t rv_HW, int32_t rv_F, int32_t rv_B, void* stream) {
  if (srcs == nullptr || n_src <= 0 || n_src > 8 || M <= 0 || C <= 0 || C % 8 || rv_HW <= 0 || rv_F <= 0 || rv_B <= 0)
    return LKGD_ESHAPE;
  SelectSrcs s;
  for (int i = 0; i < 8; ++i) {
    s.p[i] = reinterpret_cast<const uint4*>(srcs[i < n_src ? i : n_src - 1]);
    if (!aligned16(s.p[i])) return LKGD_EALIGN;
  }
  if (!aligned16(out)) return LKGD_EALIGN;
  select_rows_kernel<<<blocks_for(M * (C / 8), 256), 256, 0, ST(stream)>>>(s, n_src, reinterpret_cast<uint4*>(out), M,
                                                                            C / 8, rv_mode, rv_HW, rv_F, rv_B);
  return launch_epilogue();
}

extern "C" int lkgd_axpy_f32(const float* x, float alpha, float* y, int64_t n, void* stream) {
  if (n <= 0) return LKGD_ESHAPE;
  axpy_f32_kernel<<<blocks_for(n, 256), 256, 0, ST(stream)>>>(x, alpha, y, n);
  return launch_epilogue();
}

extern "C" int lkgd_polar(const float* a, const float* b, float* o0, float* o1, int32_t n, int32_t mode,
                          void* stream) {
  if (n <= 0 || (mode != 0 && mode != 1)) return LKGD_ESHAPE;
  polar_kernel<<<blocks_for(n, 128), 128, 0, ST(stream)>>>(a, b, o0, o1, n, mode);
  return launch_epilogue();
}

extern "C" int lkgd_scale_f32(const float* x, float alpha, float* y, int64_t n, void* stream) {
  if (n <= 0) return LKGD_ESHAPE;
  scale_f32_kernel<<<blocks_for(n, 256), 256, 0, ST(stream)>>>(x, alpha, y, n);
  return launch_epilogue();
}
