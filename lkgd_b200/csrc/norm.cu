// GroupNorm(+SiLU) and LayerNorm for channels-last bf16 activations: HBM-bound, 128-bit vectorised.
// GroupNorm is two kernels: (1) per-(sample, channel) sum / sum-of-squares -> fp64 atomics, (2) normalise.
// See include/lkgd_b200.h for the contract and the reference modules replaced.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

// thread layout shared by both GroupNorm kernels: blockDim.x = vecs * rows_par, thread -> (row lane, 8-ch vector)
struct GnGeom {
  int C1, C2, C, vecs, rows_par, R, rows_per_cta, x_f32;
};

// 8 consecutive channels of (row, vector v) from the first or the second (concatenated) source; bf16 or fp32 input
__device__ __forceinline__ void gn_load(const void* x1, const void* x2, const GnGeom& g, long long row, int v,
                                        float (&f)[8]) {
  const int c = v * 8;
  const void* base = c < g.C1 ? x1 : x2;
  const long long off = c < g.C1 ? row * g.C1 + c : row * g.C2 + (c - g.C1);
  if (g.x_f32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off)), f);
  }
}

__global__ void gn_stats_kernel(const void* __restrict__ x1, const void* __restrict__ x2, GnGeom g,
                                double* __restrict__ sums /* [NS][C][2] */) {
  extern __shared__ float sh[];  // [rows_par][vecs][16]
  const int v = threadIdx.x % g.vecs, rl = threadIdx.x / g.vecs;
  const int ns = blockIdx.y;
  const int r0 = blockIdx.x * g.rows_per_cta;
  const int r1 = min(r0 + g.rows_per_cta, g.R);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  // four rows' loads are issued before any is consumed (memory-level parallelism: one row per thread in flight left
  // the kernel latency-bound at ~60 % of HBM bandwidth)
  for (int r = r0 + rl; r < r1; r += 4 * g.rows_par) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * g.rows_par;
      if (rr < r1) gn_load(x1, x2, g, (long long)ns * g.R + rr, v, f[u]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[u][i]; q[i] = fmaf(f[u][i], f[u][i], q[i]); }
  }
  float* my = sh + (size_t)threadIdx.x * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { my[i] = s[i]; my[8 + i] = q[i]; }
  __syncthreads();
  // thread t < vecs*16 reduces one (vector, component) over the row lanes
  for (int t = threadIdx.x; t < g.vecs * 16; t += blockDim.x) {
    const int vv = t / 16, comp = t % 16;
    float acc = 0.f;
    for (int rr = 0; rr < g.rows_par; ++rr) acc += sh[((size_t)rr * g.vecs + vv) * 16 + comp];
    const int c = vv * 8 + (comp & 7);
    atomicAdd(&sums[((size_t)ns * g.C + c) * 2 + (comp >> 3)], (double)acc);
  }
}

// per (sample, channel) scale = rstd * gamma, shift = beta - mean * rstd * gamma, once per call (grid = NS): the apply
// CTAs then start streaming immediately instead of each re-deriving the statistics of all groups in a prologue
__global__ void gn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, int groups, int C, int R,
                                   float2* __restrict__ ss /* [NS][C] (scale, shift) */) {
  extern __shared__ float sh[];  // mean[groups], rstd[groups]
  float* gmean = sh;
  float* grstd = sh + groups;
  const int ns = blockIdx.x;
  const int cpg = C / groups;
  for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
      s += sums[((size_t)ns * C + c) * 2];
      q += sums[((size_t)ns * C + c) * 2 + 1];
    }
    const double n = (double)cpg * R;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    gmean[gi] = (float)mean;
    grstd[gi] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gi = c / cpg;
    const float sc = grstd[gi] * gamma[c];
    ss[(size_t)ns * C + c] = make_float2(sc, beta[c] - gmean[gi] * sc);
  }
}

// Same result as gn_finalize_kernel, but the sums are gathered from the per-(frame image, channel) sums that the
// producing GEMM launches accumulated (two sources = the channel concatenation of the up blocks).  One CTA of 128
// threads per (sample, group): cpg channels x fps frames x 2 doubles are reduced through shared memory, so the
// temporal GroupNorms (2 samples x 25 frames) are not two serial CTAs.
__global__ void __launch_bounds__(128) gn_finalize_frames_kernel(const double* __restrict__ st1, int C1,
                                                                 const double* __restrict__ st2, int C2, int fps,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, int groups,
                                                                 int R, float2* __restrict__ ss /* [NS][C] */,
                                                                 double* __restrict__ sums /* [NS][C][2] */) {
  __shared__ double red[2][128];
  const int C = C1 + C2;
  const int cpg = C / groups;
  const int ns = blockIdx.x / groups, gi = blockIdx.x % groups;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < cpg * fps; i += blockDim.x) {
    const int c = gi * cpg + i % cpg, f = i / cpg;
    const double* st = c < C1 ? st1 + (((size_t)ns * fps + f) * C1 + c) * 2
                              : st2 + (((size_t)ns * fps + f) * C2 + (c - C1)) * 2;
    s += st[0]; q += st[1];
  }
  red[0][threadIdx.x] = s; red[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const double n = (double)cpg * R;
  const double mean = red[0][0] / n;
  double var = red[1][0] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float fmean = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int c = gi * cpg + threadIdx.x; c < (gi + 1) * cpg; c += blockDim.x) {
    const float sc = rstd * gamma[c];
    ss[(size_t)ns * C + c] = make_float2(sc, beta[c] - fmean * sc);
    // per-(sample, channel) sums in the layout lkgd_groupnorm leaves behind: what the GroupNorm backward re-reads
    double cs = 0.0, cq = 0.0;
    for (int f = 0; f < fps; ++f) {
      const double* st = c < C1 ? st1 + (((size_t)ns * fps + f) * C1 + c) * 2
                                : st2 + (((size_t)ns * fps + f) * C2 + (c - C1)) * 2;
      cs += st[0]; cq += st[1];
    }
    sums[((size_t)ns * C + c) * 2] = cs;
    sums[((size_t)ns * C + c) * 2 + 1] = cq;
  }
}

__global__ void __launch_bounds__(512, 2) gn_apply_kernel(const void* __restrict__ x1, const void* __restrict__ x2, GnGeom g,
                                const float2* __restrict__ ss, int silu, __nv_bfloat16* __restrict__ out,
                                __nv_bfloat16* __restrict__ raw /* optional: the un-normalised input, narrowed */) {
  const int ns = blockIdx.y;
  const int v = threadIdx.x % g.vecs, rl = threadIdx.x / g.vecs;
  const int r0 = blockIdx.x * g.rows_per_cta;
  const int r1 = min(r0 + g.rows_per_cta, g.R);
  float sc[8], sf[8];
  {
    const float4* p4 = reinterpret_cast<const float4*>(ss + (size_t)ns * g.C + v * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(p4 + i);
      sc[2 * i] = t.x; sf[2 * i] = t.y; sc[2 * i + 1] = t.z; sf[2 * i + 1] = t.w;
    }
  }
  for (int r = r0 + rl; r < r1; r += 4 * g.rows_par) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * g.rows_par;
      if (rr < r1) gn_load(x1, x2, g, (long long)ns * g.R + rr, v, f[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * g.rows_par;
      if (rr < r1) {
        if (raw != nullptr)      // the resblock's 1x1 shortcut reads the raw (concatenated) input as a bf16 GEMM operand
          *reinterpret_cast<uint4*>(raw + ((long long)ns * g.R + rr) * g.C + v * 8) =
              make_uint4(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]), pack_bf16x2(f[u][4], f[u][5]),
                         pack_bf16x2(f[u][6], f[u][7]));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float y = fmaf(f[u][i], sc[i], sf[i]);
          f[u][i] = silu ? silu_fast(y) : y;
        }
        *reinterpret_cast<uint4*>(out + ((long long)ns * g.R + rr) * g.C + v * 8) =
            make_uint4(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]), pack_bf16x2(f[u][4], f[u][5]),
                       pack_bf16x2(f[u][6], f[u][7]));
      }
    }
  }
}

// The same pass with FOUR channels per thread (g.vecs = C / 4): 16 + 8 live values instead of 32 + 16, no spills under
// the 64-register bound of two 512-thread CTAs, 16-byte loads of fp32 rows / 8-byte stores.
__global__ void __launch_bounds__(512, 2) gn_apply4_kernel(const void* __restrict__ x1, const void* __restrict__ x2,
                                                           GnGeom g, const float2* __restrict__ ss, int silu,
                                                           __nv_bfloat16* __restrict__ out,
                                                           __nv_bfloat16* __restrict__ raw) {
  const int ns = blockIdx.y;
  const int v = threadIdx.x % g.vecs, rl = threadIdx.x / g.vecs;
  const int c0 = v * 4;
  const int r0 = blockIdx.x * g.rows_per_cta;
  const int r1 = min(r0 + g.rows_per_cta, g.R);
  float sc[4], sf[4];
  {
    const float4* p4 = reinterpret_cast<const float4*>(ss + (size_t)ns * g.C + c0);
    const float4 t0 = __ldg(p4), t1 = __ldg(p4 + 1);
    sc[0] = t0.x; sf[0] = t0.y; sc[1] = t0.z; sf[1] = t0.w; sc[2] = t1.x; sf[2] = t1.y; sc[3] = t1.z; sf[3] = t1.w;
  }
  const bool first = c0 < g.C1;
  const char* base = reinterpret_cast<const char*>(first ? x1 : x2);
  const long long ld = first ? g.C1 : g.C2;
  const int cc = first ? c0 : c0 - g.C1;
  const int es = g.x_f32 ? 4 : 2;
  for (int r = r0 + rl; r < r1; r += 4 * g.rows_par) {
    float f[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * g.rows_par;
      if (rr < r1) {
        const char* src = base + (((long long)ns * g.R + rr) * ld + cc) * es;
        if (g.x_f32) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          f[u][0] = t.x; f[u][1] = t.y; f[u][2] = t.z; f[u][3] = t.w;
        } else {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
          f[u][0] = __uint_as_float(t.x << 16); f[u][1] = __uint_as_float(t.x & 0xffff0000u);
          f[u][2] = __uint_as_float(t.y << 16); f[u][3] = __uint_as_float(t.y & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + u * g.rows_par;
      if (rr < r1) {
        const long long o = ((long long)ns * g.R + rr) * g.C + c0;
        if (raw != nullptr)
          *reinterpret_cast<uint2*>(raw + o) = make_uint2(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float y = fmaf(f[u][i], sc[i], sf[i]);
          f[u][i] = silu ? silu_fast(y) : y;
        }
        *reinterpret_cast<uint2*>(out + o) = make_uint2(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm
constexpr int LN_MAXV = 8;  // up to 8 x (32 lanes x 8 channels) = 2048 channels

__device__ __forceinline__ int ln_rowvec_index(int mode, long long m, int HW, int F, int B) {
  switch (mode) {
    case LKGD_RV_FRAME: return (int)(m / HW);
    case LKGD_RV_FRAMEPOS: return (int)((m / HW) % F);
    case LKGD_RV_BATCH: return (int)(m / ((long long)HW * F));
    case LKGD_RV_TCTX_0272: return (int)(((m / ((long long)HW * F)) * HW + (m % HW)) % B);
    case LKGD_RV_BATCH_TCTX: {
      const long long b = m / ((long long)HW * F);
      return (int)(b * B + (b * HW + (m % HW)) % B);
    }
    default: return 0;
  }
}

// One warp per row, ROWS rows per warp in flight (all loads of an iteration are issued before the first reduction:
// with a single 1.3 KB row per warp the kernel was latency-bound at ~50 % of HBM bandwidth), grid-stride over rows.
template <int NV, bool XF32, int ROWS>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ xv, int M, int C,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        const float* __restrict__ addvec, int addvec_ld, int rv_mode,
                                                        int rv_HW, int rv_F, int rv_B, void* sum_out_v,
                                                        __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row0 = warp_id * ROWS; row0 < M; row0 += warps * ROWS) {
    float f[ROWS][NV][8];
    // ---- loads (+ fused add) of all rows of this iteration
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const long long row = row0 + u;
      if (row < M) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int v = lane + 32 * j;
          if (v < nvec) {
            if (XF32) {
              const float* xr32 = reinterpret_cast<const float*>(xv) + row * C;
              const float4 x0 = *reinterpret_cast<const float4*>(xr32 + v * 8);
              const float4 x1 = *(reinterpret_cast<const float4*>(xr32 + v * 8) + 1);
              f[u][j][0] = x0.x; f[u][j][1] = x0.y; f[u][j][2] = x0.z; f[u][j][3] = x0.w;
              f[u][j][4] = x1.x; f[u][j][5] = x1.y; f[u][j][6] = x1.z; f[u][j][7] = x1.w;
            } else {
              const __nv_bfloat16* xr = reinterpret_cast<const __nv_bfloat16*>(xv) + row * C;
              unpack_bf16x8(*reinterpret_cast<const uint4*>(xr + v * 8), f[u][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const long long row = row0 + u;
      if (row >= M) break;
      const float* av = addvec ? addvec + (size_t)ln_rowvec_index(rv_mode, row, rv_HW, rv_F, rv_B) * addvec_ld : nullptr;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (v < nvec) {
          if (av) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + v * 8));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + v * 8) + 1);
            f[u][j][0] += a0.x; f[u][j][1] += a0.y; f[u][j][2] += a0.z; f[u][j][3] += a0.w;
            f[u][j][4] += a1.x; f[u][j][5] += a1.y; f[u][j][6] += a1.z; f[u][j][7] += a1.w;
            if (XF32) {
              if (sum_out_v) {
                float* so = reinterpret_cast<float*>(sum_out_v) + row * C + v * 8;
                *reinterpret_cast<float4*>(so) = make_float4(f[u][j][0], f[u][j][1], f[u][j][2], f[u][j][3]);
                *(reinterpret_cast<float4*>(so) + 1) = make_float4(f[u][j][4], f[u][j][5], f[u][j][6], f[u][j][7]);
              }
            } else {
              // bf16 residual stream: normalise what is actually stored
              uint4 o = make_uint4(pack_bf16x2(f[u][j][0], f[u][j][1]), pack_bf16x2(f[u][j][2], f[u][j][3]),
                                   pack_bf16x2(f[u][j][4], f[u][j][5]), pack_bf16x2(f[u][j][6], f[u][j][7]));
              if (sum_out_v)
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(sum_out_v) + row * C + v * 8) = o;
              unpack_bf16x8(o, f[u][j]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) s += f[u][j][i];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / C;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (lane + 32 * j < nvec) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float d = f[u][j][i] - mean; q = fmaf(d, d, q); }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / C + eps);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (v < nvec) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = (f[u][j][i] - mean) * rstd * gg[i] + bb[i];
          *reinterpret_cast<uint4*>(out + row * C + v * 8) =
              make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                         pack_bf16x2(y[6], y[7]));
        }
      }
    }
  }
}

// fp32 rows whose vector count splits evenly over LPR lanes (C = 320 / 640 / 1280: 5 vectors of 8 channels per lane on
// 8 / 16 / 32 lanes): a warp normalises 32 / LPR rows per pass, so the two reductions are log2(LPR) shuffle steps shared
// by all its rows (the row-per-warp kernel above spends 10 steps per row and leaves 3/8 of the lanes idle at C = 320),
// gamma / beta live in registers for the whole kernel, and two passes are in flight per warp.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, int M, int C,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             const float* __restrict__ addvec, int addvec_ld,
                                                             int rv_mode, int rv_HW, int rv_F, int rv_B,
                                                             float* sum_out, __nv_bfloat16* __restrict__ out) {
  constexpr int RW = 32 / LPR;     // rows per warp pass
  constexpr int U = 2;             // passes in flight
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.0f / (float)C;
  float gg[VPL][8], bb[VPL][8];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = l + LPR * k;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
    gg[k][0] = g0.x; gg[k][1] = g0.y; gg[k][2] = g0.z; gg[k][3] = g0.w; gg[k][4] = g1.x; gg[k][5] = g1.y; gg[k][6] = g1.z; gg[k][7] = g1.w;
    bb[k][0] = b0.x; bb[k][1] = b0.y; bb[k][2] = b0.z; bb[k][3] = b0.w; bb[k][4] = b1.x; bb[k][5] = b1.y; bb[k][6] = b1.z; bb[k][7] = b1.w;
  }
  for (long long row0 = warp_id * (RW * U); row0 < M; row0 += warps * (RW * U)) {
    float f[U][VPL][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * RW + sub;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (row < M) {
          const float4* p4 = reinterpret_cast<const float4*>(x + row * C + (l + LPR * k) * 8);
          const float4 a = p4[0], b = p4[1];
          f[u][k][0] = a.x; f[u][k][1] = a.y; f[u][k][2] = a.z; f[u][k][3] = a.w;
          f[u][k][4] = b.x; f[u][k][5] = b.y; f[u][k][6] = b.z; f[u][k][7] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[u][k][i] = 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * RW + sub;
      const bool live = row < M;
      if (addvec != nullptr && live) {
        const float* av = addvec + (size_t)ln_rowvec_index(rv_mode, row, rv_HW, rv_F, rv_B) * addvec_ld;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + (l + LPR * k) * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(av + (l + LPR * k) * 8) + 1);
          f[u][k][0] += a0.x; f[u][k][1] += a0.y; f[u][k][2] += a0.z; f[u][k][3] += a0.w;
          f[u][k][4] += a1.x; f[u][k][5] += a1.y; f[u][k][6] += a1.z; f[u][k][7] += a1.w;
        }
      }
      if (sum_out != nullptr && live) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          float4* o4 = reinterpret_cast<float4*>(sum_out + row * C + (l + LPR * k) * 8);
          o4[0] = make_float4(f[u][k][0], f[u][k][1], f[u][k][2], f[u][k][3]);
          o4[1] = make_float4(f[u][k][4], f[u][k][5], f[u][k][6], f[u][k][7]);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += f[u][k][i];
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = f[u][k][i] - mean; q = fmaf(d, d, q); }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + eps);
      if (live) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = (f[u][k][i] - mean) * rstd * gg[k][i] + bb[k][i];
          *reinterpret_cast<uint4*>(out + row * C + (l + LPR * k) * 8) =
              make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                         pack_bf16x2(y[6], y[7]));
        }
      }
    }
  }
}

static GnGeom gn_geom(int C1, int C2, int R, int x_f32) {
  GnGeom g;
  g.C1 = C1; g.C2 = C2; g.C = C1 + C2; g.vecs = g.C / 8; g.R = R; g.x_f32 = x_f32;
  g.rows_par = 512 / g.vecs;
  if (g.rows_par < 1) g.rows_par = 1;
  if (g.rows_par > R) g.rows_par = R;
  // 16 rows per thread per CTA (four batches of four loads in flight): >= 2 waves of CTAs at SVD sizes, and half the
  // fp64 atomics / CTA launches of the 8-row version
  static int rows_mult = 0;
  if (rows_mult == 0) { const char* e = getenv("LKGD_GN_ROWS"); rows_mult = e ? atoi(e) : 16; if (rows_mult < 4) rows_mult = 4; }
  g.rows_per_cta = g.rows_par * rows_mult;
  return g;
}

// normalise pass: four channels per thread where the row fits a CTA (C <= 2048), LKGD_GN_APPLY8 = the 8-channel kernel
static void launch_gn_apply(const void* x1, const void* x2, int C1, int C2, int NS, int R, int x_f32, const float2* ss,
                            int silu, void* out, void* raw, cudaStream_t st) {
  if ((C1 + C2) / 4 <= 512 && !getenv("LKGD_GN_APPLY8")) {
    GnGeom g = gn_geom(C1, C2, R, x_f32);
    g.vecs = g.C / 4;
    g.rows_par = 512 / g.vecs;
    if (g.rows_par < 1) g.rows_par = 1;
    if (g.rows_par > R) g.rows_par = R;
    g.rows_per_cta = g.rows_par * 16;
    dim3 grid((R + g.rows_per_cta - 1) / g.rows_per_cta, NS);
    gn_apply4_kernel<<<grid, g.vecs * g.rows_par, 0, st>>>(x1, x2, g, ss, silu, reinterpret_cast<__nv_bfloat16*>(out),
                                                          reinterpret_cast<__nv_bfloat16*>(raw));
    return;
  }
  GnGeom g = gn_geom(C1, C2, R, x_f32);
  dim3 grid((R + g.rows_per_cta - 1) / g.rows_per_cta, NS);
  gn_apply_kernel<<<grid, g.vecs * g.rows_par, 0, st>>>(x1, x2, g, ss, silu, reinterpret_cast<__nv_bfloat16*>(out),
                                                       reinterpret_cast<__nv_bfloat16*>(raw));
}

}  // namespace lkgd

using namespace lkgd;

// [NS][C][2] doubles (sum, sum of squares; what the backward re-reads) followed by [NS][C] float2 (scale, shift)
extern "C" size_t lkgd_groupnorm_workspace(int32_t NS, int32_t C) {
  return (size_t)NS * C * (2 * sizeof(double) + sizeof(float2));
}

extern "C" int lkgd_groupnorm(const void* x1, int32_t C1, const void* x2, int32_t C2, int32_t NS, int32_t R,
                              int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                              int32_t x_f32, void* out, void* workspace, size_t ws_bytes, void* stream) {
  if (x2 == nullptr) C2 = 0;
  const int C = C1 + C2;
  if (NS <= 0 || R <= 0 || C <= 0 || groups <= 0 || C % groups || C1 % 8 || C2 % 8 || C / 8 > 1024) return LKGD_ESHAPE;
  if (!aligned16(x1) || !aligned16(out) || (x2 && !aligned16(x2))) return LKGD_EALIGN;
  if (ws_bytes < lkgd_groupnorm_workspace(NS, C) || workspace == nullptr) return LKGD_EWS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnGeom g = gn_geom(C1, C2, R, x_f32);
  const size_t sums_bytes = (size_t)NS * C * 2 * sizeof(double);
  cudaError_t e = cudaMemsetAsync(workspace, 0, sums_bytes, st);
  if (e != cudaSuccess) return set_cuda_error(e);
  dim3 grid((R + g.rows_per_cta - 1) / g.rows_per_cta, NS);
  const int threads = g.vecs * g.rows_par;
  const size_t sh1 = (size_t)threads * 16 * sizeof(float);
  static DeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  }
  gn_stats_kernel<<<grid, threads, sh1, st>>>(x1, x2, g, reinterpret_cast<double*>(workspace));
  int rc = launch_epilogue();
  if (rc) return rc;
  float2* ss = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + sums_bytes);
  gn_finalize_kernel<<<NS, 256, 2 * groups * sizeof(float), st>>>(reinterpret_cast<const double*>(workspace), gamma, beta,
                                                                 eps, groups, C, R, ss);
  rc = launch_epilogue();
  if (rc) return rc;
  launch_gn_apply(x1, x2, C1, C2, NS, R, x_f32, ss, silu, out, nullptr, st);
  return launch_epilogue();
}

extern "C" int lkgd_groupnorm_from_stats(const void* x1, int32_t C1, const double* stats1, const void* x2, int32_t C2,
                                         const double* stats2, int32_t NS, int32_t R, int32_t frames_per_sample,
                                         int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                                         int32_t x_f32, void* out, void* raw_out, void* workspace, size_t ws_bytes,
                                         void* stream) {
  if (x2 == nullptr) C2 = 0;
  const int C = C1 + C2;
  if (NS <= 0 || R <= 0 || C <= 0 || groups <= 0 || C % groups || C1 % 8 || C2 % 8 || C / 8 > 1024) return LKGD_ESHAPE;
  if (frames_per_sample <= 0 || R % frames_per_sample || stats1 == nullptr || (x2 != nullptr && stats2 == nullptr))
    return LKGD_ESHAPE;
  if (!aligned16(x1) || !aligned16(out) || (x2 && !aligned16(x2)) || (raw_out && !aligned16(raw_out))) return LKGD_EALIGN;
  if (ws_bytes < lkgd_groupnorm_workspace(NS, C) || workspace == nullptr) return LKGD_EWS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnGeom g = gn_geom(C1, C2, R, x_f32);
  const size_t sums_bytes = (size_t)NS * C * 2 * sizeof(double);
  float2* ss = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + sums_bytes);
  gn_finalize_frames_kernel<<<NS * groups, 128, 0, st>>>(stats1, C1, stats2, C2, frames_per_sample, gamma, beta, eps, groups,
                                                          R, ss, reinterpret_cast<double*>(workspace));
  int rc = launch_epilogue();
  if (rc) return rc;
  launch_gn_apply(x1, x2, C1, C2, NS, R, x_f32, ss, silu, out, raw_out, st);
  return launch_epilogue();
}

extern "C" int lkgd_layernorm(const void* x, int32_t M, int32_t C, const float* gamma, const float* beta, float eps,
                              const float* addvec, int32_t addvec_ld, int32_t rv_mode, int32_t rv_HW, int32_t rv_F,
                              int32_t rv_B, int32_t x_f32, void* sum_out, void* out, void* stream) {
  if (M <= 0 || C <= 0 || C % 8 || C > LN_MAXV * 256) return LKGD_ESHAPE;
  if (!aligned16(x) || !aligned16(out) || (sum_out && !aligned16(sum_out)) || (addvec && !aligned16(addvec)) ||
      !aligned16(gamma) || !aligned16(beta))
    return LKGD_EALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nv = (C / 8 + 31) / 32;
  if (addvec_ld <= 0) addvec_ld = C;
  if (addvec_ld % 4) return LKGD_EALIGN;
  if (rv_HW <= 0) rv_HW = 1;
  if (rv_F <= 0) rv_F = 1;
  if (rv_B <= 0) rv_B = 1;
#define LN_LAUNCH(NV, ROWS)                                                                                        \
  do {                                                                                                             \
    int g_ = (M + 8 * (ROWS) - 1) / (8 * (ROWS));                                                                  \
    if (g_ > 8 * sm_count()) g_ = 8 * sm_count();                                                                  \
    if (x_f32)                                                                                                     \
      layernorm_kernel<NV, true, ROWS><<<g_, 256, 0, st>>>(x, M, C, gamma, beta, eps, addvec, addvec_ld, rv_mode,  \
                                                           rv_HW, rv_F, rv_B, sum_out,                             \
                                                           reinterpret_cast<__nv_bfloat16*>(out));                 \
    else                                                                                                           \
      layernorm_kernel<NV, false, ROWS><<<g_, 256, 0, st>>>(x, M, C, gamma, beta, eps, addvec, addvec_ld, rv_mode, \
                                                            rv_HW, rv_F, rv_B, sum_out,                            \
                                                            reinterpret_cast<__nv_bfloat16*>(out));                \
  } while (0)
  // fp32 rows of 5 vectors per lane on 8 / 16 / 32 lanes (C = 320 / 640 / 1280): the rows-per-warp kernel
  if (x_f32 && C % 40 == 0 && (C / 40 == 8 || C / 40 == 16 || C / 40 == 32) && !getenv("LKGD_LN_ROWWARP")) {
    const int lpr = C / 40, rows_per_warp_iter = (32 / lpr) * 2;
    long long g_ = ((long long)M + 8 * rows_per_warp_iter - 1) / (8 * rows_per_warp_iter);
    if (g_ > sm_count()) g_ = sm_count();        // 234 registers: one persistent CTA per SM, grid-stride over rows
    const float* xf = reinterpret_cast<const float*>(x);
    float* so = reinterpret_cast<float*>(sum_out);
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out);
    if (lpr == 8)
      layernorm_rows_kernel<8, 5><<<(unsigned)g_, 256, 0, st>>>(xf, M, C, gamma, beta, eps, addvec, addvec_ld, rv_mode, rv_HW, rv_F, rv_B, so, ob);
    else if (lpr == 16)
      layernorm_rows_kernel<16, 5><<<(unsigned)g_, 256, 0, st>>>(xf, M, C, gamma, beta, eps, addvec, addvec_ld, rv_mode, rv_HW, rv_F, rv_B, so, ob);
    else
      layernorm_rows_kernel<32, 5><<<(unsigned)g_, 256, 0, st>>>(xf, M, C, gamma, beta, eps, addvec, addvec_ld, rv_mode, rv_HW, rv_F, rv_B, so, ob);
    return launch_epilogue();
  }
  switch (nv) {
    case 1: LN_LAUNCH(1, 4); break;
    case 2: LN_LAUNCH(2, 4); break;
    case 3: LN_LAUNCH(3, 2); break;
    case 4: LN_LAUNCH(4, 2); break;
    case 5: LN_LAUNCH(5, 2); break;
    default: LN_LAUNCH(8, 1); break;
  }
#undef LN_LAUNCH
  return launch_epilogue();
}
