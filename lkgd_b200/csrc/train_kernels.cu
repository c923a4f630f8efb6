// Backward / training kernels of the LoRA fine-tuning step (reference train_models/train_svd_lora.py:1445-1689):
// GroupNorm(+SiLU) and LayerNorm backward, GEGLU forward / backward on the tile-interleaved projection, grouped column
// sums (gradient of the KV-length-1 cross-attention vectors), 2x2 sum (nearest-upsample backward), zero stuffing
// (stride-2 conv data gradient), EDM preconditioning + weighted-MSE loss forward/backward, fused AdamW, sum of squares.
// All HBM-bound, 128-bit vectorised.  Contracts: include/lkgd_b200.h.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

__device__ __forceinline__ void ld8_f32(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8_f32(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  unpack_bf16x8(*reinterpret_cast<const uint4*>(p), f);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void ld8_any(const void* base, long long off, int is_f32, float (&f)[8]) {
  if (is_f32) ld8_f32(reinterpret_cast<const float*>(base) + off, f);
  else ld8_bf16(reinterpret_cast<const __nv_bfloat16*>(base) + off, f);
}

// ------------------------------------------------------------------------------------------- GroupNorm backward
// blockDim.x = vecs * rows_par, thread -> (row lane, 4 channels).  Four channels per thread (not the forward's eight)
// keep both passes at <= 64 registers: four 256-thread CTAs per SM and two rows of loads in flight per thread - the
// 8-channel version sat at 96 registers = 2 CTAs / SM and reached a third of the HBM bandwidth on the 46 MB tensors of
// a training clip (ncu: 44 + 57 us per GroupNorm at level 0).
struct GnbGeom {
  int C1, C2, C, vecs, rows_par, R, rows_per_cta, x_f32, groups, silu;
  float eps;
};

__device__ __forceinline__ void ld4_any(const void* base, long long off, int is_f32, float (&f)[4]) {
  if (is_f32) {
    const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  }
}
__device__ __forceinline__ void gnb_load_x(const void* x1, const void* x2, const GnbGeom& g, long long row, int c,
                                           float (&f)[4]) {
  if (c < g.C1) ld4_any(x1, row * g.C1 + c, g.x_f32, f);
  else ld4_any(x2, row * g.C2 + (c - g.C1), g.x_f32, f);
}

// smem prologue shared by both passes: the groups' mean / rstd from the forward's per-channel (sum, sum of squares);
// with ``bsums`` (pass 2) also the group means of g and g * xhat from pass 1.  Eight lanes per group, four groups per
// warp and round, every load of a round in flight together (a thread-per-group loop is cpg dependent L2 round trips).
__device__ __forceinline__ void gnb_prologue(const GnbGeom& g, int ns, const double* fsums, const double* bsums,
                                             float* gmean, float* grstd, float* m1, float* m2) {
  const int cpg = g.C / g.groups;
  const double n = (double)cpg * g.R;
  const int nw = blockDim.x >> 5;
  if (nw > 0) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, sl = lane & 7;
    if (w < nw) {
      for (int g0 = w * 4; g0 < g.groups; g0 += nw * 4) {
        const int gi = g0 + sub;
        double s = 0.0, q = 0.0, bs = 0.0, bq = 0.0;
        if (gi < g.groups) {
          for (int c = gi * cpg + sl; c < (gi + 1) * cpg; c += 8) {
            const size_t at = ((size_t)ns * g.C + c) * 2;
            s += fsums[at]; q += fsums[at + 1];
            if (bsums != nullptr) { bs += bsums[at]; bq += bsums[at + 1]; }
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o);
          bs += __shfl_xor_sync(0xffffffffu, bs, o); bq += __shfl_xor_sync(0xffffffffu, bq, o);
        }
        if (gi < g.groups && sl == 0) {
          const double mean = s / n;
          double var = q / n - mean * mean;
          if (var < 0.0) var = 0.0;
          gmean[gi] = (float)mean;
          grstd[gi] = (float)(1.0 / sqrt(var + (double)g.eps));
          if (bsums != nullptr) { m1[gi] = (float)(bs / n); m2[gi] = (float)(bq / n); }
        }
      }
    }
  } else {                     // fewer than 32 threads (a handful of rows): thread per group
    for (int gi = threadIdx.x; gi < g.groups; gi += blockDim.x) {
      double s = 0.0, q = 0.0, bs = 0.0, bq = 0.0;
      for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
        const size_t at = ((size_t)ns * g.C + c) * 2;
        s += fsums[at]; q += fsums[at + 1];
        if (bsums != nullptr) { bs += bsums[at]; bq += bsums[at + 1]; }
      }
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      gmean[gi] = (float)mean;
      grstd[gi] = (float)(1.0 / sqrt(var + (double)g.eps));
      if (bsums != nullptr) { m1[gi] = (float)(bs / n); m2[gi] = (float)(bq / n); }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float silu_grad(float z) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-z));     // MUFU.RCP: the IEEE division was 12 of ~30 instructions per element
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// g = dy * act'(z) * gamma and xhat for one 4-channel piece of a row (z = xhat * gamma + beta)
__device__ __forceinline__ void gnb_piece(const float (&f)[4], const float (&d)[4], const float (&gm)[4],
                                          const float (&bt)[4], const float (&xr)[4], const float (&xmr)[4], int silu,
                                          float (&gg)[4], float (&xh)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xh[i] = fmaf(f[i], xr[i], xmr[i]);
    const float z = fmaf(xh[i], gm[i], bt[i]);
    gg[i] = d[i] * (silu ? silu_grad(z) : 1.0f) * gm[i];
  }
}

// pass 1: per (sample, channel)  sum g  and  sum g * xhat
__global__ void __launch_bounds__(1024) gn_bwd_stats_kernel(const void* __restrict__ x1, const void* __restrict__ x2,
                                    const __nv_bfloat16* __restrict__ dy, GnbGeom g,
                                    const double* __restrict__ fsums, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, double* __restrict__ bsums) {
  extern __shared__ float sh[];
  float* gmean = sh;
  float* grstd = gmean + g.groups;
  float* red = grstd + g.groups;   // [8][threads]
  const int ns = blockIdx.y;
  gnb_prologue(g, ns, fsums, nullptr, gmean, grstd, nullptr, nullptr);
  const int v = threadIdx.x % g.vecs, rl = threadIdx.x / g.vecs;
  const int cpg = g.C / g.groups;
  const int c0 = v * 4;
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.R);
  float gm[4], bt[4], xr[4], xmr[4], s1[4], s2[4];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0));
    gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w;
    bt[0] = b.x; bt[1] = b.y; bt[2] = b.z; bt[3] = b.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = (c0 + i) / cpg;
    xr[i] = grstd[gi]; xmr[i] = -gmean[gi] * grstd[gi];
    s1[i] = 0.f; s2[i] = 0.f;
  }
  for (int r = r0 + rl; r < r1; r += 2 * g.rows_par) {
    const long long row = (long long)ns * g.R + r;
    const bool two = r + g.rows_par < r1;
    const long long rowb = two ? row + g.rows_par : row;
    float fa[4], da[4], fb[4], db[4], gg[4], xh[4];
    gnb_load_x(x1, x2, g, row, c0, fa);
    ld4_any(dy, row * g.C + c0, 0, da);
    gnb_load_x(x1, x2, g, rowb, c0, fb);
    ld4_any(dy, rowb * g.C + c0, 0, db);
    gnb_piece(fa, da, gm, bt, xr, xmr, g.silu, gg, xh);
#pragma unroll
    for (int i = 0; i < 4; ++i) { s1[i] += gg[i]; s2[i] = fmaf(gg[i], xh[i], s2[i]); }
    if (two) {
      gnb_piece(fb, db, gm, bt, xr, xmr, g.silu, gg, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s1[i] += gg[i]; s2[i] = fmaf(gg[i], xh[i], s2[i]); }
    }
  }
  const int T = blockDim.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[i * T + threadIdx.x] = s1[i]; red[(4 + i) * T + threadIdx.x] = s2[i]; }
  __syncthreads();
  for (int t = threadIdx.x; t < g.C * 2; t += T) {      // consecutive threads -> consecutive (channel, which) doubles
    const int c = t >> 1, which = t & 1;
    const int vv = c >> 2, comp = which * 4 + (c & 3);
    float acc = 0.f;
    for (int rr = 0; rr < g.rows_par; ++rr) acc += red[comp * T + rr * g.vecs + vv];
    atomicAdd(&bsums[((size_t)ns * g.C + c) * 2 + which], (double)acc);
  }
}

// pass 2: dx = rstd * (g - mean(g) - xhat * mean(g xhat))  [+ add], written to up to three destinations
// the `add` tensor and (when accumulating) the destination's old value are read together with x and dy, before the
// arithmetic, so that all of a row's loads are in flight at once
__device__ __forceinline__ float* gnb_dst(const GnbGeom& g, long long row, int c0, float* out1, int acc1, float* out2,
                                          int acc2, int& acc) {
  acc = 0;
  if (c0 < g.C1) {
    if (out1) { acc = acc1; return out1 + row * g.C1 + c0; }
    return nullptr;
  }
  if (out2) { acc = acc2; return out2 + row * g.C2 + (c0 - g.C1); }
  return nullptr;
}
__device__ __forceinline__ void gnb_load_extra(const GnbGeom& g, long long row, int c0, const void* __restrict__ add,
                                               int add_f32, const float* dst, int acc, float (&e)[4]) {
  e[0] = e[1] = e[2] = e[3] = 0.f;
  if (add != nullptr) ld4_any(add, row * g.C + c0, add_f32, e);
  if (dst != nullptr && acc) {
    const float4 t = *reinterpret_cast<const float4*>(dst);
    e[0] += t.x; e[1] += t.y; e[2] += t.z; e[3] += t.w;
  }
}
__device__ __forceinline__ void gnb_store(const GnbGeom& g, long long row, int c0, const float (&o)[4], float* dst,
                                          __nv_bfloat16* __restrict__ out_bf16) {
  if (dst != nullptr) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
  if (out_bf16 != nullptr)
    *reinterpret_cast<uint2*>(out_bf16 + row * g.C + c0) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
}

// 64 registers (four CTAs per SM) with a 48-byte spill measured faster at level 0 than 92 registers without (44 vs 61 us)
__global__ void __launch_bounds__(1024) gn_bwd_apply_kernel(const void* __restrict__ x1, const void* __restrict__ x2,
                                    const __nv_bfloat16* __restrict__ dy, GnbGeom g, const double* __restrict__ fsums,
                                    const double* __restrict__ bsums, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const void* __restrict__ add, int add_f32,
                                    float* __restrict__ out1, int acc1, float* __restrict__ out2, int acc2,
                                    __nv_bfloat16* __restrict__ out_bf16) {
  extern __shared__ float sh[];
  float* gmean = sh;
  float* grstd = gmean + g.groups;
  float* m1 = grstd + g.groups;    // [groups] mean(g)
  float* m2 = m1 + g.groups;       // [groups] mean(g * xhat)
  const int ns = blockIdx.y;
  const int cpg = g.C / g.groups;
  gnb_prologue(g, ns, fsums, bsums, gmean, grstd, m1, m2);
  const int v = threadIdx.x % g.vecs, rl = threadIdx.x / g.vecs;
  const int c0 = v * 4;
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.R);
  float gm[4], bt[4], xr[4], xmr[4], k1[4], k2[4];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0));
    gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w;
    bt[0] = b.x; bt[1] = b.y; bt[2] = b.z; bt[3] = b.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = (c0 + i) / cpg;
    xr[i] = grstd[gi]; xmr[i] = -gmean[gi] * grstd[gi];
    k1[i] = m1[gi]; k2[i] = m2[gi];
  }
  for (int r = r0 + rl; r < r1; r += 2 * g.rows_par) {
    const long long row = (long long)ns * g.R + r;
    const bool two = r + g.rows_par < r1;
    const long long rowb = two ? row + g.rows_par : row;
    float fa[4], da[4], fb[4], db[4], ea[4], eb[4], gg[4], xh[4], o[4];
    int acca, accb;
    float* dsta = gnb_dst(g, row, c0, out1, acc1, out2, acc2, acca);
    float* dstb = gnb_dst(g, rowb, c0, out1, acc1, out2, acc2, accb);
    gnb_load_x(x1, x2, g, row, c0, fa);
    ld4_any(dy, row * g.C + c0, 0, da);
    gnb_load_extra(g, row, c0, add, add_f32, dsta, acca, ea);
    gnb_load_x(x1, x2, g, rowb, c0, fb);
    ld4_any(dy, rowb * g.C + c0, 0, db);
    if (two) gnb_load_extra(g, rowb, c0, add, add_f32, dstb, accb, eb);
    gnb_piece(fa, da, gm, bt, xr, xmr, g.silu, gg, xh);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fmaf(xr[i], gg[i] - k1[i] - xh[i] * k2[i], ea[i]);
    gnb_store(g, row, c0, o, dsta, out_bf16);
    if (two) {
      gnb_piece(fb, db, gm, bt, xr, xmr, g.silu, gg, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaf(xr[i], gg[i] - k1[i] - xh[i] * k2[i], eb[i]);
      gnb_store(g, rowb, c0, o, dstb, out_bf16);
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm backward
// one warp per row; G += rstd * (g - mean(g) - xhat * mean(g xhat)),  g = dy * gamma;  optional bf16 copy of the new G
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const void* __restrict__ dy,
                                                     int dy_f32, int M, int C, const float* __restrict__ gamma,
                                                     float eps, float* __restrict__ G, int accumulate,
                                                     __nv_bfloat16* __restrict__ g_bf16) {
  const int lane = threadIdx.x & 31;
  const int nvec = C / 8;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float f[NV][8], d[NV][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    if (v < nvec) {
      ld8_f32(x + row * C + v * 8, f[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[j][i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float t = f[j][i] - mean; q = fmaf(t, t, q); }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    if (v < nvec) {
      float gm[8];
      ld8_f32(gamma + v * 8, gm);
      ld8_any(dy, row * C + v * 8, dy_f32, d[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[j][i] = (f[j][i] - mean) * rstd;       // xhat
        d[j][i] *= gm[i];                        // g
        s1 += d[j][i];
        s2 = fmaf(d[j][i], f[j][i], s2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 /= C; s2 /= C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + 32 * j;
    if (v < nvec) {
      float o8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = rstd * (d[j][i] - s1 - f[j][i] * s2);
      float* gp = G + row * C + v * 8;
      if (accumulate) {
        float e[8];
        ld8_f32(gp, e);
#pragma unroll
        for (int i = 0; i < 8; ++i) o8[i] += e[i];
      }
      st8_f32(gp, o8);
      if (g_bf16) st8_bf16(g_bf16 + row * C + v * 8, o8);
    }
  }
}

// ------------------------------------------------------------------------------------------- GEGLU
// pre: [M, 2H] bf16 in the GEMM's tile-interleaved order: tile t holds 128 value columns then their 128 gate columns.
// Phi(-t), t >= 0, as 2^P(t) with ONE MUFU op: the degree-6 fit of log2(0.5 erfc(t / sqrt 2)) on [0, 6] that the GEMM's
// GEGLU epilogue uses (gemm_epilogue.cuh `gelu_erf_fast`, tools/fit_gelu.py: 4.8e-5 relative).  With erff both kernels
// were instruction-bound (ncu: issue slots 83 / 85 % busy at 3.6 / 4.8 TB/s).
__device__ __forceinline__ float phi_neg_fast(float t) {
  t = fminf(t, 6.0f);
  float p = fmaf(2.2999249e-05f, t, -6.1149016e-04f);
  p = fmaf(p, t, 7.2001889e-03f);
  p = fmaf(p, t, -5.1208213e-02f);
  p = fmaf(p, t, -4.6122226e-01f);
  p = fmaf(p, t, -1.1502144e+00f);
  p = fmaf(p, t, -1.0000589e+00f);
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(p));
  return y;
}
__global__ void geglu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ out, long long M,
                                 int H) {
  const int hv = H / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * hv) return;
  const long long m = idx / hv;
  const int j = (int)(idx % hv) * 8;            // output column
  const int t = j >> 7, jj = j & 127;
  const __nv_bfloat16* p = pre + m * 2 * H + t * 256 + jj;
  float a[8], gt[8], o[8];
  ld8_bf16(p, a);
  ld8_bf16(p + 128, gt);
#pragma unroll
  for (int i = 0; i < 8; ++i) {      // gelu(x) = relu(x) - |x| Phi(-|x|)
    const float ax = fabsf(gt[i]);
    o[i] = a[i] * fmaf(-ax, phi_neg_fast(ax), fmaxf(gt[i], 0.f));
  }
  st8_bf16(out + m * H + j, o);
}

__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dout,
                                 __nv_bfloat16* __restrict__ dpre, long long M, int H) {
  const int hv = H / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * hv) return;
  const long long m = idx / hv;
  const int j = (int)(idx % hv) * 8;
  const int t = j >> 7, jj = j & 127;
  const long long off = m * 2 * H + t * 256 + jj;
  float a[8], gt[8], d[8], da[8], dg[8];
  ld8_bf16(pre + off, a);
  ld8_bf16(pre + off + 128, gt);
  ld8_bf16(dout + m * H + j, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {      // one Phi per element serves gelu (x * cdf) and its derivative (cdf + x * pdf)
    const float hneg = phi_neg_fast(fabsf(gt[i]));
    const float cdf = gt[i] >= 0.f ? 1.0f - hneg : hneg;
    da[i] = d[i] * (gt[i] * cdf);
    dg[i] = d[i] * a[i] * fmaf(gt[i] * 0.3989422804014327f, __expf(-0.5f * gt[i] * gt[i]), cdf);
  }
  st8_bf16(dpre + off, da);
  st8_bf16(dpre + off + 128, dg);
}

// ------------------------------------------------------------------------------------------- grouped column sums
__device__ __forceinline__ int rv_index(int mode, long long m64, int HW, int F, int B) {
  const unsigned m = (unsigned)m64, hw = (unsigned)HW, hwf = (unsigned)HW * (unsigned)F;   // rows < 2^31 (GEMM M is int32)
  switch (mode) {
    case LKGD_RV_FRAME: return (int)(m / hw);
    case LKGD_RV_FRAMEPOS: return (int)((m / hw) % (unsigned)F);
    case LKGD_RV_BATCH: return (int)(m / hwf);
    case LKGD_RV_TCTX_0272: return (int)(((m / hwf) * hw + (m % hw)) % (unsigned)B);
    default: return 0;
  }
}

constexpr int CS_MAXG = 8;
// grid (column blocks of 128, row chunks); block = 128 columns x 4 row lanes; each thread walks its rows with 32-bit
// index arithmetic (frame / pixel counters advanced incrementally), per-group partial sums reduced through smem, one
// atomic per (group, column) per CTA
__global__ void __launch_bounds__(512) colsum_grouped_kernel(const float* __restrict__ G, long long M, int C,
                                                             int rows_per_cta, int mode, int HW, int F, int B,
                                                             int n_groups, float* __restrict__ out, long long ldo) {
  __shared__ float red[4][CS_MAXG][128];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int c = blockIdx.x * 128 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float acc[CS_MAXG];
#pragma unroll
  for (int i = 0; i < CS_MAXG; ++i) acc[i] = 0.f;
  if (c < C) {
    for (long long r = r0 + ty; r < r1; r += 4) {
      const int gi = rv_index(mode, r, HW, F, B);
      const float v = __ldg(G + r * C + c);
#pragma unroll
      for (int i = 0; i < CS_MAXG; ++i) acc[i] += (i == gi) ? v : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < CS_MAXG; ++i) red[ty][i][tx] = acc[i];
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 0; i < n_groups; ++i) {
      const float s = red[0][i][tx] + red[1][i][tx] + red[2][i][tx] + red[3][i][tx];
      if (s != 0.f) atomicAdd(out + (size_t)i * ldo + c, s);
    }
  }
}

// One group (batch 1 per GPU, the reference's training configuration: every row belongs to the same context vector):
// no per-row group arithmetic, 16-byte loads, four rows in flight per thread; block = 32 column quads x 16 row lanes.
__global__ void __launch_bounds__(512) colsum_all_kernel(const float* __restrict__ G, long long M, int C,
                                                         int rows_per_cta, float* __restrict__ out) {
  __shared__ float4 red[16][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, M);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (c < C) {
    const float* base = G + c;
    for (long long r = r0 + ty; r < r1; r += 64) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + 16 * u;
        v[u] = rr < r1 ? __ldg(reinterpret_cast<const float4*>(base + rr * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      a.x += v[0].x + v[2].x; a.y += v[0].y + v[2].y; a.z += v[0].z + v[2].z; a.w += v[0].w + v[2].w;
      b.x += v[1].x + v[3].x; b.y += v[1].y + v[3].y; b.z += v[1].z + v[3].z; b.w += v[1].w + v[3].w;
    }
  }
  red[ty][tx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  __syncthreads();
  if (ty == 0 && c < C) {
    float4 s = red[0][tx];
#pragma unroll
    for (int i = 1; i < 16; ++i) { s.x += red[i][tx].x; s.y += red[i][tx].y; s.z += red[i][tx].z; s.w += red[i][tx].w; }
    atomicAdd(out + c, s.x); atomicAdd(out + c + 1, s.y); atomicAdd(out + c + 2, s.z); atomicAdd(out + c + 3, s.w);
  }
}

// ------------------------------------------------------------------------------------------- resampling gradients
// nearest-2x upsample backward: out[n,h,w,:] = sum of the 2x2 block of in [N,2H,2W,C]
__global__ void downsum2x_kernel(const void* __restrict__ in, int in_f32, float* __restrict__ out, int N, int H, int W,
                                 int C) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * H * W * cv) return;
  const int v = (int)(idx % cv);
  const long long pix = idx / cv;
  const int w = (int)(pix % W), h = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float f[8];
      ld8_any(in, ((n * 2 * H + 2 * h + dy) * 2 * W + 2 * w + dx) * C + v * 8, in_f32, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += f[i];
    }
  st8_f32(out + pix * C + v * 8, o);
}

// stride-2 conv data gradient: out [N,Hin,Win,C] bf16 with out[2ho,2wo] = in[ho,wo], zero elsewhere
__global__ void zero_stuff2x_kernel(const void* __restrict__ in, int in_f32, __nv_bfloat16* __restrict__ out, int N,
                                    int Hin, int Win, int Ho, int Wo, int C) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * Hin * Win * cv) return;
  const int v = (int)(idx % cv);
  const long long pix = idx / cv;
  const int w = (int)(pix % Win), h = (int)((pix / Win) % Hin);
  const long long n = pix / ((long long)Win * Hin);
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (!(h & 1) && !(w & 1) && (h >> 1) < Ho && (w >> 1) < Wo)
    ld8_any(in, ((n * Ho + (h >> 1)) * Wo + (w >> 1)) * C + v * 8, in_f32, f);
  st8_bf16(out + pix * C + v * 8, f);
}

// ------------------------------------------------------------------------------------------- EDM wrapper + loss
// noisy = latents + noise * sigma[b];  x_in = [noisy / sqrt(sigma^2 + 1) | cond[b] | 0...]   (train_svd_lora.py:1503-1530)
__global__ void edm_precondition_kernel(const float* __restrict__ lat, const float* __restrict__ noise,
                                        const float* __restrict__ sigma, const float* __restrict__ cond,
                                        float* __restrict__ noisy, __nv_bfloat16* __restrict__ xin, int B, int F, int C,
                                        int HW, int Cpad) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * F * HW) return;
  const int p = (int)(pix % HW);
  const int f = (int)((pix / HW) % F);
  const int b = (int)(pix / ((long long)HW * F));
  const float sg = sigma[b];
  const float cin = rsqrtf(sg * sg + 1.0f);
  __nv_bfloat16* o = xin + pix * Cpad;
  for (int c8 = 0; c8 < Cpad; c8 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c8 + i;
      float val = 0.f;
      if (c < C) {
        const size_t e = (((size_t)b * F + f) * C + c) * HW + p;
        const float nz = fmaf(noise[e], sg, lat[e]);
        noisy[e] = nz;
        val = nz * cin;
      } else if (c < 2 * C) {
        val = cond[((size_t)b * C + (c - C)) * HW + p];
      }
      v[i] = val;
    }
    st8_bf16(o + c8, v);
  }
}

// denoised = pred * c_out + c_skip * noisy;  loss = mean_b mean_{f,c,h,w} w(sigma) (denoised - target)^2
// (train_svd_lora.py:1651-1672); dpred = dloss/dpred as channels-last bf16 rows [B*F*HW, Cpad] (zero padded)
__global__ void edm_loss_kernel(const float* __restrict__ pred, int ld, const float* __restrict__ noisy,
                                const float* __restrict__ target, const float* __restrict__ sigma,
                                double* __restrict__ loss, __nv_bfloat16* __restrict__ dpred, int B, int F, int C, int HW,
                                int Cpad, float grad_scale) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double part = 0.0;
  if (pix < (long long)B * F * HW) {
    const int p = (int)(pix % HW);
    const int f = (int)((pix / HW) % F);
    const int b = (int)(pix / ((long long)HW * F));
    const float sg = sigma[b];
    const float c_out = -sg * rsqrtf(sg * sg + 1.0f), c_skip = 1.0f / (sg * sg + 1.0f);
    const float wgt = (1.0f + sg * sg) / (sg * sg);
    const float inv_n = 1.0f / ((float)F * C * HW * B);
    for (int c8 = 0; c8 < Cpad; c8 += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c8 + i;
        float gval = 0.f;
        if (c < C) {
          const size_t e = (((size_t)b * F + f) * C + c) * HW + p;
          const float diff = fmaf(pred[pix * ld + c], c_out, c_skip * noisy[e]) - target[e];
          part += (double)(wgt * diff * diff * inv_n);
          gval = 2.0f * wgt * diff * c_out * inv_n * grad_scale;
        }
        v[i] = gval;
      }
      if (dpred) st8_bf16(dpred + pix * Cpad + c8, v);
    }
  }
  // block reduction -> one atomic per block
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(loss, s);
  }
}

// ------------------------------------------------------------------------------------------- optimizer
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  double part = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    part += (double)x[i] * x[i];
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(out, s);
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction); the gradient is first multiplied by
// grad_scale * min(1, max_norm / (sqrt(*sumsq) * grad_scale + 1e-6))  (clip_grad_norm_; sumsq may be NULL)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps, float wd,
                             float bc1, float bc2, float grad_scale, const double* __restrict__ sumsq, float max_norm) {
  float gs = grad_scale;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float norm = (float)sqrt(*sumsq) * grad_scale;
    const float clip = max_norm / (norm + 1e-6f);
    if (clip < 1.0f) gs *= clip;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
    m[i] = mi; v[i] = vi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    p[i] = pi;
  }
}

// contiguous-row fast path: 8 elements per thread, 128-bit loads / stores
__global__ void cast2d_bf16_vec_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                       long long ldd, int rows, int cols, float alpha) {
  const int cv = cols / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cv) return;
  const long long r = idx / cv;
  const int c = (int)(idx % cv) * 8;
  float f[8];
  ld8_f32(src + r * lds + c, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] *= alpha;
  st8_bf16(dst + r * ldd + c, f);
}

__global__ void cast2d_bf16_kernel(const float* __restrict__ src, long long lds, long long cs,
                                   __nv_bfloat16* __restrict__ dst, long long ldd, int rows, int cols, float alpha) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const long long r = idx / cols;
  const int c = (int)(idx % cols);
  dst[r * ldd + c] = __float2bfloat16(src[r * lds + c * cs] * alpha);
}

// one launch for a table of casts (blockIdx.y = job): the LoRA repack after the optimizer step
__global__ void cast2d_bf16_batch_kernel(const lkgd_cast2d_job* __restrict__ jobs) {
  const lkgd_cast2d_job j = jobs[blockIdx.y];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)j.rows * j.cols) return;
  const long long r = idx / j.cols;
  const int c = (int)(idx % j.cols);
  reinterpret_cast<__nv_bfloat16*>(j.dst)[r * j.ldd + c] = __float2bfloat16(j.src[r * j.lds + c * j.src_cs] * j.alpha);
}

// ------------------------------------------------------------------------------------------- small fp32 backward
// (latent-knowledge conditioning block, reference models/unet_spatio_temporal_condition.py:536-595 under autograd)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {       // derivative expressed with the OUTPUT y
  if (act == 3) return y > 0.f ? 1.0f : 0.1f;                               // LeakyReLU(0.1)
  return 1.0f;
}
// dx[m, k] = sum_n dy[m, n] * act'(y[m, n]) * W[n, k]
__global__ void small_linear_bwd_x_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                          int act_out, const float* __restrict__ W, float* __restrict__ dx, int lddx,
                                          int M, int N, int K, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (k >= K) return;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    float g = dy[(size_t)m * lddy + n];
    if (act_out) g *= act_grad_from_out(y[(size_t)m * ldy + n], act_out);
    acc = fmaf(g, W[(size_t)n * K + k], acc);
  }
  float* o = dx + (size_t)m * lddx + k;
  *o = accumulate ? *o + acc : acc;
}
// The same product for long contractions (N >= 512: the concatenated cross-attention matrices, [sum C, 1024] = 56 MB at
// SVD width): a CTA owns 8 output columns and all of N, each thread reads one 32-byte sector of W per n (whole sectors,
// K / 8 CTAs in flight instead of K / 128), block reduction in a fixed order (no atomics: bit-reproducible).
__global__ void __launch_bounds__(256) small_linear_bwd_x_wide_kernel(
    const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, int act_out,
    const float* __restrict__ W, float* __restrict__ dx, int lddx, int N, int K, int accumulate) {
  __shared__ float red[8][8];
  const int k0 = blockIdx.x * 8;
  const int m = blockIdx.y;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int n = threadIdx.x; n < N; n += 256) {
    float g = dy[(size_t)m * lddy + n];
    if (act_out) g *= act_grad_from_out(y[(size_t)m * ldy + n], act_out);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k0 + 4));
    acc[0] = fmaf(g, w0.x, acc[0]); acc[1] = fmaf(g, w0.y, acc[1]);
    acc[2] = fmaf(g, w0.z, acc[2]); acc[3] = fmaf(g, w0.w, acc[3]);
    acc[4] = fmaf(g, w1.x, acc[4]); acc[5] = fmaf(g, w1.y, acc[5]);
    acc[6] = fmaf(g, w1.z, acc[6]); acc[7] = fmaf(g, w1.w, acc[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x >> 5][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    float* o = dx + (size_t)m * lddx + k0 + threadIdx.x;
    *o = accumulate ? *o + s : s;
  }
}
// dW[n, k] += sum_m dy'[m, n] * act_in(x[m, k]);  db[n] += sum_m dy'[m, n]
__global__ void small_linear_bwd_w_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                          int act_out, const float* __restrict__ x, int ldx, float* __restrict__ dW,
                                          float* __restrict__ db, int M, int N, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (k >= K) return;
  float acc = 0.f, bacc = 0.f;
  for (int m = 0; m < M; ++m) {
    float g = dy[(size_t)m * lddy + n];
    if (act_out) g *= act_grad_from_out(y[(size_t)m * ldy + n], act_out);
    acc = fmaf(g, x[(size_t)m * ldx + k], acc);
    bacc += g;
  }
  if (dW) dW[(size_t)n * K + k] += acc;
  if (db && k == 0) db[n] += bacc;
}
// mode 0: forward (re, im) -> (mag, pha); given d mag, d pha -> d re, d im.
// mode 1: forward (mag, pha) -> (re, im) = mag (cos, sin) pha; given d re, d im -> d mag, d pha.
__global__ void polar_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ d0,
                                 const float* __restrict__ d1, float* __restrict__ o0, float* __restrict__ o1, int n,
                                 int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) {
    const float re = a[i], im = b[i];
    const float m2 = re * re + im * im;
    if (m2 > 0.f) {
      const float inv = rsqrtf(m2);
      o0[i] = d0[i] * re * inv - d1[i] * im / m2;
      o1[i] = d0[i] * im * inv + d1[i] * re / m2;
    } else {
      o0[i] = 0.f; o1[i] = 0.f;
    }
  } else {
    const float mag = a[i], pha = b[i];
    float sn, cs;
    sincosf(pha, &sn, &cs);
    o0[i] = d0[i] * cs + d1[i] * sn;
    o1[i] = mag * (d1[i] * cs - d0[i] * sn);
  }
}
// Conv1d(4G -> G, k=1, groups=G) weight gradient: dw[j, m] += sum_b dy[b, j] * x[b, 4 j + m]
__global__ void grouped1x1_bwd_w_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                        float* __restrict__ dw, int B, int G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * 4) return;
  const int j = idx >> 2;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc = fmaf(dy[(size_t)b * lddy + j], x[(size_t)b * ldx + idx], acc);
  dw[idx] += acc;
}
// Hamilton-product weight: dWt is the dense gradient [out, in] of y = x @ W (W [in, out] assembled from r,i,j,k blocks
// [in/4, out/4]); folds it back onto the four quaternion components (+=).
__global__ void hamilton_bwd_kernel(const float* __restrict__ dWt, int in, int out, float* __restrict__ dr,
                                    float* __restrict__ di, float* __restrict__ dj, float* __restrict__ dk) {
  const int i4 = in / 4, o4 = out / 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= i4 * o4) return;
  const int p = idx / o4, q = idx % o4;
  auto D = [&](int a, int b) { return dWt[(size_t)(b * o4 + q) * in + a * i4 + p]; };
  dr[idx] += D(0, 0) + D(1, 1) + D(2, 2) + D(3, 3);
  di[idx] += D(0, 1) - D(1, 0) + D(2, 3) - D(3, 2);
  dj[idx] += D(0, 2) - D(2, 0) + D(3, 1) - D(1, 3);
  dk[idx] += D(0, 3) - D(3, 0) + D(1, 2) - D(2, 1);
}

static GnbGeom gnb_geom(int C1, int C2, int NS, int R, int x_f32, int groups, int silu, float eps) {
  GnbGeom g;
  g.C1 = C1; g.C2 = C2; g.C = C1 + C2; g.vecs = g.C / 4; g.R = R; g.x_f32 = x_f32;
  g.groups = groups; g.silu = silu; g.eps = eps;
  g.rows_par = 256 / g.vecs;
  if (g.rows_par < 1) g.rows_par = 1;
  if (g.rows_par > R) g.rows_par = R;
  // row passes per thread (two rows in flight at a time): about four CTAs per SM in ONE wave - every CTA pays the
  // prologue and 2 C double atomics, and with few samples (temporal norms: NS = batch) all of them hit the same
  // addresses
  int passes;
  if (const char* e = getenv("LKGD_GNB_PASSES")) {       // tuning experiments only
    passes = atoi(e);
    if (passes < 1 || passes > 64) passes = 8;
  } else {
    const long long slots = ((long long)R + g.rows_par - 1) / g.rows_par * NS;
    passes = (int)((slots + 591) / 592);
    passes = (passes + 1) & ~1;
    if (passes < 2) passes = 2;
    if (passes > 32) passes = 32;
  }
  g.rows_per_cta = g.rows_par * passes;
  return g;
}

}  // namespace lkgd

using namespace lkgd;

extern "C" size_t lkgd_groupnorm_bwd_workspace(int32_t NS, int32_t C) { return (size_t)NS * C * 2 * sizeof(double); }

extern "C" int lkgd_groupnorm_bwd(const void* x1, int32_t C1, const void* x2, int32_t C2, int32_t NS, int32_t R,
                                  int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                                  int32_t x_f32, const void* dy, const void* fwd_sums, const void* add, int32_t add_f32,
                                  float* out1, int32_t acc1, float* out2, int32_t acc2, void* out_bf16, void* workspace,
                                  size_t ws_bytes, void* stream) {
  if (x2 == nullptr) C2 = 0;
  const int C = C1 + C2;
  if (NS <= 0 || R <= 0 || C <= 0 || groups <= 0 || C % groups || C1 % 8 || C2 % 8 || C / 4 > 1024) return LKGD_ESHAPE;
  if (dy == nullptr || fwd_sums == nullptr) return LKGD_ESHAPE;
  if (!aligned16(x1) || !aligned16(dy) || (x2 && !aligned16(x2)) || (add && !aligned16(add)) ||
      (out1 && !aligned16(out1)) || (out2 && !aligned16(out2)) || (out_bf16 && !aligned16(out_bf16)))
    return LKGD_EALIGN;
  if (ws_bytes < lkgd_groupnorm_bwd_workspace(NS, C) || workspace == nullptr) return LKGD_EWS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnbGeom g = gnb_geom(C1, C2, NS, R, x_f32, groups, silu, eps);
  cudaError_t e = cudaMemsetAsync(workspace, 0, lkgd_groupnorm_bwd_workspace(NS, C), st);
  if (e != cudaSuccess) return set_cuda_error(e);
  dim3 grid((R + g.rows_per_cta - 1) / g.rows_per_cta, NS);
  const int threads = g.vecs * g.rows_par;
  const size_t sh_pro = (size_t)(4 * groups) * sizeof(float);
  const size_t sh1 = (size_t)(2 * groups + threads * 8) * sizeof(float);
  if (sh_pro > 48 * 1024 || sh1 > 48 * 1024) return LKGD_ESHAPE;
  gn_bwd_stats_kernel<<<grid, threads, sh1, st>>>(x1, x2, reinterpret_cast<const __nv_bfloat16*>(dy), g,
                                                  reinterpret_cast<const double*>(fwd_sums), gamma, beta,
                                                  reinterpret_cast<double*>(workspace));
  int rc = launch_epilogue();
  if (rc) return rc;
  gn_bwd_apply_kernel<<<grid, threads, sh_pro, st>>>(x1, x2, reinterpret_cast<const __nv_bfloat16*>(dy), g,
                                                     reinterpret_cast<const double*>(fwd_sums),
                                                     reinterpret_cast<const double*>(workspace), gamma, beta, add,
                                                     add_f32, out1, acc1, out2, acc2,
                                                     reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return launch_epilogue();
}

extern "C" int lkgd_layernorm_bwd(const float* x, const void* dy, int32_t dy_f32, int32_t M, int32_t C,
                                  const float* gamma, float eps, float* G, int32_t accumulate, void* g_bf16,
                                  void* stream) {
  if (M <= 0 || C <= 0 || C % 8 || C > 2048) return LKGD_ESHAPE;
  if (!aligned16(x) || !aligned16(dy) || !aligned16(G) || !aligned16(gamma) || (g_bf16 && !aligned16(g_bf16)))
    return LKGD_EALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (M + 7) / 8;
  const int nv = (C / 8 + 31) / 32;
#define LNB(NV) ln_bwd_kernel<NV><<<grid, 256, 0, st>>>(x, dy, dy_f32, M, C, gamma, eps, G, accumulate, \
                                                        reinterpret_cast<__nv_bfloat16*>(g_bf16))
  switch (nv) {
    case 1: LNB(1); break;
    case 2: LNB(2); break;
    case 3: LNB(3); break;
    case 4: LNB(4); break;
    case 5: LNB(5); break;
    default: LNB(8); break;
  }
#undef LNB
  return launch_epilogue();
}

extern "C" int lkgd_geglu_fwd(const void* pre, void* out, int64_t M, int32_t H, void* stream) {
  if (M <= 0 || H <= 0 || H % 128) return LKGD_ESHAPE;
  if (!aligned16(pre) || !aligned16(out)) return LKGD_EALIGN;
  const long long n = M * (H / 8);
  geglu_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<__nv_bfloat16*>(out), M, H);
  return launch_epilogue();
}

extern "C" int lkgd_geglu_bwd(const void* pre, const void* dout, void* dpre, int64_t M, int32_t H, void* stream) {
  if (M <= 0 || H <= 0 || H % 128) return LKGD_ESHAPE;
  if (!aligned16(pre) || !aligned16(dout) || !aligned16(dpre)) return LKGD_EALIGN;
  const long long n = M * (H / 8);
  geglu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dpre), M, H);
  return launch_epilogue();
}

extern "C" int lkgd_colsum_grouped(const float* G, int64_t M, int32_t C, int32_t rv_mode, int32_t rv_HW, int32_t rv_F,
                                   int32_t rv_B, int32_t n_groups, float* out, int64_t ldo, void* stream) {
  if (M <= 0 || C <= 0 || n_groups <= 0 || n_groups > CS_MAXG || ldo < C) return LKGD_ESHAPE;
  if (rv_HW <= 0) rv_HW = 1;
  if (rv_F <= 0) rv_F = 1;
  if (rv_B <= 0) rv_B = 1;
  // ~4 CTAs per SM; at least 64 rows per CTA so the atomics stay a small share
  const int col_blocks = (C + 127) / 128;
  long long want = (4LL * sm_count() + col_blocks - 1) / col_blocks;
  long long rpc = (M + want - 1) / want;
  if (rpc < 64) rpc = 64;
  const int rows_per_cta = (int)rpc;
  dim3 grid(col_blocks, (unsigned)((M + rows_per_cta - 1) / rows_per_cta));
  if (n_groups == 1 && C % 4 == 0 && aligned16(G)) {
    colsum_all_kernel<<<grid, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(G, M, C, rows_per_cta, out);
    return launch_epilogue();
  }
  colsum_grouped_kernel<<<grid, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(G, M, C, rows_per_cta, rv_mode, rv_HW,
                                                                                   rv_F, rv_B, n_groups, out, ldo);
  return launch_epilogue();
}

extern "C" int lkgd_downsum2x(const void* in, int32_t in_f32, float* out, int32_t N, int32_t H, int32_t W, int32_t C,
                              void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return LKGD_ESHAPE;
  if (!aligned16(in) || !aligned16(out)) return LKGD_EALIGN;
  const long long n = (long long)N * H * W * (C / 8);
  downsum2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, in_f32, out, N,
                                                                                                     H, W, C);
  return launch_epilogue();
}

extern "C" int lkgd_zero_stuff2x(const void* in, int32_t in_f32, void* out, int32_t N, int32_t Hin, int32_t Win,
                                 int32_t C, void* stream) {
  if (N <= 0 || Hin <= 0 || Win <= 0 || C <= 0 || C % 8) return LKGD_ESHAPE;
  if (!aligned16(in) || !aligned16(out)) return LKGD_EALIGN;
  const int Ho = (Hin - 1) / 2 + 1, Wo = (Win - 1) / 2 + 1;
  const long long n = (long long)N * Hin * Win * (C / 8);
  zero_stuff2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, in_f32, reinterpret_cast<__nv_bfloat16*>(out), N, Hin, Win, Ho, Wo, C);
  return launch_epilogue();
}

extern "C" int lkgd_edm_precondition(const float* latents, const float* noise, const float* sigma, const float* cond,
                                     float* noisy, void* x_in, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                                     int32_t Cpad, void* stream) {
  if (B <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad % 8 || Cpad < 2 * C) return LKGD_ESHAPE;
  if (!aligned16(x_in)) return LKGD_EALIGN;
  const long long n = (long long)B * F * H * W;
  edm_precondition_kernel<<<(unsigned)((n + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      latents, noise, sigma, cond, noisy, reinterpret_cast<__nv_bfloat16*>(x_in), B, F, C, H * W, Cpad);
  return launch_epilogue();
}

extern "C" int lkgd_edm_loss(const float* pred, int32_t ld, const float* noisy, const float* target, const float* sigma,
                             double* loss, void* dpred, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                             int32_t Cpad, float grad_scale, void* stream) {
  if (B <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0 || Cpad % 8 || Cpad < C || ld < C) return LKGD_ESHAPE;
  if (dpred && !aligned16(dpred)) return LKGD_EALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(double), st);
  if (e != cudaSuccess) return set_cuda_error(e);
  const long long n = (long long)B * F * H * W;
  edm_loss_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pred, ld, noisy, target, sigma, loss,
                                                               reinterpret_cast<__nv_bfloat16*>(dpred), B, F, C, H * W,
                                                               Cpad, grad_scale);
  return launch_epilogue();
}

extern "C" int lkgd_sumsq(const float* x, int64_t n, double* out, void* stream) {
  if (n <= 0) return LKGD_ESHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
  if (e != cudaSuccess) return set_cuda_error(e);
  long long blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, out);
  return launch_epilogue();
}

extern "C" int lkgd_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int32_t step, float grad_scale, const double* sumsq,
                          float max_norm, void* stream) {
  if (n <= 0 || step <= 0) return LKGD_ESHAPE;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  adamw_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale, sumsq, max_norm);
  return launch_epilogue();
}

extern "C" int lkgd_cast2d_bf16(const float* src, int64_t lds, int64_t src_cs, void* dst, int64_t ldd, int32_t rows,
                                int32_t cols, float alpha, void* stream) {
  if (rows <= 0 || cols <= 0 || ldd < cols || src_cs <= 0) return LKGD_ESHAPE;
  if (src_cs == 1 && cols % 8 == 0 && lds % 4 == 0 && ldd % 8 == 0 && aligned16(src) && aligned16(dst)) {
    const long long nv = (long long)rows * (cols / 8);
    cast2d_bf16_vec_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, alpha);
    return launch_epilogue();
  }
  const long long n = (long long)rows * cols;
  cast2d_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, lds, src_cs, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, alpha);
  return launch_epilogue();
}

extern "C" int lkgd_cast2d_bf16_batch(const lkgd_cast2d_job* jobs, int32_t n_jobs, int32_t max_elems, void* stream) {
  if (jobs == nullptr || n_jobs <= 0 || n_jobs > 65535 || max_elems <= 0) return LKGD_ESHAPE;
  cast2d_bf16_batch_kernel<<<dim3((max_elems + 255) / 256, n_jobs), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(jobs);
  return launch_epilogue();
}

extern "C" int lkgd_small_linear_bwd(const float* dy, int32_t lddy, const float* y, int32_t ldy, int32_t act_out,
                                     const float* x, int32_t ldx, const float* W, float* dx, int32_t lddx,
                                     int32_t dx_accumulate, float* dW, float* db, int32_t M, int32_t N, int32_t K,
                                     void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || dy == nullptr) return LKGD_ESHAPE;
  if (act_out != 0 && act_out != 3) return LKGD_ESHAPE;
  if (act_out && y == nullptr) return LKGD_ESHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = LKGD_OK;
  if (dx != nullptr) {
    if (W == nullptr) return LKGD_ESHAPE;
    if (N >= 512 && K % 8 == 0 && aligned16(W)) {
      small_linear_bwd_x_wide_kernel<<<dim3(K / 8, M), 256, 0, st>>>(dy, lddy, y, ldy, act_out, W, dx, lddx, N, K,
                                                                    dx_accumulate);
    } else {
      dim3 grid((K + 127) / 128, M);
      small_linear_bwd_x_kernel<<<grid, 128, 0, st>>>(dy, lddy, y, ldy, act_out, W, dx, lddx, M, N, K, dx_accumulate);
    }
    if ((rc = launch_epilogue())) return rc;
  }
  if (dW != nullptr || db != nullptr) {
    if (x == nullptr) return LKGD_ESHAPE;
    dim3 grid((K + 127) / 128, N);
    small_linear_bwd_w_kernel<<<grid, 128, 0, st>>>(dy, lddy, y, ldy, act_out, x, ldx, dW, db, M, N, K);
    rc = launch_epilogue();
  }
  return rc;
}

extern "C" int lkgd_polar_bwd(const float* a, const float* b, const float* d0, const float* d1, float* o0, float* o1,
                              int32_t n, int32_t mode, void* stream) {
  if (n <= 0 || (mode != 0 && mode != 1)) return LKGD_ESHAPE;
  polar_bwd_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, d0, d1, o0, o1, n, mode);
  return launch_epilogue();
}

extern "C" int lkgd_grouped1x1_bwd_w(const float* dy, int32_t lddy, const float* x, int32_t ldx, float* dw, int32_t B,
                                     int32_t G, void* stream) {
  if (B <= 0 || G <= 0) return LKGD_ESHAPE;
  grouped1x1_bwd_w_kernel<<<(G * 4 + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, lddy, x, ldx, dw,
                                                                                                  B, G);
  return launch_epilogue();
}

extern "C" int lkgd_hamilton_bwd(const float* dWt, int32_t in, int32_t out, float* dr, float* di, float* dj, float* dk,
                                 void* stream) {
  if (in <= 0 || out <= 0 || in % 4 || out % 4) return LKGD_ESHAPE;
  const int n = (in / 4) * (out / 4);
  hamilton_bwd_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dWt, in, out, dr, di, dj, dk);
  return launch_epilogue();
}
