// Plain SIMT implementations of the GEMM / attention contracts.  TESTS ONLY: they are the on-GPU checkers for
// sizes where the CPU oracle is too slow; nothing in the product path calls them.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

__device__ __forceinline__ int chk_rowvec_index(int mode, long long m, int HW, int F, int B) {
  switch (mode) {
    case LKGD_RV_FRAME: return (int)(m / HW);
    case LKGD_RV_FRAMEPOS: return (int)((m / HW) % F);
    case LKGD_RV_BATCH: return (int)(m / ((long long)HW * F));
    case LKGD_RV_TCTX_0272: return (int)(((m / ((long long)HW * F)) * HW + (m % HW)) % B);
    default: return 0;
  }
}

__device__ float chk_dot(const lkgd_gemm_args& a, long long m, int n) {
  const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(a.A);
  const __nv_bfloat16* Bw = reinterpret_cast<const __nv_bfloat16*>(a.Bw) + (size_t)n * a.ldb;
  float acc = 0.f;
  if (a.a_mode == LKGD_A_LINEAR) {
    const __nv_bfloat16* ar = A + m * a.lda;
    for (int k = 0; k < a.K0; ++k) acc = fmaf(__bfloat162float(ar[k]), __bfloat162float(Bw[k]), acc);
  } else if (a.a_mode == LKGD_A_CONV3X3) {
    const int s = a.stride;
    const int pad = (s == 2 && a.pad_br) ? 0 : 1;
    const int Ho = pad ? (a.Hin - 1) / s + 1 : (a.Hin - 2) / 2 + 1, Wo = pad ? (a.Win - 1) / s + 1 : (a.Win - 2) / 2 + 1;
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho);
    const long long img = m / ((long long)Wo * Ho);
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int hi = ho * s + ky - pad, wi = wo * s + kx - pad;
        if (hi < 0 || hi >= a.Hin || wi < 0 || wi >= a.Win) continue;
        const __nv_bfloat16* ar = A + ((img * a.Hin + hi) * a.Win + wi) * a.K0;
        const __nv_bfloat16* br = Bw + (ky * 3 + kx) * a.K0;
        for (int k = 0; k < a.K0; ++k) acc = fmaf(__bfloat162float(ar[k]), __bfloat162float(br[k]), acc);
      }
  } else {
    const int pp = (int)(m % a.HW), f = (int)((m / a.HW) % a.F);
    const long long b = m / ((long long)a.HW * a.F);
    for (int kt = 0; kt < 3; ++kt) {
      const int fi = f + kt - 1;
      if (fi < 0 || fi >= a.F) continue;
      const __nv_bfloat16* ar = A + ((b * a.F + fi) * a.HW + pp) * a.K0;
      const __nv_bfloat16* br = Bw + kt * a.K0;
      for (int k = 0; k < a.K0; ++k) acc = fmaf(__bfloat162float(ar[k]), __bfloat162float(br[k]), acc);
    }
  }
  if (a.K1 > 0) {
    const __nv_bfloat16* ar = reinterpret_cast<const __nv_bfloat16*>(a.A1) +
                              m * (a.a_mode == LKGD_A_LINEAR ? a.lda1 : a.K1);
    const __nv_bfloat16* br = reinterpret_cast<const __nv_bfloat16*>(a.Bw1) + (size_t)n * a.ldb1;
    for (int k = 0; k < a.K1; ++k) acc = fmaf(__bfloat162float(ar[k]), __bfloat162float(br[k]), acc);
  }
  return acc;
}

__global__ void gemm_check_kernel(const lkgd_gemm_args a) {
  const bool geglu = a.act == LKGD_ACT_GEGLU;
  const int n_cols = geglu ? a.N / 2 : a.N;
  const int n_store = a.n_store > 0 ? a.n_store : n_cols;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.M * n_store) return;
  const long long m = idx / n_store;
  const int n_out = (int)(idx % n_store);
  float v;
  if (geglu) {
    const int tile = n_out / 128, i = n_out % 128;
    const int nh = tile * 256 + i, ng = nh + 128;
    const float h = chk_dot(a, m, nh) + (a.bias ? a.bias[nh] : 0.f);
    const float g = chk_dot(a, m, ng) + (a.bias ? a.bias[ng] : 0.f);
    v = h * gelu_erf_f(g);
  } else {
    v = chk_dot(a, m, n_out) + (a.bias ? a.bias[n_out] : 0.f);
  }
  if (a.rowvec) v += a.rowvec[(size_t)chk_rowvec_index(a.rv_mode, m, max(a.rv_HW, 1), max(a.rv_F, 1), max(a.rv_B, 1)) * (a.rv_ld > 0 ? a.rv_ld : n_cols) + n_out];
  if (a.act == LKGD_ACT_SILU) v = silu_f(v);
  else if (a.act == LKGD_ACT_GELU) v = gelu_erf_f(v);
  else if (a.act == LKGD_ACT_QUICK_GELU) v = v / (1.0f + expf(-1.702f * v));
  v *= a.s0;
  if (a.res1)
    v += a.s1 * (a.res1_f32 ? reinterpret_cast<const float*>(a.res1)[m * a.ldr1 + n_out]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.res1)[m * a.ldr1 + n_out]));
  if (a.res2)
    v += a.s2 * (a.res2_f32 ? reinterpret_cast<const float*>(a.res2)[m * a.ldr2 + n_out]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.res2)[m * a.ldr2 + n_out]));
  if (a.out_f32) reinterpret_cast<float*>(a.out)[m * a.ldo + n_out] = v;
  else reinterpret_cast<__nv_bfloat16*>(a.out)[m * a.ldo + n_out] = __float2bfloat16(v);
  if (a.out2 && a.out_f32) reinterpret_cast<__nv_bfloat16*>(a.out2)[m * a.ldo2 + n_out] = __float2bfloat16(v);
}

// one thread per (image, head, query); two passes over the keys (max, then exp-sum + PV), d <= 128
__global__ void attention_check_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                       const __nv_bfloat16* __restrict__ k, int ldk,
                                       const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ out,
                                       int ldo, int n_img, int heads, int d, int Nq, int Nk, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_img * heads * Nq) return;
  const int i = (int)(idx % Nq);
  const int h = (int)((idx / Nq) % heads);
  const long long img = idx / ((long long)Nq * heads);
  const __nv_bfloat16* qr = q + (img * Nq + i) * ldq + h * d;
  float qf[128];
  for (int t = 0; t < d; ++t) qf[t] = __bfloat162float(qr[t]);
  float mx = -INFINITY;
  for (int j = 0; j < Nk; ++j) {
    const __nv_bfloat16* kr = k + (img * Nk + j) * ldk + h * d;
    float s = 0.f;
    for (int t = 0; t < d; ++t) s = fmaf(qf[t], __bfloat162float(kr[t]), s);
    mx = fmaxf(mx, s * scale);
  }
  float o[128];
  for (int t = 0; t < d; ++t) o[t] = 0.f;
  float l = 0.f;
  for (int j = 0; j < Nk; ++j) {
    const __nv_bfloat16* kr = k + (img * Nk + j) * ldk + h * d;
    float s = 0.f;
    for (int t = 0; t < d; ++t) s = fmaf(qf[t], __bfloat162float(kr[t]), s);
    const float pj = expf(s * scale - mx);
    l += pj;
    const __nv_bfloat16* vr = v + (img * Nk + j) * ldv + h * d;
    for (int t = 0; t < d; ++t) o[t] = fmaf(pj, __bfloat162float(vr[t]), o[t]);
  }
  __nv_bfloat16* orow = out + (img * Nq + i) * ldo + h * d;
  for (int t = 0; t < d; ++t) orow[t] = __float2bfloat16(o[t] / l);
}

}  // namespace lkgd

using namespace lkgd;

extern "C" int lkgd_gemm_simt_check(const lkgd_gemm_args* a, void* stream) {
  if (a == nullptr) return LKGD_ESHAPE;
  const bool geglu = a->act == LKGD_ACT_GEGLU;
  const int n_cols = geglu ? a->N / 2 : a->N;
  const int n_store = a->n_store > 0 ? a->n_store : n_cols;
  const long long total = (long long)a->M * n_store;
  gemm_check_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a);
  return launch_epilogue();
}

extern "C" int lkgd_attention_simt_check(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                                         int32_t ldv, void* out, int32_t ldo, int32_t n_img, int32_t heads,
                                         int32_t d, int32_t Nq, int32_t Nk, float scale, void* stream) {
  if (d > 128 || d <= 0) return LKGD_ESHAPE;
  const long long total = (long long)n_img * heads * Nq;
  attention_check_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k), ldk,
      reinterpret_cast<const __nv_bfloat16*>(v), ldv, reinterpret_cast<__nv_bfloat16*>(out), ldo, n_img, heads, d,
      Nq, Nk, scale);
  return launch_epilogue();
}
