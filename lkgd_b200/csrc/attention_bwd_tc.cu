// Flash-attention backward on tcgen05 / TMEM for 64-wide heads (the SVD head size): the training step's
// attn_bwd_dq_kernel / attn_bwd_dkv_kernel (attention_bwd.cu, mma.sync) re-built on the forward kernel's machinery
// (attention.cu): TMA tiles in SWIZZLE_128B shared memory, accumulators in tensor memory, the bf16 operands P / dS handed
// back to the tensor core THROUGH tensor memory (TS-form MMA), streamed tiles consumed both K-major (score MMAs) and
// MN-major (gradient MMAs) from the same shared-memory tile - no transposes, no atomics.
//
// ONE kernel template serves both directions:
//                         resident 128-row tiles R0, R1     streamed 64-row tiles S0, S1 (ring of two stages)
//   DKV = false  (dQ)     Q, dO  of 128 queries             K_j, V_j
//   DKV = true   (dK,dV)  K, V   of 128 keys                Q_j, dO_j
//   T1 = R0 S0^T   scores (S or S^T)                 128 x 64 fp32   TMEM columns   0.. 63
//   T2 = R1 S1^T   dP or dP^T                        128 x 64 fp32   TMEM columns  64..127
//   P  = 2^(T1 c - lse)          row statistics (dQ) or column statistics (dK,dV: the row is a key, the column a query)
//   dS = P (T2 - Dv)             both packed to bf16 IN PLACE over the first 32 columns of T1 / T2 (each thread owns its row)
//   A0 += dS S0    dQ (x scale) or dK (x scale)      128 x 64 fp32   TMEM columns 128..191   (B = S0 MN-major)
//   A1 += P  S1    dV                  (DKV only)    128 x 64 fp32   TMEM columns 192..255   (B = S1 MN-major)
// 256 TMEM columns and ~66 KB of shared memory per CTA: TWO CTAs per SM - inside a CTA the score MMAs of step j+1 wait for
// the gradient MMAs of step j (P / dS live in the score columns), the other CTA's softmax fills the gap.  The exponentials
// bound this kernel as they bound the forward (16 ex2 / clock / SM): every third pair runs as a polynomial on the FMA pipes.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

constexpr int TB_ROWS = 128, TB_STEP = 64, TB_D = 64;
constexpr int TB_RES = TB_ROWS * TB_D * 2;      // 16 KB resident tile
constexpr int TB_STR = TB_STEP * TB_D * 2;      // 8 KB streamed tile
constexpr int TB_NS = 3;                        // streamed stages
constexpr int TB_THREADS = 192;
constexpr int TB_SMEM = 2 * TB_RES + TB_NS * 2 * TB_STR + 2 * 2 * TB_STEP * 4 + 256;

struct AttnBwdTcParams {
  CUtensorMap tmR0, tmR1, tmS0, tmS1;
  const float* lse;     // [n_img, heads, N] log2-domain log-sum-exp of the scaled scores (forward)
  const float* dvec;    // [n_img, heads, N] sum_d dO * O
  __nv_bfloat16* out0;  // dQ or dK rows (head slice), row pitch ld0
  __nv_bfloat16* out1;  // dV rows (DKV only)
  int ld0, ld1, heads, N;
  float scale, scale_log2;
};

__device__ __forceinline__ float tb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (attention.cu::at_exp2_poly2, scalar form): Cody-Waite split + degree-3 minimax, 7.5e-5 relative
__device__ __forceinline__ float tb_ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float n = t - 12582912.0f;
  const float r = x - n;
  float p = fmaf(0.0551716685f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// ---- packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2: one issue slot for two elements), as in attention.cu
__device__ __forceinline__ uint64_t tb_pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void tb_upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t tb_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t tb_add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t tb_mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t tb_ex2_poly2(uint64_t X) {
  float x0, x1;
  tb_upk2(X, x0, x1);
  X = tb_pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t MAGIC = tb_pk2(12582912.0f, 12582912.0f), NMAGIC = tb_pk2(-12582912.0f, -12582912.0f);
  const uint64_t T = tb_add2(X, MAGIC);
  const uint64_t N = tb_add2(T, NMAGIC);
  const uint64_t R = tb_fma2(N, tb_pk2(-1.0f, -1.0f), X);
  uint64_t P = tb_fma2(tb_pk2(0.0551716685f, 0.0551716685f), R, tb_pk2(0.2426111251f, 0.2426111251f));
  P = tb_fma2(P, R, tb_pk2(0.6932609677f, 0.6932609677f));
  P = tb_fma2(P, R, tb_pk2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  tb_upk2(T, t0, t1);
  tb_upk2(P, p0, p1);
  return tb_pk2(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)),
                __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}

template <bool DKV>
__global__ void __launch_bounds__(TB_THREADS, 2) attn_bwd_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR0 = smem;
  uint8_t* sR1 = smem + TB_RES;
  uint8_t* sS = sR1 + TB_RES;                                 // [stage][S0 | S1]
  float* s_stat = reinterpret_cast<float*>(sS + TB_NS * 2 * TB_STR);     // DKV: [2][lse 64 | Dv 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 2 * TB_STEP);
  uint64_t* r_full = bars;
  uint64_t* st_full = bars + 1;            // [TB_NS]
  uint64_t* st_empty = bars + 1 + TB_NS;   // [TB_NS]
  uint64_t* t_full = bars + 1 + 2 * TB_NS;   // T1, T2 of this step are in TMEM
  uint64_t* ds_full = t_full + 1;            // P / dS of this step are in TMEM
  uint64_t* acc_done = t_full + 2;           // the gradient MMAs of this step have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * TB_ROWS, head = blockIdx.y, img = blockIdx.z;
  const int T = (p.N + TB_STEP - 1) / TB_STEP;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(r_full, 1);
      for (int i = 0; i < TB_NS; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
      mbar_init(t_full, 1);
      mbar_init(ds_full, 128);
      mbar_init(acc_done, 1);
      fence_barrier_init();
      tma_prefetch_desc(&p.tmR0); tma_prefetch_desc(&p.tmR1); tma_prefetch_desc(&p.tmS0); tma_prefetch_desc(&p.tmS1);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_T1 = tmem_base, tmem_T2 = tmem_base + 64, tmem_A0 = tmem_base + 128, tmem_A1 = tmem_base + 192;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(r_full, 2 * TB_RES);
      tma_load_4d(sR0, &p.tmR0, r_full, 0, head, r0, img);
      tma_load_4d(sR1, &p.tmR1, r_full, 0, head, r0, img);
      int st = 0, ph = 1;
      for (int j = 0; j < T; ++j) {
        while (!mbar_try_wait(&st_empty[st], ph)) __nanosleep(64);
        mbar_expect_tx(&st_full[st], 2 * TB_STR);
        tma_load_4d(sS + st * 2 * TB_STR, &p.tmS0, &st_full[st], 0, head, j * TB_STEP, img);
        tma_load_4d(sS + st * 2 * TB_STR + TB_STR, &p.tmS1, &st_full[st], 0, head, j * TB_STEP, img);
        if (++st == TB_NS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc_t = umma_idesc_bf16(TB_STEP);               // N = 64, A and B K-major
      const uint32_t idesc_g = umma_idesc_bf16(TB_D, 128, 0, 1);       // N = 64 head channels, B MN-major
      const uint32_t aR0 = smem_u32(sR0), aR1 = smem_u32(sR1), aS = smem_u32(sS);
      mbar_wait(r_full, 0);
      int st = 0, ph = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&st_full[st], ph);
        tc_fence_after();
        const uint32_t s0 = aS + st * 2 * TB_STR, s1 = s0 + TB_STR;
        // score MMAs (in issue order after the gradient MMAs of step j-1, which read P / dS from these columns)
        {
          const uint64_t a0 = umma_desc_sw128(aR0), b0 = umma_desc_sw128(s0);
          const uint64_t a1 = umma_desc_sw128(aR1), b1 = umma_desc_sw128(s1);
#pragma unroll
          for (int k = 0; k < TB_D / 16; ++k) umma_bf16(tmem_T1, a0 + 2 * k, b0 + 2 * k, idesc_t, k != 0);
#pragma unroll
          for (int k = 0; k < TB_D / 16; ++k) umma_bf16(tmem_T2, a1 + 2 * k, b1 + 2 * k, idesc_t, k != 0);
        }
        umma_commit(t_full);
        mbar_wait(ds_full, j & 1);            // P / dS (bf16) are in the score columns
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TB_STEP / 16; ++ks) {
          umma_bf16_ts(tmem_A0, tmem_T2 + ks * 8, umma_desc_sw128(s0 + ks * 2048), idesc_g, (j | ks) != 0);     // dS S0
          if (DKV) umma_bf16_ts(tmem_A1, tmem_T1 + ks * 8, umma_desc_sw128(s1 + ks * 2048), idesc_g, (j | ks) != 0);   // P S1
        }
        umma_commit(acc_done);
        umma_commit(&st_empty[st]);
        if (++st == TB_NS) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax-side warps: one row per thread
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const size_t stat_base = ((size_t)img * p.heads + head) * p.N;
    float lse_r = INFINITY, dv_r = 0.f;
    if (!DKV && r0 + r < p.N) { lse_r = p.lse[stat_base + r0 + r]; dv_r = p.dvec[stat_base + r0 + r]; }
    const float sc = p.scale_log2;
    for (int j = 0; j < T; ++j) {
      float* st_ = s_stat + (j & 1) * 2 * TB_STEP;
      if (DKV) {
        // NEGATED statistics of the 64 queries of this step: thread t < 64 fetches -lse, t >= 64 fetches -Dv (out of range:
        // -inf / 0, so that 2^(-inf) = 0 removes the query)
        const int q = j * TB_STEP + (r & 63);
        float v = r < 64 ? -INFINITY : 0.f;
        if (q < p.N) v = r < 64 ? -p.lse[stat_base + q] : -p.dvec[stat_base + q];
        st_[r] = v;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(t_full, j & 1);
      tc_fence_after();
      const int valid = p.N - j * TB_STEP;       // dQ: keys of this step inside the sequence
      // two halves of 32 columns: the packed bf16 results of a half overwrite score columns this thread has already read
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t s[32], dp[32];
        tmem_ld32(tmem_T1 + lane_addr + hf * 32, s);
        tmem_ld32(tmem_T2 + lane_addr + hf * 32, dp);
        tmem_ld_wait();
        const uint64_t SC2 = tb_pk2(sc, sc);
        const uint64_t NL2r = tb_pk2(-lse_r, -lse_r), ND2r = tb_pk2(-dv_r, -dv_r);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = hf * 32 + i;
          uint64_t NL2 = NL2r, ND2 = ND2r;
          if (DKV) {
            NL2 = *reinterpret_cast<const uint64_t*>(st_ + col);
            ND2 = *reinterpret_cast<const uint64_t*>(st_ + TB_STEP + col);
          }
          const uint64_t X = tb_fma2(tb_pk2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), SC2, NL2);
          uint64_t P;
          if ((i >> 1) % 3 == 2) {
            P = tb_ex2_poly2(X);
          } else {
            float x0, x1;
            tb_upk2(X, x0, x1);
            P = tb_pk2(tb_ex2(x0), tb_ex2(x1));
          }
          if (!DKV && valid < TB_STEP) {             // zero-filled keys beyond the sequence would still give 2^(-lse)
            float p0, p1;
            tb_upk2(P, p0, p1);
            P = tb_pk2(col < valid ? p0 : 0.f, col + 1 < valid ? p1 : 0.f);
          }
          const uint64_t G = tb_mul2(P, tb_add2(tb_pk2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), ND2));
          float p0, p1, g0, g1;
          tb_upk2(P, p0, p1);
          tb_upk2(G, g0, g1);
          s[i >> 1] = pack_bf16x2(p0, p1);
          dp[i >> 1] = pack_bf16x2(g0, g1);
        }
        tmem_st16(tmem_T2 + lane_addr + hf * 16, *reinterpret_cast<uint32_t(*)[16]>(&dp[0]));
        if (DKV) tmem_st16(tmem_T1 + lane_addr + hf * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_done, (T - 1) & 1);
    tc_fence_after();
    const bool row_ok = r0 + r < p.N;
#pragma unroll
    for (int which = 0; which < (DKV ? 2 : 1); ++which) {
      __nv_bfloat16* orow = (which == 0 ? p.out0 : p.out1) + ((size_t)img * p.N + r0 + r) * (which == 0 ? p.ld0 : p.ld1) +
                            head * TB_D;
      const float mul = which == 0 ? p.scale : 1.0f;
#pragma unroll
      for (int c = 0; c < TB_D; c += 32) {
        uint32_t t[32];
        tmem_ld32((which == 0 ? tmem_A0 : tmem_A1) + lane_addr + c, t);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(orow + c + i) =
                make_uint4(pack_bf16x2(__uint_as_float(t[i]) * mul, __uint_as_float(t[i + 1]) * mul),
                           pack_bf16x2(__uint_as_float(t[i + 2]) * mul, __uint_as_float(t[i + 3]) * mul),
                           pack_bf16x2(__uint_as_float(t[i + 4]) * mul, __uint_as_float(t[i + 5]) * mul),
                           pack_bf16x2(__uint_as_float(t[i + 6]) * mul, __uint_as_float(t[i + 7]) * mul));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

static int tb_tmap(CUtensorMap* tm, const void* base, int ld, int heads, int N, int n_img, int rows) {
  uint64_t dims[4] = {(uint64_t)TB_D, (uint64_t)heads, (uint64_t)N, (uint64_t)n_img};
  uint64_t strides[3] = {(uint64_t)TB_D * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * N};
  uint32_t box[4] = {TB_D, 1, (uint32_t)rows, 1};
  return make_tmap(tm, base, 4, dims, strides, box);
}

// d == 64 only.  q / k / v / dO are head slices inside token rows (pitches in elements), lse / dvec are [n_img, heads, N].
int attn_bwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* dO, int ldo,
                const float* lse, const float* dvec, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int n_img,
                int heads, int N, float scale, cudaStream_t st) {
  static DeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  dim3 grid((N + TB_ROWS - 1) / TB_ROWS, heads, n_img);
  int rc;
  AttnBwdTcParams p;
  p.lse = lse; p.dvec = dvec; p.heads = heads; p.N = N; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  // dQ: resident Q, dO; streamed K, V
  if ((rc = tb_tmap(&p.tmR0, q, ldq, heads, N, n_img, TB_ROWS))) return rc;
  if ((rc = tb_tmap(&p.tmR1, dO, ldo, heads, N, n_img, TB_ROWS))) return rc;
  if ((rc = tb_tmap(&p.tmS0, k, ldk, heads, N, n_img, TB_STEP))) return rc;
  if ((rc = tb_tmap(&p.tmS1, v, ldv, heads, N, n_img, TB_STEP))) return rc;
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dq); p.ld0 = lddq; p.out1 = nullptr; p.ld1 = 0;
  attn_bwd_tc_kernel<false><<<grid, TB_THREADS, TB_SMEM, st>>>(p);
  if ((rc = launch_epilogue())) return rc;
  // dK, dV: resident K, V; streamed Q, dO
  if ((rc = tb_tmap(&p.tmR0, k, ldk, heads, N, n_img, TB_ROWS))) return rc;
  if ((rc = tb_tmap(&p.tmR1, v, ldv, heads, N, n_img, TB_ROWS))) return rc;
  if ((rc = tb_tmap(&p.tmS0, q, ldq, heads, N, n_img, TB_STEP))) return rc;
  if ((rc = tb_tmap(&p.tmS1, dO, ldo, heads, N, n_img, TB_STEP))) return rc;
  p.out0 = reinterpret_cast<__nv_bfloat16*>(dk); p.ld0 = lddk;
  p.out1 = reinterpret_cast<__nv_bfloat16*>(dv); p.ld1 = lddv;
  attn_bwd_tc_kernel<true><<<grid, TB_THREADS, TB_SMEM, st>>>(p);
  return launch_epilogue();
}

}  // namespace lkgd
