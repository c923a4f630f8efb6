// Attention kernels.
//  * attn_flash_kernel: spatial self-attention (and general cross-attention), flash-style on tcgen05.
//      warps 0-3: online softmax (one query row per thread), O accumulation in registers, epilogue
//      warp  4  : TMEM alloc + TMA producer (Q once, K/V double-buffered)
//      warp  5  : tcgen05.mma issuer:  S = Q K^T (TMEM cols 0..127),  PV = P V (TMEM cols 128..191)
//    P (bf16) is written by the softmax warps into SWIZZLE_128B K-major smem and consumed as the A operand;
//    V is consumed as an MN-major B operand straight from its TMA tile (no transpose anywhere).
//    ~112 KB smem and 256 TMEM columns per CTA -> two CTAs per SM overlap softmax with MMA.
//  * attn_temporal_kernel: attention over the frame axis (F <= 32), one warp per (batch, pixel, head),
//    reading the fused qkv projection in place (frame stride HW*3C) - HBM-bound, SIMT.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

constexpr int AT_BQ = 128, AT_D = 64;     // AT_D: width of one SWIZZLE_128B sub-tile (64 bf16 = 128 bytes)
constexpr int AT_SUBQ = AT_BQ * AT_D * 2;  // 16 KB: 128 query rows x 64 head channels
constexpr int AT_THREADS = 192;

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  const __nv_bfloat16* q;   // QT variant: the softmax threads load their own Q row (row pitch ldq elements)
  int ldq;
  __nv_bfloat16* out;
  int ldo, heads, d, Nq, Nk;
  float scale_log2;
  float* lse;       // optional [n_img, heads, Nq]: log2-domain log-sum-exp of the scaled scores (training backward)
};

// three-input maximum (FMNMX3): one ALU-pipe instruction per key pair of the row-maximum scan instead of two
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int AT_BH = 64;                 // keys per softmax step == keys per K/V stage
constexpr int AT_SUBKV = AT_BH * AT_D * 2; // 8 KB: 64 keys x 64 head channels
// Head width D = 64 (also serves d = 16 / 32 through zero-filled TMA boxes) or 128 (the reference's default heads
// (5,10,10,20) give 1280 / 10 = 128 at level 2, models/unet_spatio_temporal_condition_controlnet.py:93).  A D = 128 tile
// is two 64-channel SWIZZLE_128B sub-tiles side by side: Q K^T runs 8 k-steps over them, P V two N = 64 halves.
// QT ("Q in tensor memory", D = 64 only): Q never touches shared memory - every softmax thread loads its own query row
// from global memory and stores it into 32 TMEM columns, and Q K^T is a TS-form MMA (A from TMEM) like P V.  Per 64-key
// step the tensor core then reads 16 KB of operands from smem instead of 32 KB (the 16 KB Q tile re-read by every one of
// the four N = 64 MMAs was half of it), at the price of 192 TMEM columns per CTA: two CTAs per SM instead of three.
template <int D, bool QT = false>
struct AttnCfg {
  static constexpr int NSUB = D / AT_D;                       // 1 or 2
  static constexpr int TILE = QT ? 0 : NSUB * AT_SUBQ;        // Q tile bytes in smem
  static constexpr int KV = NSUB * AT_SUBKV;                  // one 64-key K (or V) stage
  static constexpr int NS = QT ? 4 : (D == 64 ? 3 : 2);       // K/V stages (ring)
  static constexpr int SMEM = TILE + 2 * NS * KV + 128;
  static constexpr int TMEM_MAIN = (D == 64 && !QT) ? 128 : 256;   // S (64 columns) | O (D columns) [| P | Q]
  static constexpr int CTAS = (D == 64 && !QT) ? 3 : 2;       // resident CTAs per SM
};

// Pipeline (per CTA; THREE CTAs share an SM - 64 KB smem, 160 TMEM columns and <= 96 registers each - so that three
// softmax warps per scheduler hide each other's barrier / TMEM / fence latencies and keep the MUFU pipe, which bounds
// d = 64 attention, busy):
//   TMA warp     : Q once; K/V in 64-key stages (ring of three)
//   MMA warp     : S_h = Q K_h^T into ONE 64-column TMEM buffer (S_{h+1} is issued as soon as the softmax warps have
//                  copied S_h into registers), O += P_h V_h with P read from TMEM (no smem round trip for P)
//   softmax warps: ONE tcgen05.ld of the 64 scores of a row into registers -> s_free -> row max -> lazy rescale ->
//                  p = 2^(s c - m) -> packed bf16 pairs -> tcgen05.st into the P columns -> p_full.
// Shared-memory bandwidth was the co-bottleneck of the earlier versions (P written with st.shared and read back by the
// tensor core: 32 KB of the 64 KB smem traffic per step).
// ---- packed fp32x2 helpers (FFMA2 / FADD2: one issue slot for two elements)
__device__ __forceinline__ uint64_t at_pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void at_upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t at_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t at_add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for a pair on the FMA / ALU pipes (no MUFU): x = n + r with n = round(x) taken from the mantissa of x + 1.5 * 2^23,
// 2^r on [-0.5, 0.5] as a degree-3 minimax polynomial (max relative error 7.5e-5, 15x below the bf16 rounding of P),
// 2^n added into the exponent field.  x is clamped at -126 (also maps the -inf of masked keys to ~1e-38).
__device__ __forceinline__ uint64_t at_exp2_poly2(uint64_t X) {
  float x0, x1;
  at_upk2(X, x0, x1);
  X = at_pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t MAGIC = at_pk2(12582912.0f, 12582912.0f), NMAGIC = at_pk2(-12582912.0f, -12582912.0f);
  const uint64_t T = at_add2(X, MAGIC);
  const uint64_t N = at_add2(T, NMAGIC);
  const uint64_t R = at_fma2(N, at_pk2(-1.0f, -1.0f), X);
  uint64_t P = at_fma2(at_pk2(0.0551716685f, 0.0551716685f), R, at_pk2(0.2426111251f, 0.2426111251f));
  P = at_fma2(P, R, at_pk2(0.6932609677f, 0.6932609677f));
  P = at_fma2(P, R, at_pk2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  at_upk2(T, t0, t1);
  at_upk2(P, p0, p1);
  return at_pk2(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)),
                __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}

// POLY: every POLY-th pair of exponentials is evaluated on the FMA pipes instead of the MUFU pipe (0 = none).  d = 64
// attention is bound by the 16 ex2 / clock / SM of the MUFU pipe; the FMA pipes are otherwise nearly idle here.
template <int POLY, int D, bool QT = false>
__global__ void __launch_bounds__(AT_THREADS, AttnCfg<D, QT>::CTAS) attn_flash_kernel(const __grid_constant__ AttnParams p) {
  static_assert(!QT || D == 64, "Q in TMEM is built for 64-wide heads");
  using Cfg = AttnCfg<D, QT>;
  constexpr int AT_TILE = Cfg::TILE, AT_KV = Cfg::KV, AT_NS = Cfg::NS, NSUB = Cfg::NSUB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = sK + AT_NS * AT_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AT_NS * AT_KV);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;            // [AT_NS]
  uint64_t* kv_empty = bars + 1 + AT_NS;   // [AT_NS]
  uint64_t* s_full = bars + 1 + 2 * AT_NS; //      S holds Q K_h^T
  uint64_t* s_free = s_full + 1;           //      every softmax warp has copied S into registers
  uint64_t* p_full = s_full + 2;           //      P_h is in TMEM
  uint64_t* pv_done = s_full + 3;          //      P_h V_h has landed in O (and the P columns may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);     // [2]: 128 columns (S | O) + 32 columns (P)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, img = blockIdx.z;
  const int H = (p.Nk + AT_BH - 1) / AT_BH;     // 64-key steps

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, QT ? 128 : 1);
      for (int i = 0; i < AT_NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      mbar_init(s_full, 1);
      mbar_init(s_free, 4);
      mbar_init(p_full, 128);
      mbar_init(pv_done, 1);
      fence_barrier_init();
      tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV);
    }
    __syncwarp();
    if (D == 64 && !QT) {
      tmem_alloc_keep_permit(tmem_slot, Cfg::TMEM_MAIN);
      tmem_alloc(tmem_slot + 1, 32);
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_MAIN);     // D = 128: S | O | P fit one 256-column allocation (two CTAs per SM)
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + AT_BH;   // S at columns 0..63, O at 64..127
  const uint32_t tmem_P = (D == 64 && !QT) ? tmem_slot[1] : tmem_base + AT_BH + D;   // 128 x 64 bf16 = 32 columns
  const uint32_t tmem_Q = tmem_base + AT_BH + D + 32;                                // QT: 128 x 64 bf16 = 32 columns

  if (warp == 4) {
    if (elect_one()) {
      if (!QT) {
        mbar_expect_tx(q_full, AT_TILE);
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tma_load_4d(sQ + sb * AT_SUBQ, &p.tmQ, q_full, sb * AT_D, head, q0, img);
      }
      int st = 0, ph = 1;
      for (int j = 0; j < H; ++j) {
        while (!mbar_try_wait(&kv_empty[st], ph)) __nanosleep(64);   // off the critical path: back off
        mbar_expect_tx(&kv_full[st], 2 * AT_KV);
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {
          tma_load_4d(sK + st * AT_KV + sb * AT_SUBKV, &p.tmK, &kv_full[st], sb * AT_D, head, j * AT_BH, img);
          tma_load_4d(sV + st * AT_KV + sb * AT_SUBKV, &p.tmV, &kv_full[st], sb * AT_D, head, j * AT_BH, img);
        }
        if (++st == AT_NS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BH);             // N = 64 keys
      const uint32_t idesc_o = umma_idesc_bf16(AT_D, 128, 0, 1);   // N = 64 (per 64-channel half), B (= V) is MN-major
      const uint32_t sQ0 = smem_u32(sQ), sK0 = smem_u32(sK), sV0 = smem_u32(sV);
      int qst = 0, qph = 0;            // K/V stage (and its phase) of the next Q K^T
      auto issue_qk = [&]() {
        mbar_wait(&kv_full[qst], qph);
        tc_fence_after();
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {        // contraction over the head channels: 64 per sub-tile
          const uint64_t qdesc = umma_desc_sw128(sQ0 + sb * AT_SUBQ);
          const uint64_t kdesc = umma_desc_sw128(sK0 + qst * AT_KV + sb * AT_SUBKV);
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) {
            if (QT) umma_bf16_ts(tmem_S, tmem_Q + k * 8, kdesc + 2 * k, idesc_s, k != 0);   // A = Q from TMEM
            else umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, (sb | k) != 0);
          }
        }
        umma_commit(s_full);
        if (++qst == AT_NS) { qst = 0; qph ^= 1; }
      };
      mbar_wait(q_full, 0);
      if (QT) tc_fence_after();
      issue_qk();
      int st = 0;                      // K/V stage of P_h V_h
      for (int h = 0; h < H; ++h) {
        if (h + 1 < H) {               // S is in the softmax warps' registers: start the next Q K^T now
          mbar_wait(s_free, h & 1);
          tc_fence_after();
          issue_qk();
        }
        mbar_wait(p_full, h & 1);      // P_h is in smem
        tc_fence_after();
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {        // output channels: one N = 64 half per V sub-tile
#pragma unroll
          for (int ks = 0; ks < AT_BH / 16; ++ks) {
            const uint64_t vdesc = umma_desc_sw128(sV0 + st * AT_KV + sb * AT_SUBKV + ks * 2048);
            umma_bf16_ts(tmem_O + sb * AT_D, tmem_P + ks * 8, vdesc, idesc_o, (h | ks) != 0);   // O accumulates in TMEM
          }
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
        if (++st == AT_NS) st = 0;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // One query row per thread.  O stays in TMEM and accumulates across KV blocks; the running maximum is only
    // raised when a row's block maximum exceeds it by more than 2^8 ("lazy rescale"): then the warp multiplies its
    // O rows in TMEM by 2^(m_old - m_new).  p = 2^(s*c - m) <= 256 otherwise, exact enough in bf16 / fp32.
    const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    if (QT) {
      // this thread's query row: 64 bf16 (zero beyond d / beyond Nq) -> 32 TMEM columns, word j = channels 2j, 2j+1
      uint32_t qw[32];
      const bool row_ok = q0 + r < p.Nq;
      const __nv_bfloat16* qrow = p.q + ((size_t)img * p.Nq + (row_ok ? q0 + r : 0)) * p.ldq + head * p.d;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && c * 8 < p.d) u = __ldg(reinterpret_cast<const uint4*>(qrow) + c);
        qw[4 * c] = u.x; qw[4 * c + 1] = u.y; qw[4 * c + 2] = u.z; qw[4 * c + 3] = u.w;
      }
      tmem_st32(tmem_Q + lane_addr, qw);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_full);
    }
    float m_run = -INFINITY, l_run = 0.f;
    const float sc = p.scale_log2;
    for (int h = 0; h < H; ++h) {
      mbar_wait(s_full, h & 1);                 // S_h = Q K_h^T is in TMEM
      tc_fence_after();
      uint32_t s[AT_BH];
      tmem_ld32(tmem_S + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_S + lane_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(s_free);       // the tensor core may overwrite S (with S_{h+1})
      const int kv_valid = p.Nk - h * AT_BH;
      if (kv_valid < AT_BH) {                   // last, partial step: -inf scores give p = 0
#pragma unroll
        for (int i = 0; i < AT_BH; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < AT_BH; i += 2) {
        if (POLY == 4)      // A/B: the two-instruction form (POLY = 4 is the comparison variant of tools/bench_attn.py)
          m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
        else
          m4[(i >> 1) & 3] = fmax3(m4[(i >> 1) & 3], __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
      }
      const float m_blk = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
      // lazy rescale: decided now, applied to O (below) only after P_{h-1} V_{h-1} has landed
      float alpha = 1.0f;
      bool rescale = false;
      if (h == 0) {
        m_run = m_blk;
      } else {
        const bool grow = m_blk > m_run + 8.0f;
        rescale = __any_sync(0xffffffffu, grow);
        if (grow) {
          alpha = ex2f(m_run - m_blk);
          l_run *= alpha;
          m_run = m_blk;
        }
      }
      // p = 2^(s*c - m) in f32, packed to bf16 pairs in place (word j = keys 2j, 2j+1)
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
      if (POLY < 0) {        // scalar reference formulation (A/B only)
#pragma unroll
        for (int i = 0; i < AT_BH; i += 2) {
          const float p0 = ex2f(fmaf(__uint_as_float(s[i]), sc, -m_run));
          const float p1 = ex2f(fmaf(__uint_as_float(s[i + 1]), sc, -m_run));
          ls4[(i >> 1) & 3] += p0 + p1;
          s[i >> 1] = pack_bf16x2(p0, p1);
        }
      } else {
        const uint64_t SC2 = at_pk2(sc, sc), NM2 = at_pk2(-m_run, -m_run);
        uint64_t LS[2] = {0ull, 0ull};
#pragma unroll
        for (int j = 0; j < AT_BH / 2; ++j) {
          const uint64_t X = at_fma2(at_pk2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), SC2, NM2);
          uint64_t P;
          constexpr int PEFF = POLY == 4 ? 3 : POLY;      // POLY = 4: the r01 kernel (every third pair, two-instruction max)
          if (PEFF > 0 && (j % (PEFF > 0 ? PEFF : 1)) == PEFF - 1) {
            P = at_exp2_poly2(X);
          } else {
            float x0, x1;
            at_upk2(X, x0, x1);
            P = at_pk2(ex2f(x0), ex2f(x1));
          }
          LS[j & 1] = at_add2(LS[j & 1], P);
          float p0, p1;
          at_upk2(P, p0, p1);
          s[j] = pack_bf16x2(p0, p1);
        }
        at_upk2(LS[0], ls4[0], ls4[1]);
        at_upk2(LS[1], ls4[2], ls4[3]);
      }
      if (h > 0) {
        mbar_wait(pv_done, (h - 1) & 1);        // P_{h-1} V_{h-1} has landed: O may be rescaled, P may be overwritten
        if (rescale) {
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < D; c += 16) {
            uint32_t t[16];
            tmem_ld16(tmem_O + lane_addr + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st16(tmem_O + lane_addr + c, t);
          }
        }
      }
      tmem_st32(tmem_P + lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      l_run += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      tmem_st_wait();
      tc_fence_before();          // P (and an O rescale, if any) is in TMEM before the MMA warp is signalled
      mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (H - 1) & 1);
    tc_fence_after();
    if (q0 + r < p.Nq) {
      if (p.lse != nullptr) p.lse[((size_t)img * p.heads + head) * p.Nq + q0 + r] = m_run + log2f(l_run);
    }
    const float inv = 1.0f / l_run;
    __nv_bfloat16* orow = p.out + ((size_t)img * p.Nq + q0 + r) * p.ldo + head * p.d;
#pragma unroll
    for (int c = 0; c < D; c += 32) {
      uint32_t t[32];
      tmem_ld32(tmem_O + lane_addr + c, t);
      tmem_ld_wait();
      if (q0 + r < p.Nq) {
        if (p.d == D) {
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(orow + c + i) =
                make_uint4(pack_bf16x2(__uint_as_float(t[i]) * inv, __uint_as_float(t[i + 1]) * inv),
                           pack_bf16x2(__uint_as_float(t[i + 2]) * inv, __uint_as_float(t[i + 3]) * inv),
                           pack_bf16x2(__uint_as_float(t[i + 4]) * inv, __uint_as_float(t[i + 5]) * inv),
                           pack_bf16x2(__uint_as_float(t[i + 6]) * inv, __uint_as_float(t[i + 7]) * inv));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < p.d) orow[c + i] = __float2bfloat16(__uint_as_float(t[i]) * inv);
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
    tmem_dealloc(tmem_base, Cfg::TMEM_MAIN);
    if (D == 64 && !QT) tmem_dealloc(tmem_P, 32);
  }
}

// ------------------------------------------------------------------------------------------- temporal
// One warp per (batch, pixel, head): the F <= 32 frames of that pixel are a 32 x D problem - far too small for a
// tcgen05 tile, HBM-bound as a whole (reads q|k|v once, writes o once).  cp.async pulls the three F x D tiles out of
// the fused projection in place (frame stride HW*3C, 16-byte pieces coalesced along d), ldmatrix + mma.sync
// m16n8k16 do S = Q K^T and O = P V with the softmax between them in the accumulator fragments, and the result
// leaves through smem as full 16-byte row pieces.
template <int D>
struct TAttn {
  static constexpr int P = D / 8;            // 16-byte chunks per row
  static constexpr int RPL = 8 / P > 0 ? 8 / P : 1;
  static constexpr int ROW = D * 2;          // row pitch in bytes
  static constexpr int TILE = 32 * ROW;      // one 32 x D tile
  __device__ static __forceinline__ uint32_t off(int row, int chunk) {   // XOR swizzle: ldmatrix conflict-free
    return static_cast<uint32_t>(row * ROW + ((chunk ^ ((row / RPL) & (P - 1))) << 4));
  }
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <int D>
__global__ void __launch_bounds__(128) attn_temporal_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int B, int F, int HW,
                                                            int heads, float scale_log2) {
  using T = TAttn<D>;
  extern __shared__ __align__(128) uint8_t tsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)B * HW * heads;
  if (seq >= total) return;
  const int head = (int)(seq % heads);
  const int pix = (int)((seq / heads) % HW);
  const int b = (int)(seq / ((long long)heads * HW));
  const int C = heads * D;
  const size_t ld = (size_t)3 * C;
  const uint32_t sQ = smem_u32(tsm + warp * 3 * T::TILE), sK = sQ + T::TILE, sV = sK + T::TILE;
  const __nv_bfloat16* base = qkv + ((size_t)b * F * HW + pix) * ld + head * D;
  // ---- global -> smem: 32 rows (frames, zero-filled beyond F) x P chunks per tile
#pragma unroll
  for (int j = 0; j < T::P; ++j) {
    const int i = lane + 32 * j;
    const int row = i / T::P, ch = i % T::P;
    const bool ok = row < F;
    const __nv_bfloat16* src = base + (ok ? (size_t)row * HW * ld + ch * 8 : 0);
    const uint32_t o = T::off(row, ch);
    cp_async_16(sQ + o, src, ok ? 16u : 0u);
    cp_async_16(sK + o, src + C, ok ? 16u : 0u);
    cp_async_16(sV + o, src + 2 * C, ok ? 16u : 0u);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  // ---- S = Q K^T : 2 m-tiles (query frames) x 4 n-tiles (key frames)
  float sacc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    uint32_t a[2][4], kb[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) ldsm_x4(sQ + T::off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), a[mt]);
#pragma unroll
    for (int np = 0; np < 2; ++np)     // two key n-tiles per ldmatrix.x4: {nt0 k-lo, nt0 k-hi, nt1 k-lo, nt1 k-hi}
      ldsm_x4(sK + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), kb[np]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(sacc[mt][nt], a[mt], kb[nt >> 1][(nt & 1) * 2], kb[nt >> 1][(nt & 1) * 2 + 1]);
  }
  // ---- softmax over the key frames (columns nt*8 + 2*(lane%4) + {0,1}); rows lane/4 and lane/4 + 8 of each m-tile
  uint32_t pa[2][2][4];     // P as bf16 A fragments: [m-tile][k-step of 16 keys]
  float inv_l[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = nt * 8 + 2 * (lane & 3) + e;
          if (key >= F) sacc[mt][nt][hr * 2 + e] = -INFINITY;
          mx = fmaxf(mx, sacc[mt][nt][hr * 2 + e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float moff = mx * scale_log2;
      float l = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pv = ex2f(fmaf(sacc[mt][nt][hr * 2 + e], scale_log2, -moff));
          sacc[mt][nt][hr * 2 + e] = pv;
          l += pv;
        }
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv_l[mt][hr] = 1.0f / l;
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      pa[mt][kk][0] = pack_bf16x2(sacc[mt][2 * kk][0], sacc[mt][2 * kk][1]);
      pa[mt][kk][1] = pack_bf16x2(sacc[mt][2 * kk][2], sacc[mt][2 * kk][3]);
      pa[mt][kk][2] = pack_bf16x2(sacc[mt][2 * kk + 1][0], sacc[mt][2 * kk + 1][1]);
      pa[mt][kk][3] = pack_bf16x2(sacc[mt][2 * kk + 1][2], sacc[mt][2 * kk + 1][3]);
    }
  }
  // ---- O = P V : V tile is [key][d] = K x N row-major -> transposed ldmatrix
  float oacc[2][D / 8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[mt][nd][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int np = 0; np < D / 16; ++np) {   // {keys lo d0, keys hi d0, keys lo d1, keys hi d1}
      uint32_t vb[4];
      ldsm_x4_t(sV + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), vb);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_bf16_16816(oacc[mt][np * 2], pa[mt][kk], vb[0], vb[1]);
        mma_bf16_16816(oacc[mt][np * 2 + 1], pa[mt][kk], vb[2], vb[3]);
      }
    }
  }
  // ---- normalise, stage through the (now free) Q tile, write full 16-byte pieces of each frame's row
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int row = mt * 16 + hr * 8 + (lane >> 2);
        const uint32_t v = pack_bf16x2(oacc[mt][nd][hr * 2] * inv_l[mt][hr], oacc[mt][nd][hr * 2 + 1] * inv_l[mt][hr]);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + T::off(row, nd) + (lane & 3) * 4), "r"(v) : "memory");
      }
  __syncwarp();
  __nv_bfloat16* obase = out + ((size_t)b * F * HW + pix) * C + head * D;
#pragma unroll
  for (int j = 0; j < T::P; ++j) {
    const int i = lane + 32 * j;
    const int row = i / T::P, ch = i % T::P;
    if (row < F) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + T::off(row, ch)));
      *reinterpret_cast<uint4*>(obase + (size_t)row * HW * C + ch * 8) = v;
    }
  }
}

static int attn_tmap(CUtensorMap* tm, const void* base, int ld, int heads, int d, int N, int n_img, int rows) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)n_img};
  uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * N};
  uint32_t box[4] = {AT_D, 1, (uint32_t)rows, 1};
  return make_tmap(tm, base, 4, dims, strides, box);
}

}  // namespace lkgd

using namespace lkgd;

static int attention_impl(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                          void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq,
                          int32_t Nk, float scale, float* lse, void* stream) {
  if (n_img <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0 || d % 8 || d <= 0 || d > 128) return LKGD_ESHAPE;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return LKGD_EALIGN;
  if (heads > 65535 || n_img > 65535) return LKGD_ESHAPE;
  AttnParams p;
  int rc;
  if ((rc = attn_tmap(&p.tmQ, q, ldq, heads, d, Nq, n_img, AT_BQ))) return rc;
  if ((rc = attn_tmap(&p.tmK, k, ldk, heads, d, Nk, n_img, AT_BH))) return rc;
  if ((rc = attn_tmap(&p.tmV, v, ldv, heads, d, Nk, n_img, AT_BH))) return rc;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.ldo = ldo; p.heads = heads; p.d = d; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  constexpr int AT_SMEM = AttnCfg<64>::SMEM, AT_SMEM128 = AttnCfg<128>::SMEM;
  static DeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<-1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM128);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<4, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64, true>::SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  dim3 grid((Nq + AT_BQ - 1) / AT_BQ, heads, n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d > AT_D) {      // 72 .. 128: the two-sub-tile kernel; TMA zero-fills the channels beyond d (CLIP ViT-H: d = 80)
    attn_flash_kernel<4, 128><<<grid, AT_THREADS, AT_SMEM128, st>>>(p);
    return launch_epilogue();
  }
  // tuning switch (tools/bench_attn.py, LKGD_ATTN_POLY_AB): -1 scalar formulation, 0 packed without offload, 3 every
  // third pair on the FMA pipes + FMNMX3 row maximum, 4 = 3 with the two-instruction maximum (the r01 kernel).
  const char* pe = getenv("LKGD_ATTN_POLY");
  const int poly = pe ? atoi(pe) : 4;      // FMNMX3 (3) measured 3 % SLOWER than two FMNMX at L0 (6.18 vs 6.01 ms): half rate
  // LKGD_ATTN_QT=1: Q as a TMEM operand (two CTAs per SM) - A/B switch of tools/bench_attn.py; needs 16-byte aligned rows
  if (const char* qt = getenv("LKGD_ATTN_QT")) {
    if (atoi(qt) == 1 && aligned16(q) && (heads * d) % 8 == 0 && d % 8 == 0) {
      attn_flash_kernel<4, 64, true><<<grid, AT_THREADS, AttnCfg<64, true>::SMEM, st>>>(p);
      return launch_epilogue();
    }
  }
  switch (poly) {
    case -1: attn_flash_kernel<-1, 64><<<grid, AT_THREADS, AT_SMEM, st>>>(p); break;
    case 0: attn_flash_kernel<0, 64><<<grid, AT_THREADS, AT_SMEM, st>>>(p); break;
    case 3: attn_flash_kernel<3, 64><<<grid, AT_THREADS, AT_SMEM, st>>>(p); break;
    default: attn_flash_kernel<4, 64><<<grid, AT_THREADS, AT_SMEM, st>>>(p); break;
  }
  return launch_epilogue();
}

extern "C" int lkgd_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                              void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq,
                              int32_t Nk, float scale, void* stream) {
  return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, n_img, heads, d, Nq, Nk, scale, nullptr, stream);
}

extern "C" int lkgd_attention_lse(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                  void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq,
                                  int32_t Nk, float scale, float* lse, void* stream) {
  if (lse == nullptr) return LKGD_ESHAPE;
  return attention_impl(q, ldq, k, ldk, v, ldv, out, ldo, n_img, heads, d, Nq, Nk, scale, lse, stream);
}

extern "C" int lkgd_attention_temporal(const void* qkv, void* out, int32_t B, int32_t F, int32_t HW, int32_t heads,
                                       int32_t d, float scale, void* stream) {
  if (B <= 0 || F <= 0 || F > 32 || HW <= 0 || heads <= 0) return LKGD_ESHAPE;
  if (!aligned16(qkv) || !aligned16(out)) return LKGD_EALIGN;
  const long long total = (long long)B * HW * heads;
  const unsigned grid = (unsigned)((total + 3) / 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const float sl2 = scale * 1.4426950408889634f;
  static DeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * TAttn<64>::TILE);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_temporal_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * TAttn<128>::TILE);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  switch (d) {
    case 16: attn_temporal_kernel<16><<<grid, 128, 4 * 3 * TAttn<16>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    case 32: attn_temporal_kernel<32><<<grid, 128, 4 * 3 * TAttn<32>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    case 64: attn_temporal_kernel<64><<<grid, 128, 4 * 3 * TAttn<64>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    case 128: attn_temporal_kernel<128><<<grid, 128, 4 * 3 * TAttn<128>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    default: return LKGD_ESHAPE;
  }
  return launch_epilogue();
}
