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

constexpr int AT_BQ = 128, AT_BK = 128, AT_D = 64;
constexpr int AT_TILE = AT_BQ * AT_D * 2;  // 16 KB
constexpr int AT_THREADS = 192;
constexpr int AT_SMEM = AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + 2 * AT_TILE /*V*/ + 2 * AT_TILE /*P*/ + 128;

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  int ldo, heads, d, Nq, Nk;
  float scale_log2;
  float* lse;       // optional [n_img, heads, Nq]: log2-domain log-sum-exp of the scaled scores (training backward)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int AT_BH = 64;   // keys per softmax step: half of a 128-key K/V stage

// Pipeline (per CTA; two CTAs share an SM):
//   TMA warp     : Q once; K/V in 128-key stages (double-buffered)
//   MMA warp     : S_h = Q K_h^T for 64-key half-blocks h into TWO 64-column TMEM buffers (S_{h+2} is issued as soon as
//                  the softmax warps have copied S_h into registers), O += P_h V_h
//   softmax warps: ONE tcgen05.ld of the 64 scores of a row into registers -> s_free -> row max -> lazy rescale ->
//                  p = 2^(s c - m) -> bf16 P tile in smem -> p_full.  The scores are read from TMEM once, and the next
//                  two QK^T products are already done or in flight while a half-block's exponentials are computed, so the
//                  softmax warps never wait for the tensor core (the previous kernel waited ~20 % of its time for S).
__global__ void __launch_bounds__(AT_THREADS, 2) attn_flash_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 3 * AT_TILE;
  uint8_t* sP = smem + 5 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * AT_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]  S buffer b holds Q K_h^T (h & 1 == b)
  uint64_t* s_free = bars + 7;    // [2]  every softmax warp has copied S buffer b into registers
  uint64_t* p_full = bars + 9;    // [2]  P_h is in smem buffer b (one barrier per buffer: a consumer never lags two phases)
  uint64_t* pv_done = bars + 11;  // [2]  P_h V_h has landed in O (and P buffer b may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, img = blockIdx.z;
  const int T = (p.Nk + AT_BK - 1) / AT_BK;     // 128-key K/V stages
  const int H = (p.Nk + AT_BH - 1) / AT_BH;     // 64-key softmax steps

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
        mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&pv_done[i], 1);
        mbar_init(&p_full[i], 128);
      }
      fence_barrier_init();
      tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;   // S buffers at columns 0 / 64, O at 128..191

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, head, q0, img);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        while (!mbar_try_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1)) __nanosleep(64);   // off the critical path: back off
        mbar_expect_tx(&kv_full[st], 2 * AT_TILE);
        tma_load_4d(sK + st * AT_TILE, &p.tmK, &kv_full[st], 0, head, j * AT_BK, img);
        tma_load_4d(sV + st * AT_TILE, &p.tmV, &kv_full[st], 0, head, j * AT_BK, img);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BH);             // N = 64 keys
      const uint32_t idesc_o = umma_idesc_bf16(AT_D, 128, 0, 1);   // N = 64, B (= V) is MN-major
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      const uint32_t sK0 = smem_u32(sK), sV0 = smem_u32(sV), sP0 = smem_u32(sP);
      auto issue_qk = [&](int h) {
        const int st = (h >> 1) & 1;
        if ((h & 1) == 0) {            // first half of K/V stage h / 2
          mbar_wait(&kv_full[st], (h >> 2) & 1);
          tc_fence_after();
        }
        const uint64_t kdesc = umma_desc_sw128(sK0 + st * AT_TILE + (h & 1) * (AT_BH * 128));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_bf16(tmem_S + (h & 1) * AT_BH, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[h & 1]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (H > 1) issue_qk(1);
      for (int h = 0; h < H; ++h) {
        const int b = h & 1, st = (h >> 1) & 1;
        if (h + 2 < H) {               // S buffer b is in the softmax warps' registers: refill it two steps ahead
          mbar_wait(&s_free[b], (h >> 1) & 1);
          tc_fence_after();
          issue_qk(h + 2);
        }
        mbar_wait(&p_full[b], (h >> 1) & 1);      // P_h is in smem
        tc_fence_after();
        const uint64_t pdesc = umma_desc_sw128(sP0 + b * AT_TILE);
#pragma unroll
        for (int ks = 0; ks < AT_BH / 16; ++ks) {
          const uint64_t vdesc = umma_desc_sw128(sV0 + st * AT_TILE + (b * 4 + ks) * 2048);
          umma_bf16(tmem_O, pdesc + 2 * ks, vdesc, idesc_o, (h | ks) != 0);     // O accumulates in TMEM
        }
        umma_commit(&pv_done[b]);
        if (b == 1 || h == H - 1) umma_commit(&kv_empty[st]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // One query row per thread.  O stays in TMEM and accumulates across KV blocks; the running maximum is only
    // raised when a row's block maximum exceeds it by more than 2^8 ("lazy rescale"): then the warp multiplies its
    // O rows in TMEM by 2^(m_old - m_new).  p = 2^(s*c - m) <= 256 otherwise, exact enough in bf16 / fp32.
    const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t prow = smem_u32(sP) + (r >> 3) * 1024 + (r & 7) * 128;
    const float sc = p.scale_log2;
    for (int h = 0; h < H; ++h) {
      const int b = h & 1;
      mbar_wait(&s_full[b], (h >> 1) & 1);      // S_h = Q K_h^T is in TMEM
      tc_fence_after();
      uint32_t s[AT_BH];
      tmem_ld32(tmem_S + lane_addr + b * AT_BH, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_S + lane_addr + b * AT_BH + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(&s_free[b]);   // the tensor core may overwrite this S buffer (with S_{h+2})
      const int kv_valid = p.Nk - h * AT_BH;
      if (kv_valid < AT_BH) {                   // last, partial half-block: -inf scores give p = 0
#pragma unroll
        for (int i = 0; i < AT_BH; ++i)
          if (i >= kv_valid) s[i] = 0xff800000u;
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < AT_BH; i += 2)
        m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
      const float m_blk = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
      if (h == 0) {
        m_run = m_blk;
      } else {
        const bool grow = m_blk > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&pv_done[(h - 1) & 1], ((h - 1) >> 1) & 1);   // P_{h-1} V_{h-1} must have landed before O is rescaled
          tc_fence_after();
          const float m_new = grow ? m_blk : m_run;
          const float alpha = ex2f(m_run - m_new);     // 1 for the rows that keep their maximum
#pragma unroll
          for (int c = 0; c < AT_D; c += 32) {
            uint32_t t[32];
            tmem_ld32(tmem_O + lane_addr + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(tmem_O + lane_addr + c, t);
          }
          tmem_st_wait();
          l_run *= alpha;
          m_run = m_new;
        }
      }
      // p = 2^(s*c - m) in f32, packed to bf16 into the swizzled K-major P tile b (free once P_{h-2} V_{h-2} has been
      // consumed by the tensor core)
      if (h >= 2) mbar_wait(&pv_done[b], ((h - 2) >> 1) & 1);
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t blk = prow + b * AT_TILE;
#pragma unroll
      for (int c = 0; c < AT_BH; c += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float p0 = ex2f(fmaf(__uint_as_float(s[c + i]), sc, -m_run));
          const float p1 = ex2f(fmaf(__uint_as_float(s[c + i + 1]), sc, -m_run));
          ls4[(i >> 1) & 3] += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (((c >> 3) ^ (r & 7)) << 4)), "r"(pk[0]),
                     "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      }
      l_run += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      tc_fence_before();          // O rescale (if any) is complete before the MMA warp may touch O
      fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&p_full[b]);
    }
    mbar_wait(&pv_done[(H - 1) & 1], ((H - 1) >> 1) & 1);
    tc_fence_after();
    float o[AT_D];
#pragma unroll
    for (int c = 0; c < AT_D; c += 32) {
      uint32_t t[32];
      tmem_ld32(tmem_O + lane_addr + c, t);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(t[i]);
    }
    tc_fence_before();
    if (q0 + r < p.Nq) {
      const float inv = 1.0f / l_run;
      if (p.lse != nullptr) p.lse[((size_t)img * p.heads + head) * p.Nq + q0 + r] = m_run + log2f(l_run);
      __nv_bfloat16* orow = p.out + ((size_t)img * p.Nq + q0 + r) * p.ldo + head * p.d;
      if (p.d == AT_D) {
#pragma unroll
        for (int c = 0; c < AT_D; c += 8)
          *reinterpret_cast<uint4*>(orow + c) =
              make_uint4(pack_bf16x2(o[c] * inv, o[c + 1] * inv), pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv),
                         pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv), pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv));
      } else {
#pragma unroll
        for (int c = 0; c < AT_D; ++c)
          if (c < p.d) orow[c] = __float2bfloat16(o[c] * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// Previous pipeline (128-key softmax steps, S read from TMEM twice, single S buffer): kept for A/B timing only
// (LKGD_ATTN_V1=1).
__global__ void __launch_bounds__(AT_THREADS, 2) attn_flash_v1_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 3 * AT_TILE;
  uint8_t* sP = smem + 5 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * AT_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, img = blockIdx.z;
  const int T = (p.Nk + AT_BK - 1) / AT_BK;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
      mbar_init(o_full, 1);
      fence_barrier_init();
      tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, head, q0, img);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        while (!mbar_try_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1)) __nanosleep(64);   // off the critical path: back off
        mbar_expect_tx(&kv_full[st], 2 * AT_TILE);
        tma_load_4d(sK + st * AT_TILE, &p.tmK, &kv_full[st], 0, head, j * AT_BK, img);
        tma_load_4d(sV + st * AT_TILE, &p.tmV, &kv_full[st], 0, head, j * AT_BK, img);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BK);             // N = 128 keys
      const uint32_t idesc_o = umma_idesc_bf16(AT_D, 128, 0, 1);   // N = 64, B (= V) is MN-major
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);     // P_j is in smem, S columns are free again
        tc_fence_after();
        // S_{j+1} = Q K_{j+1}^T first: the softmax warps wait for it, nobody waits for P_j V_j
        if (j + 1 < T) {
          const int sn = (j + 1) & 1;
          mbar_wait(&kv_full[sn], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + sn * AT_TILE));
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
#pragma unroll
        for (int ks = 0; ks < AT_BK / 16; ++ks) {
          const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + (ks >> 2) * AT_TILE)) + 2 * (ks & 3);
          const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + st * AT_TILE + ks * 2048));
          umma_bf16(tmem_O, pdesc, vdesc, idesc_o, (j | ks) != 0);     // O accumulates in TMEM across KV blocks
        }
        umma_commit(&kv_empty[st]);
        umma_commit(o_full);          // phase j: P_j V_j has landed in O (and the P tile may be overwritten)
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    // One query row per thread.  O stays in TMEM and accumulates across KV blocks; the running maximum is only
    // raised when a row's block maximum exceeds it by more than 2^8 ("lazy rescale"): then the warp multiplies its
    // O rows in TMEM by 2^(m_old - m_new).  p = 2^(s*c - m) <= 256 otherwise, exact enough in bf16 / fp32.
    const int r = warp * 32 + lane;  // query row in the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    for (int j = 0; j < T; ++j) {
      mbar_wait(s_full, j & 1);      // S_j = Q K_j^T is in TMEM
      tc_fence_after();
      const int kv_valid = min(AT_BK, p.Nk - j * AT_BK);
      // pass 1: row maximum (64 columns per TMEM wait, four independent chains)
      float mx;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < AT_BK; c += 64) {
          uint32_t s0[32], s1[32];
          tmem_ld32(tmem_S + lane_addr + c, s0);
          tmem_ld32(tmem_S + lane_addr + c + 32, s1);
          tmem_ld_wait();
          if (c + 64 <= kv_valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              m4[(i >> 1) & 1] = fmaxf(m4[(i >> 1) & 1], fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));
              m4[2 + ((i >> 1) & 1)] = fmaxf(m4[2 + ((i >> 1) & 1)], fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (c + i < kv_valid) m4[i & 1] = fmaxf(m4[i & 1], __uint_as_float(s0[i]));
              if (c + 32 + i < kv_valid) m4[2 + (i & 1)] = fmaxf(m4[2 + (i & 1)], __uint_as_float(s1[i]));
            }
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      const float m_blk = mx * p.scale_log2;
      if (j == 0) {
        m_run = m_blk;
      } else {
        const bool grow = m_blk > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(o_full, (j - 1) & 1);              // P_{j-1} V_{j-1} must have landed before O is rescaled
          tc_fence_after();
          const float m_new = grow ? m_blk : m_run;
          const float alpha = ex2f(m_run - m_new);     // 1 for the rows that keep their maximum
#pragma unroll
          for (int c = 0; c < AT_D; c += 32) {
            uint32_t t[32];
            tmem_ld32(tmem_O + lane_addr + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(tmem_O + lane_addr + c, t);
          }
          tmem_st_wait();
          l_run *= alpha;
          m_run = m_new;
        }
      }
      float lsum = 0.f;
      // pass 2: p = 2^(s*c - m) in f32, packed to bf16 into the swizzled K-major P tile (free once P_{j-1} V_{j-1}
      // has been consumed by the tensor core)
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
      if (j > 0) mbar_wait(o_full, (j - 1) & 1);
      const bool full_blk = kv_valid == AT_BK;     // every block but possibly the last: no per-element masking
#pragma unroll
      for (int c = 0; c < AT_BK; c += 32) {
        uint32_t s[32];
        tmem_ld32(tmem_S + lane_addr + c, s);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full_blk) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2f(fmaf(__uint_as_float(s[i]), p.scale_log2, -m_run));
            const float p1 = ex2f(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -m_run));
            ls4[(i >> 1) & 3] += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c + i < kv_valid) ? ex2f(fmaf(__uint_as_float(s[i]), p.scale_log2, -m_run)) : 0.f;
            const float p1 = (c + i + 1 < kv_valid) ? ex2f(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -m_run)) : 0.f;
            ls4[(i >> 1) & 3] += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
        }
        uint8_t* blk = prow + (c >> 6) * AT_TILE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 63) >> 3) + q;   // 16-byte chunk index within the 128-byte row
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      lsum = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      l_run += lsum;
      tc_fence_before();          // S reads (and O rescale) are complete before the MMA warp may touch S / O
      fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor-core (async) proxy
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (T - 1) & 1);
    tc_fence_after();
    float o[AT_D];
#pragma unroll
    for (int c = 0; c < AT_D; c += 32) {
      uint32_t t[32];
      tmem_ld32(tmem_O + lane_addr + c, t);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(t[i]);
    }
    tc_fence_before();
    if (q0 + r < p.Nq) {
      const float inv = 1.0f / l_run;
      if (p.lse != nullptr) p.lse[((size_t)img * p.heads + head) * p.Nq + q0 + r] = m_run + log2f(l_run);
      __nv_bfloat16* orow = p.out + ((size_t)img * p.Nq + q0 + r) * p.ldo + head * p.d;
      if (p.d == AT_D) {
#pragma unroll
        for (int c = 0; c < AT_D; c += 8)
          *reinterpret_cast<uint4*>(orow + c) =
              make_uint4(pack_bf16x2(o[c] * inv, o[c + 1] * inv), pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv),
                         pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv), pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv));
      } else {
#pragma unroll
        for (int c = 0; c < AT_D; ++c)
          if (c < p.d) orow[c] = __float2bfloat16(o[c] * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
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

static int attn_tmap(CUtensorMap* tm, const void* base, int ld, int heads, int d, int N, int n_img) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)n_img};
  uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * N};
  uint32_t box[4] = {AT_D, 1, AT_BQ, 1};
  return make_tmap(tm, base, 4, dims, strides, box);
}

}  // namespace lkgd

using namespace lkgd;

static int attention_impl(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                          void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq,
                          int32_t Nk, float scale, float* lse, void* stream) {
  if (n_img <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0 || d > AT_D || d % 8 || d <= 0) return LKGD_ESHAPE;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return LKGD_EALIGN;
  if (heads > 65535 || n_img > 65535) return LKGD_ESHAPE;
  AttnParams p;
  int rc;
  if ((rc = attn_tmap(&p.tmQ, q, ldq, heads, d, Nq, n_img))) return rc;
  if ((rc = attn_tmap(&p.tmK, k, ldk, heads, d, Nk, n_img))) return rc;
  if ((rc = attn_tmap(&p.tmV, v, ldv, heads, d, Nk, n_img))) return rc;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo; p.heads = heads; p.d = d; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
    e = cudaFuncSetAttribute(attn_flash_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr = true;
  }
  dim3 grid((Nq + AT_BQ - 1) / AT_BQ, heads, n_img);
  static const bool v1 = getenv("LKGD_ATTN_V1") != nullptr;
  if (v1) attn_flash_v1_kernel<<<grid, AT_THREADS, AT_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  else attn_flash_kernel<<<grid, AT_THREADS, AT_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
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
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 3 * TAttn<64>::TILE);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr = true;
  }
  switch (d) {
    case 16: attn_temporal_kernel<16><<<grid, 128, 4 * 3 * TAttn<16>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    case 32: attn_temporal_kernel<32><<<grid, 128, 4 * 3 * TAttn<32>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    case 64: attn_temporal_kernel<64><<<grid, 128, 4 * 3 * TAttn<64>::TILE, st>>>(x, o, B, F, HW, heads, sl2); break;
    default: return LKGD_ESHAPE;
  }
  return launch_epilogue();
}
