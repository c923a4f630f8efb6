// Epilogue of the tcgen05 GEMM (included by gemm_tcgen05.cu after GemmParams / tile_row are defined).
//
// Eight epilogue warps, two per TMEM lane quarter, alternate 64-byte column chunks of the 128 x BN accumulator tile:
//   cp.async prefetch of the residual chunk into a per-warp, XOR-swizzled smem buffer (double-buffered, so one chunk
//   of HBM latency is always in flight per warp)  ->  tcgen05.ld  ->  bias (from smem) / row vector / SiLU / GEGLU /
//   scaled residuals in registers  ->  per-warp smem transpose  ->  16-byte global stores coalesced along rows.
#pragma once

namespace lkgd {

// ---- per-warp staging buffers: 32 rows x RB bytes (RB = 32 or 64), 16-byte pieces XOR-swizzled so that both the
// row-per-lane accesses and the coalesced (several lanes per row) accesses are bank-conflict free.
template <int RB>
__device__ __forceinline__ uint32_t stg_off(int row, int piece) {
  constexpr int P = RB / 16;             // pieces per row: 2 or 4
  constexpr int RPL = 8 / P;             // rows per 128-byte line
  return static_cast<uint32_t>(row * RB + ((piece ^ ((row / RPL) & (P - 1))) << 4));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf-GELU with ONE MUFU op.  gelu(x) = relu(x) - |x| h(|x|) with h(t) = 0.5 erfc(t / sqrt 2) = Phi(-t), and
// h(t) = 2^P(t): log2 h is smooth (nearly quadratic), a degree-6 polynomial on [0, 6] (Chebyshev fit, tools/fit_gelu.py)
// reproduces it to 7e-5, i.e. h to 4.8e-5 relative and gelu to 6.9e-6 absolute - 40x below the bf16 rounding of the
// product it feeds (the reference computes F.gelu(gate) with the exact erf form).  t is clamped at 6 (|x| h < 1e-8
// beyond).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = fminf(ax, 6.0f);
  float p = fmaf(2.2999249e-05f, t, -6.1149016e-04f);
  p = fmaf(p, t, 7.2001889e-03f);
  p = fmaf(p, t, -5.1208213e-02f);
  p = fmaf(p, t, -4.6122226e-01f);
  p = fmaf(p, t, -1.1502144e+00f);
  p = fmaf(p, t, -1.0000589e+00f);
  return fmaf(-ax, ex2_approx(p), fmaxf(x, 0.f));
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: one issue slot for two lanes' worth of work)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// GEGLU on NP column pairs in lock step (the Horner chains of the pairs interleave, so no FFMA waits for the previous
// one): v[i] = (a[i] + bv[i]) * gelu(g[i] + bg[i]).  a / g are accumulator bits, bv / bg point to the tile's bias in smem.
template <int NP>
__device__ __forceinline__ void geglu_pairs(const uint32_t* a, const uint32_t* g, uint32_t bv, uint32_t bg, float* v) {
  const uint64_t C6 = pk2(2.2999249e-05f, 2.2999249e-05f), C5 = pk2(-6.1149016e-04f, -6.1149016e-04f),
                 C4 = pk2(7.2001889e-03f, 7.2001889e-03f), C3 = pk2(-5.1208213e-02f, -5.1208213e-02f),
                 C2 = pk2(-4.6122226e-01f, -4.6122226e-01f), C1 = pk2(-1.1502144e+00f, -1.1502144e+00f),
                 C0 = pk2(-1.0000589e+00f, -1.0000589e+00f);
  uint64_t V[NP], T[NP], NA[NP], R[NP], P[NP];
#pragma unroll
  for (int i = 0; i < NP; i += 2) {      // two pairs per 16-byte bias load
    const float4 b4 = lds_f4(bv + 8 * i);          // shared-space loads (LDS), not generic LD
    const float4 g4 = lds_f4(bg + 8 * i);
    V[i] = add2(pk2(__uint_as_float(a[2 * i]), __uint_as_float(a[2 * i + 1])), pk2(b4.x, b4.y));
    V[i + 1] = add2(pk2(__uint_as_float(a[2 * i + 2]), __uint_as_float(a[2 * i + 3])), pk2(b4.z, b4.w));
    const uint64_t G0 = add2(pk2(__uint_as_float(g[2 * i]), __uint_as_float(g[2 * i + 1])), pk2(g4.x, g4.y));
    const uint64_t G1 = add2(pk2(__uint_as_float(g[2 * i + 2]), __uint_as_float(g[2 * i + 3])), pk2(g4.z, g4.w));
    float x0, x1, x2, x3;
    upk2(G0, x0, x1); upk2(G1, x2, x3);
    T[i] = pk2(fminf(fabsf(x0), 6.0f), fminf(fabsf(x1), 6.0f));
    T[i + 1] = pk2(fminf(fabsf(x2), 6.0f), fminf(fabsf(x3), 6.0f));
    NA[i] = pk2(fminf(x0, -x0), fminf(x1, -x1));            // -|x|
    NA[i + 1] = pk2(fminf(x2, -x2), fminf(x3, -x3));
    R[i] = pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));             // relu(x)
    R[i + 1] = pk2(fmaxf(x2, 0.f), fmaxf(x3, 0.f));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(C6, T[i], C5);
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(P[i], T[i], C4);
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(P[i], T[i], C3);
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(P[i], T[i], C2);
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(P[i], T[i], C1);
#pragma unroll
  for (int i = 0; i < NP; ++i) P[i] = fma2(P[i], T[i], C0);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float p0, p1;
    upk2(P[i], p0, p1);
    P[i] = pk2(ex2_approx(p0), ex2_approx(p1));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const uint64_t Y = mul2(V[i], fma2(NA[i], P[i], R[i]));
    upk2(Y, v[2 * i], v[2 * i + 1]);
  }
}

// Epilogue of one 128 x BN tile for one warp (32 TMEM lanes = 32 tile rows, alternate column chunks).
//   CW   accumulator columns per chunk: 32 when every operand is bf16, else 16 (a chunk is 64 B of the widest row)
//   OES  output element size (2 / 4);  RES residual element size (0 = none, 2, 4);  NRES number of residuals
// Row-per-lane accesses touch this lane's own staging row; "coalesced" accesses walk 16-byte pieces so that
// consecutive lanes cover consecutive bytes of a global row (P pieces per row, 32 / P rows per instruction).
template <int EG, int CW, int OES, int RES, int NRES, bool GEGLU>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const TileCoord& tc, int n_tile, uint32_t taddr,
                                              int lane_base, int lane, int half, uint32_t stg,
                                              const float* __restrict__ sbias) {
  constexpr int OB = CW * OES;                 // output row bytes per chunk: 64 or 32
  constexpr int PO = OB / 16;
  constexpr int RB = RES ? CW * RES : 64;      // residual row bytes per chunk
  constexpr int PR = RB / 16;
  const int bn_out = GEGLU ? p.BN / 2 : p.BN;
  const int n_cols = GEGLU ? p.N / 2 : p.N;
  const int out_col_base = n_tile * bn_out;
  // columns this tile may touch: inside the matrix (n_store) and inside the tile (a chunk may overhang bn_out)
  const int n_store = min(p.n_store > 0 ? p.n_store : n_cols, out_col_base + bn_out);
  const int nchunks = (bn_out + CW - 1) / CW;
  const long long m = tile_row(p, tc, lane_base + lane);
  const uint32_t sb_addr = smem_u32(sbias);
  const float* rv = nullptr;
  if (p.rowvec != nullptr && m >= 0)
    rv = p.rowvec + (size_t)rowvec_index(p.rv_mode, (int)m, p.rv_HW, p.rv_F, p.rv_B) * p.rv_ld;

  // coalesced-side geometry, fixed for the tile: row pointers (nullptr = row outside the tensor) and smem offsets
  char* orow[PO];
  const char* r1row[PR];
  const char* r2row[PR];
#pragma unroll
  for (int j = 0; j < PO; ++j) {
    const long long mr = tile_row(p, tc, lane_base + lane / PO + (32 / PO) * j);
    orow[j] = mr >= 0 ? reinterpret_cast<char*>(p.out) + (size_t)mr * p.ldo * OES + (lane % PO) * 16 : nullptr;
  }
#pragma unroll
  for (int j = 0; j < PR; ++j) {
    r1row[j] = r2row[j] = nullptr;
    if (NRES >= 1) {
      const long long mr = tile_row(p, tc, lane_base + lane / PR + (32 / PR) * j);
      if (mr >= 0) {
        r1row[j] = reinterpret_cast<const char*>(p.res1) + (size_t)mr * p.ldr1 * RES + (lane % PR) * 16;
        if (NRES == 2) r2row[j] = reinterpret_cast<const char*>(p.res2) + (size_t)mr * p.ldr2 * RES + (lane % PR) * 16;
      }
    }
  }
  const uint32_t o_co = stg_off<OB>(lane / PO, lane % PO);        // + 512 * j
  const uint32_t r_co = stg_off<RB>(lane / PR, lane % PR);
  const int o_pc_col = (lane % PO) * (16 / OES);                  // first column of this lane's coalesced piece
  const int r_pc_col = RES ? (lane % PR) * (16 / (RES ? RES : 4)) : 0;

  auto issue = [&](int c, uint32_t buf) {
    const int col0 = out_col_base + c * CW;
    const bool col_ok = col0 + r_pc_col < n_store;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const bool ok = col_ok && r1row[j] != nullptr;
      cp_async16(buf + r_co + 512 * j, ok ? r1row[j] + (size_t)col0 * RES : reinterpret_cast<const char*>(p.res1),
                 ok ? 16u : 0u);
      if (NRES == 2)
        cp_async16(buf + 2048 + r_co + 512 * j,
                   ok ? r2row[j] + (size_t)col0 * RES : reinterpret_cast<const char*>(p.res2), ok ? 16u : 0u);
    }
    cp_async_commit();
  };

  // Residual ring: a slot is NRES buffers of 2 KB; `dist` chunks of this warp are in flight ahead of the one being
  // consumed.  Exactly one cp.async group is committed per chunk (empty past the end), so wait_group<dist> means
  // "this chunk's residual has landed".
  const int slots = p.epi_bufs / (NRES == 2 ? 2 : 1);      // 1 / 2 (two residuals), 2 / 4 (one)
  const int dist = slots - 1;
  constexpr uint32_t SLOT = (NRES == 2 ? 2 : 1) * 2048;
  if (NRES >= 1) {
    for (int d = 0; d < dist; ++d) {
      if (half + d * EG < nchunks) issue(half + d * EG, stg + d * SLOT); else cp_async_commit();
    }
  }
  int k = 0, slot = 0, pslot = dist;       // slot of chunk k, slot of chunk k + dist
  for (int c = half; c < nchunks; c += EG, ++k) {
    const uint32_t buf = stg + slot * SLOT;
    if (NRES >= 1) {
      if (dist == 0) issue(c, buf);                        // single slot: fetch, then wait
      else if (c + dist * EG < nchunks) issue(c + dist * EG, stg + pslot * SLOT);
      else cp_async_commit();
    }
    if (++slot == slots) slot = 0;
    if (++pslot >= slots) pslot = 0;
    uint32_t a[CW], g[GEGLU ? CW : 1];
    if (CW == 16) {
      tmem_ld16(taddr + c * CW, *reinterpret_cast<uint32_t(*)[16]>(a));
      if (GEGLU) tmem_ld16(taddr + bn_out + c * CW, *reinterpret_cast<uint32_t(*)[16]>(g));
    } else {
      tmem_ld32(taddr + c * CW, *reinterpret_cast<uint32_t(*)[32]>(a));
      if (GEGLU) tmem_ld32(taddr + bn_out + c * CW, *reinterpret_cast<uint32_t(*)[32]>(g));
    }
    const int n_out = out_col_base + c * CW;       // output column
    const bool use_rv = rv != nullptr;
    float v[CW];
    tmem_ld_wait();
    // bias of the tile sits in smem (value half, then - GEGLU - the gate half at +bn_out): broadcast LDS.128
    if (GEGLU) {
#pragma unroll
      for (int j = 0; j < CW; j += 16)     // 8 pairs in lock step: +1.2 % over 4 (same-box A/B)
        geglu_pairs<8>(a + j, g + (GEGLU ? j : 0), sb_addr + (c * CW + j) * 4, sb_addr + (bn_out + c * CW + j) * 4, v + j);
    } else {
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 b4 = lds_f4(sb_addr + (c * CW + j) * 4);
        v[j] = __uint_as_float(a[j]) + b4.x;
        v[j + 1] = __uint_as_float(a[j + 1]) + b4.y;
        v[j + 2] = __uint_as_float(a[j + 2]) + b4.z;
        v[j + 3] = __uint_as_float(a[j + 3]) + b4.w;
      }
    }
    if (use_rv) {                                  // per-row vector (time embedding / context term), L1/L2 resident
      if (n_out + CW <= n_cols) {      // rv_ld % 4 == 0 and 16-byte aligned base: vector loads
#pragma unroll
        for (int j = 0; j < CW; j += 4) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rv + n_out + j));
          v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] += (n_out + j < n_cols) ? __ldg(rv + n_out + j) : 0.f;
      }
    }
    if (p.act == LKGD_ACT_SILU) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = silu_fast(v[j]);
    } else if (p.act == LKGD_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = gelu_erf_fast(v[j]);
    } else if (p.act == LKGD_ACT_QUICK_GELU) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-1.702f * v[j]));
    }
    if (p.s0 != 1.0f) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] *= p.s0;
    }
    if (NRES >= 1) {
      if (dist == 3) cp_async_wait<3>(); else if (dist == 1) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
      const uint32_t own = buf + lane * RB;
      const int sw = (lane / (8 / PR)) & (PR - 1);
#pragma unroll
      for (int r = 0; r < NRES; ++r) {
        const float scale = r == 0 ? p.s1 : p.s2;
#pragma unroll
        for (int q = 0; q < PR; ++q) {
          const uint4 u = lds128(own + r * 2048 + ((q ^ sw) << 4));
          if (RES == 4) {
            v[(4 * q) % CW] = fmaf(scale, __uint_as_float(u.x), v[(4 * q) % CW]);
            v[(4 * q + 1) % CW] = fmaf(scale, __uint_as_float(u.y), v[(4 * q + 1) % CW]);
            v[(4 * q + 2) % CW] = fmaf(scale, __uint_as_float(u.z), v[(4 * q + 2) % CW]);
            v[(4 * q + 3) % CW] = fmaf(scale, __uint_as_float(u.w), v[(4 * q + 3) % CW]);
          } else {
            float f[8];
            unpack_bf16x8(u, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[(8 * q + e) % CW] = fmaf(scale, f[e], v[(8 * q + e) % CW]);
          }
        }
      }
      __syncwarp();            // every lane has consumed its residual row before the buffer is reused for the output
    }
    if (OES == 4 && p.out2 != nullptr && m >= 0) {
      // bf16 copy of the fp32 output for the GEMM that reads this tensor next (ControlNet zero convs over the skips):
      // the owning lane writes its row's 16 columns as one 32-byte sector, straight from registers
      __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)m * p.ldo2 + n_out;
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) {
        if (n_out + 8 * q < n_store)
          *reinterpret_cast<uint4*>(o2 + 8 * q) =
              make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
      }
    }
    if (p.fast_io) {
      // stage the output chunk (row per lane), then write it out with 16-byte pieces coalesced along rows
      const uint32_t own = buf + lane * OB;
      const int sw = (lane / (8 / PO)) & (PO - 1);
#pragma unroll
      for (int q = 0; q < PO; ++q) {
        uint4 u;
        if (OES == 4) {
          u = make_uint4(__float_as_uint(v[(4 * q) % CW]), __float_as_uint(v[(4 * q + 1) % CW]),
                         __float_as_uint(v[(4 * q + 2) % CW]), __float_as_uint(v[(4 * q + 3) % CW]));
        } else {
          u = make_uint4(pack_bf16x2(v[(8 * q) % CW], v[(8 * q + 1) % CW]), pack_bf16x2(v[(8 * q + 2) % CW], v[(8 * q + 3) % CW]),
                         pack_bf16x2(v[(8 * q + 4) % CW], v[(8 * q + 5) % CW]), pack_bf16x2(v[(8 * q + 6) % CW], v[(8 * q + 7) % CW]));
        }
        sts128(own + ((q ^ sw) << 4), u);
      }
      __syncwarp();
      const bool col_ok = n_out + o_pc_col < n_store;
      if (OES == 4 && p.gn_stats != nullptr) {
        // Fused GroupNorm statistics.  After the transpose this lane holds 4 consecutive columns of 4 rows; the 8 lanes
        // with the same lane % 4 cover the warp's 32 rows of those columns.  Column sums / sums of squares are reduced
        // over them with a halving butterfly (7 shuffles) that leaves ONE of the 8 totals in each lane, which then
        // issues one fp64 atomic: 32 atomics per 32 x 16 chunk.
        float w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 us[PO];
#pragma unroll
        for (int j = 0; j < PO; ++j) us[j] = lds128(buf + o_co + 512 * j);     // all loads first (see below)
#pragma unroll
        for (int j = 0; j < PO; ++j) {
          const uint4 u = us[j];
          const bool ok = col_ok && orow[j] != nullptr;
          if (ok) *reinterpret_cast<uint4*>(orow[j] + (size_t)n_out * OES) = u;
          const float x0 = ok ? __uint_as_float(u.x) : 0.f, x1 = ok ? __uint_as_float(u.y) : 0.f,
                      x2 = ok ? __uint_as_float(u.z) : 0.f, x3 = ok ? __uint_as_float(u.w) : 0.f;
          w[0] += x0; w[1] += x1; w[2] += x2; w[3] += x3;
          w[4] = fmaf(x0, x0, w[4]); w[5] = fmaf(x1, x1, w[5]); w[6] = fmaf(x2, x2, w[6]); w[7] = fmaf(x3, x3, w[7]);
        }
#pragma unroll
        for (int n = 4, off = 16; n >= 1; n >>= 1, off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send = up ? w[i] : w[i + n];
            const float keep = up ? w[i + n] : w[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        // w[0]: quantity (lane & 16 ? sum of squares : sum) of column o_pc_col + 2 * bit3 + bit2 of this chunk
        const int col = n_out + o_pc_col + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        if (col < n_store) {
          long long frame;
          if (p.mode == LKGD_A_LINEAR) frame = tc.c1 / p.gn_rows;
          else if (p.mode == LKGD_A_CONV3X3) frame = tc.c3 + (lane_base >> p.px_shift);   // >= 32 pixels per image
          else frame = (long long)tc.c3 * p.F + tc.c2;
          if (p.mode != LKGD_A_CONV3X3 || frame < p.nimg)      // the last tile of a multi-image group may have no image
            atomicAdd(p.gn_stats + ((size_t)frame * n_cols + col) * 2 + (lane >> 4), (double)w[0]);
        }
      } else {
        // every LDS is issued before the first store: with the load inside the predicated store the compiler emitted
        // LDS -> STG -> (branch) -> LDS into the SAME registers -> STG ..., paying the shared-memory latency three times
        // per chunk and a write-after-read wait on the store's operand read (20 % of the samples of the qkv L0 launch)
        uint4 us[PO];
#pragma unroll
        for (int j = 0; j < PO; ++j) us[j] = lds128(buf + o_co + 512 * j);
#pragma unroll
        for (int j = 0; j < PO; ++j) {
          if (col_ok && orow[j] != nullptr) *reinterpret_cast<uint4*>(orow[j] + (size_t)n_out * OES) = us[j];
        }
      }
      __syncwarp();            // staging buffer free again (next prefetch may overwrite it)
    } else if (m >= 0) {
      // generic path: unaligned pitches / odd column counts - scalar stores from the owning lane
      if (OES == 4) {
        float* op = reinterpret_cast<float*>(p.out) + (size_t)m * p.ldo + n_out;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (n_out + j < n_store) op[j] = v[j];
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.ldo + n_out;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (n_out + j < n_store) op[j] = __float2bfloat16(v[j]);
      }
    }
  }
}

// runtime -> compile-time epilogue variant
template <int EG>
__device__ __forceinline__ void epilogue_dispatch(const GemmParams& p, const TileCoord& tc, int n_tile, uint32_t taddr,
                                                  int lane_base, int lane, int half, uint32_t stg,
                                                  const float* sbias) {
#define LKGD_EPI(CW, OES, RES, NRES, GG) \
  epilogue_tile<EG, CW, OES, RES, NRES, GG>(p, tc, n_tile, taddr, lane_base, lane, half, stg, sbias)
  const int nres = p.res1 == nullptr ? 0 : (p.res2 == nullptr ? 1 : 2);
  if (p.act == LKGD_ACT_GEGLU) {
    if (p.out_f32) LKGD_EPI(16, 4, 0, 0, true); else LKGD_EPI(32, 2, 0, 0, true);
  } else if (p.out_f32) {
    if (nres == 0) LKGD_EPI(16, 4, 0, 0, false);
    else if (p.res1_f32) { if (nres == 1) LKGD_EPI(16, 4, 4, 1, false); else LKGD_EPI(16, 4, 4, 2, false); }
    else { if (nres == 1) LKGD_EPI(16, 4, 2, 1, false); else LKGD_EPI(16, 4, 2, 2, false); }
  } else {
    if (nres == 0) LKGD_EPI(32, 2, 0, 0, false);
    else if (p.res1_f32) { if (nres == 1) LKGD_EPI(16, 2, 4, 1, false); else LKGD_EPI(16, 2, 4, 2, false); }
    else { if (nres == 1) LKGD_EPI(32, 2, 2, 1, false); else LKGD_EPI(32, 2, 2, 2, false); }
  }
#undef LKGD_EPI
}

}  // namespace lkgd
