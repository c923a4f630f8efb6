// Attention backward (LoRA fine-tuning step, reference train_models/train_svd_lora.py:1683 -> autograd of
// F.scaled_dot_product_attention in the spatial / temporal transformer blocks).
//  * attn_bwd_prep_kernel : Dv[row, head] = sum_d dO * O
//  * attn_bwd_dq_kernel   : one CTA per 64 query rows of an (image, head); loops over key blocks; recomputes
//                           P = 2^(S c - lse), dP = dO V^T, dS = P (dP - Dv);  dQ = scale * dS K
//  * attn_bwd_dkv_kernel  : one CTA per 64 key rows; loops over query blocks on the TRANSPOSED problem
//                           (S^T = K Q^T, dP^T = V dO^T) so that P^T / dS^T are produced directly in the A-fragment
//                           layout of the next MMA:  dV = P^T dO,  dK = scale * dS^T Q.  No atomics, no smem round trip
//                           of P.
//  * attn_temporal_bwd_kernel : one warp per (batch, pixel, head), F <= 32 frames, both of the above in one pass.
// mma.sync m16n8k16 (bf16 in, fp32 accumulate) fed by ldmatrix from XOR-swizzled smem tiles; the problems are small
// (training runs 14 frames at 40x64 latents) and the backward is a fraction of the step, so the legacy tensor path is
// adequate here; the forward stays on tcgen05.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

template <int D>
struct BT {
  static constexpr int P = D / 8;            // 16-byte chunks per row
  static constexpr int RPL = 8 / P > 0 ? 8 / P : 1;
  static constexpr int ROW = D * 2;
  __device__ static __forceinline__ uint32_t off(int row, int chunk) {
    return static_cast<uint32_t>(row * ROW + ((chunk ^ ((row / RPL) & (P - 1))) << 4));
  }
};

__device__ __forceinline__ void b_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void b_ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void b_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void b_cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void b_cp_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void b_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b_cp_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float b_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rows [r0, r0+ROWS) of a strided [N, ld] matrix (head slice of D columns) -> swizzled smem tile; rows >= N zero-filled
template <int D, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(uint32_t smem, const __nv_bfloat16* base, long long ld, int r0, int N) {
  using T = BT<D>;
  for (int i = threadIdx.x; i < ROWS * T::P; i += THREADS) {
    const int row = i / T::P, ch = i % T::P;
    const bool ok = r0 + row < N;
    b_cp16(smem + T::off(row, ch), base + (ok ? (long long)(r0 + row) * ld + ch * 8 : 0), ok ? 16u : 0u);
  }
}

__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO, int ldo,
                                     long long rows, int heads, int d, int N, float* __restrict__ dvec) {
  // thread per (row, head); dvec layout [img][head][N]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * heads) return;
  const long long row = idx / heads;
  const int h = (int)(idx % heads);
  const __nv_bfloat16* po = o + row * ldo + h * d;
  const __nv_bfloat16* pd = dO + row * ldo + h * d;
  float s = 0.f;
  for (int c = 0; c < d; c += 8) {
    float a[8], b[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(po + c), a);
    unpack_bf16x8(*reinterpret_cast<const uint4*>(pd + c), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
  }
  const long long img = row / N;
  dvec[(img * heads + h) * N + row % N] = s;
}

constexpr int BB = 64;   // block of 64 rows (queries or keys): 4 warps x 16 rows

struct AttnBwdParams {
  const __nv_bfloat16 *q, *k, *v, *dO;
  __nv_bfloat16 *dq, *dk, *dv;
  const float *lse, *dvec;
  long long ldq, ldk, ldv, ldo, lddq, lddk, lddv;
  int heads, N;
  float scale, scale_log2;
};

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdParams p) {
  using T = BT<D>;
  constexpr int TILE = BB * T::ROW;
  extern __shared__ __align__(128) uint8_t bsm[];
  const uint32_t sQ = smem_u32(bsm), sdO = sQ + TILE, sKV = sdO + TILE;      // then [K0 | V0 | K1 | V1]: double-buffered stream
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BB, head = blockIdx.y, img = blockIdx.z;
  const long long img_row = (long long)img * p.N;
  const __nv_bfloat16* qb = p.q + img_row * p.ldq + head * D;
  const __nv_bfloat16* kb = p.k + img_row * p.ldk + head * D;
  const __nv_bfloat16* vb = p.v + img_row * p.ldv + head * D;
  const __nv_bfloat16* dob = p.dO + img_row * p.ldo + head * D;
  load_tile<D, BB, 128>(sQ, qb, p.ldq, q0, p.N);
  load_tile<D, BB, 128>(sdO, dob, p.ldo, q0, p.N);
  b_cp_wait_all();
  __syncthreads();
  // A fragments of this warp's 16 query rows, kept for the whole loop
  uint32_t aq[D / 16][4], ado[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    b_ldsm_x4(sQ + T::off(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), aq[ks]);
    b_ldsm_x4(sdO + T::off(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), ado[ks]);
  }
  const int r_lo = q0 + warp * 16 + (lane >> 2), r_hi = r_lo + 8;
  const float* lse_b = p.lse + ((long long)img * p.heads + head) * p.N;
  const float* dv_b = p.dvec + ((long long)img * p.heads + head) * p.N;
  const float lse0 = r_lo < p.N ? lse_b[r_lo] : INFINITY, lse1 = r_hi < p.N ? lse_b[r_hi] : INFINITY;
  const float dd0 = r_lo < p.N ? dv_b[r_lo] : 0.f, dd1 = r_hi < p.N ? dv_b[r_hi] : 0.f;
  float dq[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;

  const int T_kv = (p.N + BB - 1) / BB;
  load_tile<D, BB, 128>(sKV, kb, p.ldk, 0, p.N);
  load_tile<D, BB, 128>(sKV + TILE, vb, p.ldv, 0, p.N);
  b_cp_commit();
  for (int j = 0; j < T_kv; ++j) {
    b_cp_wait0();
    __syncthreads();                       // tile j has landed; every warp is done with tile j-1, whose buffer is refilled now
    if (j + 1 < T_kv) {
      const uint32_t nb = sKV + ((j + 1) & 1) * 2 * TILE;
      load_tile<D, BB, 128>(nb, kb, p.ldk, (j + 1) * BB, p.N);
      load_tile<D, BB, 128>(nb + TILE, vb, p.ldv, (j + 1) * BB, p.N);
      b_cp_commit();
    }
    const uint32_t sK = sKV + (j & 1) * 2 * TILE, sV = sK + TILE;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[n][e] = 0.f; dp[n][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        b_ldsm_x4(sK + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), b);
        b_mma(s[2 * np], aq[ks], b[0], b[1]);
        b_mma(s[2 * np + 1], aq[ks], b[2], b[3]);
        b_ldsm_x4(sV + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), b);
        b_mma(dp[2 * np], ado[ks], b[0], b[1]);
        b_mma(dp[2 * np + 1], ado[ks], b[2], b[3]);
      }
    }
    // dS = P * (dP - Dv), packed straight into A fragments (16 keys per k-step)
    uint32_t ads[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = j * BB + n * 8 + 2 * (lane & 3) + (e & 1);
        const float l = (e < 2) ? lse0 : lse1, dd = (e < 2) ? dd0 : dd1;
        const float pr = key < p.N ? b_ex2(fmaf(s[n][e], p.scale_log2, -l)) : 0.f;
        ds[e] = pr * (dp[n][e] - dd);
      }
      ads[n >> 1][(n & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
      ads[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dQ += dS K   (K tile is [key][d] = k x n row-major -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t b[4];
        b_ldsm_x4_t(sK + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), b);
        b_mma(dq[np * 2], ads[kk], b[0], b[1]);
        b_mma(dq[np * 2 + 1], ads[kk], b[2], b[3]);
      }
    }
  }
  __nv_bfloat16* dqb = p.dq + img_row * p.lddq + head * D;
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int c = n * 8 + 2 * (lane & 3);
    if (r_lo < p.N)
      *reinterpret_cast<uint32_t*>(dqb + (long long)r_lo * p.lddq + c) = pack_bf16x2(dq[n][0] * p.scale, dq[n][1] * p.scale);
    if (r_hi < p.N)
      *reinterpret_cast<uint32_t*>(dqb + (long long)r_hi * p.lddq + c) = pack_bf16x2(dq[n][2] * p.scale, dq[n][3] * p.scale);
  }
}

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdParams p) {
  using T = BT<D>;
  constexpr int TILE = BB * T::ROW;
  extern __shared__ __align__(128) uint8_t bsm[];
  const uint32_t sK = smem_u32(bsm), sV = sK + TILE, sQD = sV + TILE;       // then [Q0 | dO0 | Q1 | dO1]: double-buffered stream
  float* s_stat = reinterpret_cast<float*>(bsm + 6 * TILE);                 // [lse0 | Dv0 | lse1 | Dv1], BB floats each
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BB, head = blockIdx.y, img = blockIdx.z;
  const long long img_row = (long long)img * p.N;
  const __nv_bfloat16* qb = p.q + img_row * p.ldq + head * D;
  const __nv_bfloat16* kb = p.k + img_row * p.ldk + head * D;
  const __nv_bfloat16* vb = p.v + img_row * p.ldv + head * D;
  const __nv_bfloat16* dob = p.dO + img_row * p.ldo + head * D;
  load_tile<D, BB, 128>(sK, kb, p.ldk, k0, p.N);
  load_tile<D, BB, 128>(sV, vb, p.ldv, k0, p.N);
  b_cp_wait_all();
  __syncthreads();
  uint32_t ak[D / 16][4], av[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    b_ldsm_x4(sK + T::off(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), ak[ks]);
    b_ldsm_x4(sV + T::off(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), av[ks]);
  }
  const float* lse_b = p.lse + ((long long)img * p.heads + head) * p.N;
  const float* dv_b = p.dvec + ((long long)img * p.heads + head) * p.N;
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int n = 0; n < D / 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[n][e] = 0.f; dv[n][e] = 0.f; }

  const int T_q = (p.N + BB - 1) / BB;
  auto stream_tile = [&](int t) {
    const uint32_t nb = sQD + (t & 1) * 2 * TILE;
    load_tile<D, BB, 128>(nb, qb, p.ldq, t * BB, p.N);
    load_tile<D, BB, 128>(nb + TILE, dob, p.ldo, t * BB, p.N);
    b_cp_commit();
    if (threadIdx.x < BB) {
      const int r = t * BB + threadIdx.x;
      float* st_ = s_stat + (t & 1) * 2 * BB;
      st_[threadIdx.x] = r < p.N ? lse_b[r] : INFINITY;        // 2^(-inf) = 0: out-of-range queries contribute nothing
      st_[BB + threadIdx.x] = r < p.N ? dv_b[r] : 0.f;
    }
  };
  stream_tile(0);
  for (int j = 0; j < T_q; ++j) {
    b_cp_wait0();
    __syncthreads();                       // tile j has landed; every warp is done with tile j-1, whose buffer is refilled now
    if (j + 1 < T_q) stream_tile(j + 1);
    const uint32_t sQ = sQD + (j & 1) * 2 * TILE, sdO = sQ + TILE;
    const float* s_lse = s_stat + (j & 1) * 2 * BB;
    const float* s_dv = s_lse + BB;
    float st[8][4], dpt[8][4];       // S^T and dP^T: rows = this warp's 16 keys, columns = 64 queries
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { st[n][e] = 0.f; dpt[n][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        b_ldsm_x4(sQ + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), b);
        b_mma(st[2 * np], ak[ks], b[0], b[1]);
        b_mma(st[2 * np + 1], ak[ks], b[2], b[3]);
        b_ldsm_x4(sdO + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), b);
        b_mma(dpt[2 * np], av[ks], b[0], b[1]);
        b_mma(dpt[2 * np + 1], av[ks], b[2], b[3]);
      }
    }
    uint32_t apt[4][4], adst[4][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int qc = n * 8 + 2 * (lane & 3);
      const float l0 = s_lse[qc], l1 = s_lse[qc + 1], d0 = s_dv[qc], d1 = s_dv[qc + 1];
      float pt[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float l = (e & 1) ? l1 : l0, dd = (e & 1) ? d1 : d0;
        pt[e] = b_ex2(fmaf(st[n][e], p.scale_log2, -l));
        ds[e] = pt[e] * (dpt[n][e] - dd);
      }
      apt[n >> 1][(n & 1) * 2] = pack_bf16x2(pt[0], pt[1]);
      apt[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(pt[2], pt[3]);
      adst[n >> 1][(n & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
      adst[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dV += P^T dO ; dK += dS^T Q     (dO / Q tiles are [query][d] = k x n row-major -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t b[4];
        b_ldsm_x4_t(sdO + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), b);
        b_mma(dv[np * 2], apt[kk], b[0], b[1]);
        b_mma(dv[np * 2 + 1], apt[kk], b[2], b[3]);
        b_ldsm_x4_t(sQ + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), b);
        b_mma(dk[np * 2], adst[kk], b[0], b[1]);
        b_mma(dk[np * 2 + 1], adst[kk], b[2], b[3]);
      }
    }
  }
  const int r_lo = k0 + warp * 16 + (lane >> 2), r_hi = r_lo + 8;
  __nv_bfloat16* dkb = p.dk + img_row * p.lddk + head * D;
  __nv_bfloat16* dvb = p.dv + img_row * p.lddv + head * D;
#pragma unroll
  for (int n = 0; n < D / 8; ++n) {
    const int c = n * 8 + 2 * (lane & 3);
    if (r_lo < p.N) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)r_lo * p.lddk + c) = pack_bf16x2(dk[n][0] * p.scale, dk[n][1] * p.scale);
      *reinterpret_cast<uint32_t*>(dvb + (long long)r_lo * p.lddv + c) = pack_bf16x2(dv[n][0], dv[n][1]);
    }
    if (r_hi < p.N) {
      *reinterpret_cast<uint32_t*>(dkb + (long long)r_hi * p.lddk + c) = pack_bf16x2(dk[n][2] * p.scale, dk[n][3] * p.scale);
      *reinterpret_cast<uint32_t*>(dvb + (long long)r_hi * p.lddv + c) = pack_bf16x2(dv[n][2], dv[n][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------- temporal backward
// One warp per (batch, pixel, head); the F <= 32 frames are a 32 x D problem (rows >= F zero-filled).
//   phase 1 (rows = queries):  S, P = softmax(S c), dP = dO V^T, Dv = rowsum(P dP), dS = P (dP - Dv), dQ = c dS K
//   phase 2 (rows = keys)   :  S^T = K Q^T, P^T via the row statistics of phase 1 (through smem), dP^T = V dO^T,
//                              dV = P^T dO, dK = c dS^T Q
template <int D>
__global__ void __launch_bounds__(128) attn_temporal_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                const __nv_bfloat16* __restrict__ dO,
                                                                __nv_bfloat16* __restrict__ dqkv, int B, int F, int HW,
                                                                int heads, float scale, float scale_log2) {
  using T = BT<D>;
  constexpr int TILE = 32 * T::ROW;
  extern __shared__ __align__(128) uint8_t bsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long seq = (long long)blockIdx.x * 4 + warp;
  const long long total = (long long)B * HW * heads;
  if (seq >= total) return;
  const int head = (int)(seq % heads);
  const int pix = (int)((seq / heads) % HW);
  const int b = (int)(seq / ((long long)heads * HW));
  const int C = heads * D;
  const size_t ld = (size_t)3 * C;
  uint8_t* wsm = bsm + warp * (4 * TILE + 3 * 32 * 4);
  const uint32_t sQ = smem_u32(wsm), sK = sQ + TILE, sV = sK + TILE, sdO = sV + TILE;
  float* s_m = reinterpret_cast<float*>(wsm + 4 * TILE);   // row maximum (log2 domain)
  float* s_il = s_m + 32;                                  // 1 / row sum
  float* s_dd = s_il + 32;                                 // Dv
  const __nv_bfloat16* base = qkv + ((size_t)b * F * HW + pix) * ld + head * D;
  const __nv_bfloat16* dobase = dO + ((size_t)b * F * HW + pix) * C + head * D;
#pragma unroll
  for (int j = 0; j < T::P; ++j) {
    const int i = lane + 32 * j;
    const int row = i / T::P, ch = i % T::P;
    const bool ok = row < F;
    const __nv_bfloat16* src = base + (ok ? (size_t)row * HW * ld + ch * 8 : 0);
    const uint32_t o = T::off(row, ch);
    b_cp16(sQ + o, src, ok ? 16u : 0u);
    b_cp16(sK + o, src + C, ok ? 16u : 0u);
    b_cp16(sV + o, src + 2 * C, ok ? 16u : 0u);
    b_cp16(sdO + o, dobase + (ok ? (size_t)row * HW * C + ch * 8 : 0), ok ? 16u : 0u);
  }
  b_cp_wait_all();
  __syncwarp();
  __nv_bfloat16* obase = dqkv + ((size_t)b * F * HW + pix) * ld + head * D;

  // ---------------------------------------------------------------- phase 1: rows = queries
  {
    float s[2][4][4], dp[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[mt][nt][e] = 0.f; dp[mt][nt][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      uint32_t a[2][4], ad[2][4], kb[2][4], vb[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        b_ldsm_x4(sQ + T::off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), a[mt]);
        b_ldsm_x4(sdO + T::off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), ad[mt]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        b_ldsm_x4(sK + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), kb[np]);
        b_ldsm_x4(sV + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), vb[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          b_mma(s[mt][nt], a[mt], kb[nt >> 1][(nt & 1) * 2], kb[nt >> 1][(nt & 1) * 2 + 1]);
          b_mma(dp[mt][nt], ad[mt], vb[nt >> 1][(nt & 1) * 2], vb[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    uint32_t ads[2][2][4];
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
            if (key >= F) s[mt][nt][hr * 2 + e] = -INFINITY;
            mx = fmaxf(mx, s[mt][nt][hr * 2 + e]);
          }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float moff = mx * scale_log2;
        float l = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = b_ex2(fmaf(s[mt][nt][hr * 2 + e], scale_log2, -moff));
            s[mt][nt][hr * 2 + e] = pv;
            l += pv;
          }
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float il = 1.0f / l;
        float dd = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            s[mt][nt][hr * 2 + e] *= il;
            dd = fmaf(s[mt][nt][hr * 2 + e], dp[mt][nt][hr * 2 + e], dd);
          }
        dd += __shfl_xor_sync(0xffffffffu, dd, 1);
        dd += __shfl_xor_sync(0xffffffffu, dd, 2);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            s[mt][nt][hr * 2 + e] *= (dp[mt][nt][hr * 2 + e] - dd);      // dS
        if ((lane & 3) == 0) {
          const int row = mt * 16 + hr * 8 + (lane >> 2);
          s_m[row] = moff; s_il[row] = il; s_dd[row] = dd;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        ads[mt][kk][0] = pack_bf16x2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
        ads[mt][kk][1] = pack_bf16x2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
        ads[mt][kk][2] = pack_bf16x2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
        ads[mt][kk][3] = pack_bf16x2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
      }
    }
    float dq[2][D / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[mt][nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int np = 0; np < D / 16; ++np) {
        uint32_t kb[4];
        b_ldsm_x4_t(sK + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), kb);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          b_mma(dq[mt][np * 2], ads[mt][kk], kb[0], kb[1]);
          b_mma(dq[mt][np * 2 + 1], ads[mt][kk], kb[2], kb[3]);
        }
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int row = mt * 16 + hr * 8 + (lane >> 2);
          if (row < F)
            *reinterpret_cast<uint32_t*>(obase + (size_t)row * HW * ld + nd * 8 + 2 * (lane & 3)) =
                pack_bf16x2(dq[mt][nd][hr * 2] * scale, dq[mt][nd][hr * 2 + 1] * scale);
        }
  }
  __syncwarp();
  // ---------------------------------------------------------------- phase 2: rows = keys
  {
    float st[2][4][4], dpt[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { st[mt][nt][e] = 0.f; dpt[mt][nt][e] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      uint32_t a[2][4], av[2][4], qb[2][4], db[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        b_ldsm_x4(sK + T::off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), a[mt]);
        b_ldsm_x4(sV + T::off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), av[mt]);
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        b_ldsm_x4(sQ + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), qb[np]);
        b_ldsm_x4(sdO + T::off(np * 16 + (lane >> 4) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)), db[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          b_mma(st[mt][nt], a[mt], qb[nt >> 1][(nt & 1) * 2], qb[nt >> 1][(nt & 1) * 2 + 1]);
          b_mma(dpt[mt][nt], av[mt], db[nt >> 1][(nt & 1) * 2], db[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    uint32_t apt[2][2][4], adst[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int qc = nt * 8 + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qq = qc + (e & 1);
          const int key = mt * 16 + (e >> 1) * 8 + (lane >> 2);
          const float pt = key < F ? b_ex2(fmaf(st[mt][nt][e], scale_log2, -s_m[qq])) * s_il[qq] : 0.f;
          st[mt][nt][e] = pt;
          dpt[mt][nt][e] = pt * (dpt[mt][nt][e] - s_dd[qq]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        apt[mt][kk][0] = pack_bf16x2(st[mt][2 * kk][0], st[mt][2 * kk][1]);
        apt[mt][kk][1] = pack_bf16x2(st[mt][2 * kk][2], st[mt][2 * kk][3]);
        apt[mt][kk][2] = pack_bf16x2(st[mt][2 * kk + 1][0], st[mt][2 * kk + 1][1]);
        apt[mt][kk][3] = pack_bf16x2(st[mt][2 * kk + 1][2], st[mt][2 * kk + 1][3]);
        adst[mt][kk][0] = pack_bf16x2(dpt[mt][2 * kk][0], dpt[mt][2 * kk][1]);
        adst[mt][kk][1] = pack_bf16x2(dpt[mt][2 * kk][2], dpt[mt][2 * kk][3]);
        adst[mt][kk][2] = pack_bf16x2(dpt[mt][2 * kk + 1][0], dpt[mt][2 * kk + 1][1]);
        adst[mt][kk][3] = pack_bf16x2(dpt[mt][2 * kk + 1][2], dpt[mt][2 * kk + 1][3]);
      }
    }
    // two output passes (dV then dK) to bound register pressure
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float acc[2][D / 8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nd][e] = 0.f;
      const uint32_t sB = which == 0 ? sdO : sQ;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int np = 0; np < D / 16; ++np) {
          uint32_t bb[4];
          b_ldsm_x4_t(sB + T::off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), bb);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (which == 0) {
              b_mma(acc[mt][np * 2], apt[mt][kk], bb[0], bb[1]);
              b_mma(acc[mt][np * 2 + 1], apt[mt][kk], bb[2], bb[3]);
            } else {
              b_mma(acc[mt][np * 2], adst[mt][kk], bb[0], bb[1]);
              b_mma(acc[mt][np * 2 + 1], adst[mt][kk], bb[2], bb[3]);
            }
          }
        }
      const float sc = which == 0 ? 1.0f : scale;
      __nv_bfloat16* ob = obase + (which == 0 ? 2 * C : C);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const int row = mt * 16 + hr * 8 + (lane >> 2);
            if (row < F)
              *reinterpret_cast<uint32_t*>(ob + (size_t)row * HW * ld + nd * 8 + 2 * (lane & 3)) =
                  pack_bf16x2(acc[mt][nd][hr * 2] * sc, acc[mt][nd][hr * 2 + 1] * sc);
          }
    }
  }
}

template <int D>
static int launch_attn_bwd(const AttnBwdParams& p, int n_img, cudaStream_t st) {
  constexpr int TILE = BB * D * 2;
  const int smem_dq = 6 * TILE, smem_dkv = 6 * TILE + 4 * BB * 4;
  static DeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
  }
  dim3 grid((p.N + BB - 1) / BB, p.heads, n_img);
  attn_bwd_dq_kernel<D><<<grid, 128, smem_dq, st>>>(p);
  int rc = launch_epilogue();
  if (rc) return rc;
  attn_bwd_dkv_kernel<D><<<grid, 128, smem_dkv, st>>>(p);
  return launch_epilogue();
}

// attention_bwd_tc.cu: the same two passes on tcgen05 / TMEM (64-wide heads)
int attn_bwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* dO, int ldo,
                const float* lse, const float* dvec, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int n_img,
                int heads, int N, float scale, cudaStream_t st);

}  // namespace lkgd

using namespace lkgd;

extern "C" size_t lkgd_attention_bwd_workspace(int32_t n_img, int32_t heads, int32_t N) {
  return (size_t)n_img * heads * N * sizeof(float);
}

extern "C" int lkgd_attention_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                  const void* o, const void* dO, int32_t ldo, const float* lse, void* dq, int32_t lddq,
                                  void* dk, int32_t lddk, void* dv, int32_t lddv, int32_t n_img, int32_t heads,
                                  int32_t d, int32_t N, float scale, void* workspace, size_t ws_bytes, void* stream) {
  if (n_img <= 0 || heads <= 0 || N <= 0 || (d != 16 && d != 32 && d != 64 && d != 128)) return LKGD_ESHAPE;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || lddq % 8 || lddk % 8 || lddv % 8) return LKGD_EALIGN;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o) || !aligned16(dO) || !aligned16(dq) ||
      !aligned16(dk) || !aligned16(dv))
    return LKGD_EALIGN;
  if (heads > 65535 || n_img > 65535) return LKGD_ESHAPE;
  if (workspace == nullptr || ws_bytes < lkgd_attention_bwd_workspace(n_img, heads, N)) return LKGD_EWS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)n_img * N;
  float* dvec = reinterpret_cast<float*>(workspace);
  attn_bwd_prep_kernel<<<(unsigned)((rows * heads + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dO), ldo, rows, heads, d, N, dvec);
  int rc = launch_epilogue();
  if (rc) return rc;
  // 64-wide heads (the SVD checkpoints): tcgen05 / TMEM kernels; LKGD_ATTN_BWD_MMA=1 keeps the mma.sync pair (A/B switch)
  if (d == 64 && !getenv("LKGD_ATTN_BWD_MMA"))
    return attn_bwd_tc(q, ldq, k, ldk, v, ldv, dO, ldo, lse, dvec, dq, lddq, dk, lddk, dv, lddv, n_img, heads, N, scale, st);
  AttnBwdParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v); p.dO = reinterpret_cast<const __nv_bfloat16*>(dO);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lse = lse; p.dvec = dvec;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.heads = heads; p.N = N; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  switch (d) {
    case 16: return launch_attn_bwd<16>(p, n_img, st);
    case 32: return launch_attn_bwd<32>(p, n_img, st);
    case 128: return launch_attn_bwd<128>(p, n_img, st);   // the reference-default heads (5,10,10,20): d = 128 at level 2
    default: return launch_attn_bwd<64>(p, n_img, st);
  }
}

extern "C" int lkgd_attention_temporal_bwd(const void* qkv, const void* dO, void* dqkv, int32_t B, int32_t F, int32_t HW,
                                           int32_t heads, int32_t d, float scale, void* stream) {
  if (B <= 0 || F <= 0 || F > 32 || HW <= 0 || heads <= 0) return LKGD_ESHAPE;
  if (!aligned16(qkv) || !aligned16(dO) || !aligned16(dqkv)) return LKGD_EALIGN;
  const long long total = (long long)B * HW * heads;
  const unsigned grid = (unsigned)((total + 3) / 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(dO);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dqkv);
  const float sl2 = scale * 1.4426950408889634f;
#define TB_SMEM(D) (4 * (4 * 32 * (D) * 2 + 3 * 32 * 4))
  static DeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM(64));
    if (e != cudaSuccess) return set_cuda_error(e);
    e = cudaFuncSetAttribute(attn_temporal_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM(128));
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  switch (d) {
    case 16: attn_temporal_bwd_kernel<16><<<grid, 128, TB_SMEM(16), st>>>(x, g, o, B, F, HW, heads, scale, sl2); break;
    case 32: attn_temporal_bwd_kernel<32><<<grid, 128, TB_SMEM(32), st>>>(x, g, o, B, F, HW, heads, scale, sl2); break;
    case 64: attn_temporal_bwd_kernel<64><<<grid, 128, TB_SMEM(64), st>>>(x, g, o, B, F, HW, heads, scale, sl2); break;
    case 128: attn_temporal_bwd_kernel<128><<<grid, 128, TB_SMEM(128), st>>>(x, g, o, B, F, HW, heads, scale, sl2); break;
    default: return LKGD_ESHAPE;
  }
#undef TB_SMEM
  return launch_epilogue();
}
