// Weight-gradient GEMM of the LoRA adapters: out[i, j] += alpha * sum_m X[m, i] * Y[m, j]  (contraction over the
// token axis M, both operands token-major bf16, fp32 output accumulated with atomics across an M split).
// Replaces the autograd of lora_B(lora_A(x)) in the reference (models/lora_layer.py:437 under
// train_models/train_svd_lora.py:1683): dA = dT^T x, dB = dY^T t.  The outputs are tiny ([r, C] / [C, r]) and the
// contraction is long, so the grid is (I/64, J/64, M-splits); mma.sync m16n8k16 fed by transposed ldmatrix.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

__device__ __forceinline__ uint32_t tn_off(int row, int chunk) {      // 64 bf16 (128 B) per row, 8 chunks XOR-swizzled
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void tn_ldsm_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void tn_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile of 64 token rows x 64 columns starting at (m0, c0); out-of-range rows / 8-column chunks are zero-filled
__device__ __forceinline__ void tn_load(uint32_t smem, const __nv_bfloat16* base, long long ld, long long m0, long long m_end,
                                        int c0, int ncols) {
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const bool ok = (m0 + row < m_end) && (c0 + ch * 8 < ncols);
    const __nv_bfloat16* src = base + (ok ? (m0 + row) * ld + c0 + ch * 8 : 0);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem + tn_off(row, ch)), "l"(src),
                 "r"(ok ? 16u : 0u) : "memory");
  }
}

__global__ void __launch_bounds__(128) gemm_tn_kernel(const __nv_bfloat16* __restrict__ X, long long ldx, int I,
                                                      const __nv_bfloat16* __restrict__ Y, long long ldy, int J,
                                                      long long M, long long m_per_split, float alpha,
                                                      float* __restrict__ out, long long ldo) {
  __shared__ __align__(128) uint8_t sm[2 * 2 * 8192];     // 2 stages x (X tile, Y tile)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const long long m_begin = (long long)blockIdx.z * m_per_split;
  const long long m_end = min(m_begin + m_per_split, M);
  if (m_begin >= m_end) return;
  const uint32_t s0 = smem_u32(sm);
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
  const int iters = (int)((m_end - m_begin + 63) / 64);
  tn_load(s0, X, ldx, m_begin, m_end, i0, I);
  tn_load(s0 + 8192, Y, ldy, m_begin, m_end, j0, J);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int it = 0; it < iters; ++it) {
    const uint32_t cur = s0 + (it & 1) * 16384, nxt = s0 + ((it + 1) & 1) * 16384;
    if (it + 1 < iters) {
      tn_load(nxt, X, ldx, m_begin + (long long)(it + 1) * 64, m_end, i0, I);
      tn_load(nxt + 8192, Y, ldy, m_begin + (long long)(it + 1) * 64, m_end, j0, J);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      tn_ldsm_t(cur + tn_off(kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, warp * 2 + ((lane >> 3) & 1)), a);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        tn_ldsm_t(cur + 8192 + tn_off(kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), b);
        tn_mma(acc[np * 2], a, b[0], b[1]);
        tn_mma(acc[np * 2 + 1], a, b[2], b[3]);
      }
    }
    __syncthreads();
  }
  const int r_lo = i0 + warp * 16 + (lane >> 2), r_hi = r_lo + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = j0 + n * 8 + 2 * (lane & 3);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = (e < 2) ? r_lo : r_hi, cc = c + (e & 1);
      if (r < I && cc < J) atomicAdd(out + (long long)r * ldo + cc, alpha * acc[n][e]);
    }
  }
}

}  // namespace lkgd

using namespace lkgd;

extern "C" int lkgd_gemm_tn(const void* X, int64_t ldx, int32_t I, const void* Y, int64_t ldy, int32_t J, int64_t M,
                            float alpha, float* out, int64_t ldo, void* stream) {
  if (I <= 0 || J <= 0 || M <= 0) return LKGD_ESHAPE;
  if (ldx % 8 || ldy % 8 || !aligned16(X) || !aligned16(Y)) return LKGD_EALIGN;
  const int gi = (I + 63) / 64, gj = (J + 63) / 64;
  long long splits = (2LL * sm_count() + gi * gj - 1) / (gi * gj);
  const long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = ((M + splits - 1) / splits + 63) / 64 * 64;
  splits = (M + per - 1) / per;
  dim3 grid(gi, gj, (unsigned)splits);
  gemm_tn_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(X), ldx, I, reinterpret_cast<const __nv_bfloat16*>(Y), ldy, J, M, per, alpha,
      out, ldo);
  return launch_epilogue();
}
