// Thin 3x3 convolutions of the ControlNet condition encoder at pixel resolution
// (reference models/controlnet_sdv.py:64-119: conv3x3 Cc -> 16, SiLU, conv3x3 16 -> 16, SiLU, ... over [B*F, Cc, 576, 1024]).
//
// These layers have 2-32 channels on 0.6 M pixels per frame: as implicit GEMMs on the tcgen05 kernel their 128 x 16
// accumulator tiles are bound by the TMA box rate, not by the tensor core (profiles/r02d_thin_conv.txt: 2.4 ms / 3.1 ms at
// 28 frames, DRAM 1 %, tensor pipe 2 %).  They are HBM-bound by nature (0.25 KFLOP per output byte), so they get what a
// bandwidth-bound layer needs: one pass over the input, one over the output, arithmetic from registers / smem.
//
//  * cond_conv_in_kernel: fp32 planar input [N, Cc, H, W] (the reference's condition layout) -> bf16 channels-last
//    [N, H, W, 16] with bias + SiLU.  One thread per output pixel, SIMT fp32 (Cc <= 4: 27-36 inputs x 16 outputs), loads
//    coalesced along W, neighbours from L1.  Fuses the fp32 -> bf16 channels-last pack that preceded the GEMM version.
//  * thin_conv3x3_kernel<CIN, COUT>: bf16 channels-last [N, H, W, CIN] -> [N, H, W, COUT], stride 1, bias + SiLU, CIN / COUT
//    in {16, 32}.  A block takes a 4 x 32 pixel tile: the 6 x 34 halo tile arrives with cp.async (zero-filled outside the
//    image), the 9 x COUT x CIN weights sit in smem; each warp owns one row of 32 pixels = two m16 tiles and runs
//    9 taps x CIN/16 k-steps x COUT/8 n-tiles of mma.sync.m16n8k16 with ldmatrix fragments straight from the halo tile
//    (a pixel's channels are one or two 16-byte chunks; the chunk index is XOR-swizzled with the pixel index so that the
//    eight rows of an ldmatrix phase hit eight different bank groups).  Results leave through smem as 16-byte row pieces.
//    mma.sync, not tcgen05: a 128-row UMMA tile with N = 16 cannot amortise its set-up here, and the layer is bandwidth-bound.
#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

constexpr int TC_OUT_C = 16;

template <int CC>
__global__ void __launch_bounds__(256) cond_conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ out, int N, int H, int W) {
  __shared__ float sw[TC_OUT_C * CC * 9 + TC_OUT_C];
  for (int i = threadIdx.x; i < TC_OUT_C * CC * 9 + TC_OUT_C; i += blockDim.x)
    sw[i] = i < TC_OUT_C * CC * 9 ? w[i] : bias[i - TC_OUT_C * CC * 9];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W;
  if (pix >= total) return;
  const int wx = (int)(pix % W), hy = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  float in[CC * 9];
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    const float* xc = x + ((n * CC + c) * H) * (long long)W;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = hy + dy - 1;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = wx + dx - 1;
        in[c * 9 + dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xc + (long long)yy * W + xx) : 0.f;
      }
    }
  }
  float acc[TC_OUT_C];
#pragma unroll
  for (int o = 0; o < TC_OUT_C; ++o) {
    float a = sw[TC_OUT_C * CC * 9 + o];
#pragma unroll
    for (int k = 0; k < CC * 9; ++k) a = fmaf(in[k], sw[o * CC * 9 + k], a);   // weight [o][c][ky][kx], as nn.Conv2d stores it
    acc[o] = silu_fast(a);
  }
  uint4* o4 = reinterpret_cast<uint4*>(out + pix * TC_OUT_C);
  o4[0] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                     pack_bf16x2(acc[6], acc[7]));
  o4[1] = make_uint4(pack_bf16x2(acc[8], acc[9]), pack_bf16x2(acc[10], acc[11]), pack_bf16x2(acc[12], acc[13]),
                     pack_bf16x2(acc[14], acc[15]));
}

// ---------------------------------------------------------------------------------------------- mma.sync thin conv
__device__ __forceinline__ void tc_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void tc_ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void tc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tc_cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

constexpr int TC_TH = 4, TC_TW = 32;                 // output tile: 4 rows x 32 columns, one row per warp
constexpr int TC_HH = TC_TH + 2, TC_HW = TC_TW + 2;  // halo tile

template <int CIN, int COUT>
struct ThinCfg {
  static constexpr int PCH = CIN / 8;                                  // 16-byte chunks per input pixel: 2 or 4
  static constexpr int PIX_B = CIN * 2;                                // bytes per input pixel
  static constexpr int IN_BYTES = TC_HH * TC_HW * PIX_B;
  static constexpr int W_BYTES = 9 * COUT * CIN * 2;                   // [tap][cout][cin]
  static constexpr int OUT_BYTES = TC_TH * TC_TW * COUT * 2;
  static constexpr int SMEM = IN_BYTES + W_BYTES + OUT_BYTES;
  // chunk c of halo pixel q lives at q * PIX_B + ((c ^ swz(q)) * 16): eight consecutive pixels -> eight bank groups
  __device__ static __forceinline__ uint32_t in_off(int q, int c) {
    return static_cast<uint32_t>(q * PIX_B + ((c ^ ((q >> (PCH == 2 ? 2 : 1)) & (PCH - 1))) << 4));
  }
  // weights: row (tap, cout) of CIN elements, chunks swizzled with the row index
  __device__ static __forceinline__ uint32_t w_off(int row, int c) {
    return static_cast<uint32_t>(row * PIX_B + ((c ^ ((row >> (PCH == 2 ? 2 : 1)) & (PCH - 1))) << 4));
  }
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128) thin_conv3x3_kernel(const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ wt,   // [9][COUT][CIN]
                                                           const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ out, int N, int H, int W,
                                                           int tiles_w, int tiles_h, int silu) {
  using T = ThinCfg<CIN, COUT>;
  extern __shared__ __align__(128) uint8_t tsm[];
  const uint32_t s_in = smem_u32(tsm), s_w = s_in + T::IN_BYTES, s_out = s_w + T::W_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tx = tile % tiles_w;
  tile /= tiles_w;
  const int ty = tile % tiles_h, n = tile / tiles_h;
  const int w0 = tx * TC_TW, h0 = ty * TC_TH;
  // ---- halo tile + weights -> smem
  for (int i = threadIdx.x; i < TC_HH * TC_HW * T::PCH; i += 128) {
    const int c = i % T::PCH, q = i / T::PCH;
    const int hy = h0 + q / TC_HW - 1, wx = w0 + q % TC_HW - 1;
    const bool ok = hy >= 0 && hy < H && wx >= 0 && wx < W;
    const __nv_bfloat16* src = x + (((long long)n * H + (ok ? hy : 0)) * W + (ok ? wx : 0)) * CIN + c * 8;
    tc_cp_async16(s_in + T::in_off(q, c), src, ok ? 16u : 0u);
  }
  for (int i = threadIdx.x; i < 9 * COUT * T::PCH; i += 128) {
    const int c = i % T::PCH, row = i / T::PCH;
    tc_cp_async16(s_w + T::w_off(row, c), wt + (long long)row * CIN + c * 8, 16u);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- warp `warp` computes output row h0 + warp: pixels w0 .. w0 + 31 as two m16 tiles
  float acc[2][COUT / 8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < COUT / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
#pragma unroll
    for (int ks = 0; ks < CIN / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        // A fragment rows = 16 consecutive pixels of halo row (warp + dy), starting at halo column mt * 16 + dx;
        // lanes 0-15 address the k-low chunk of pixel (lane & 15), lanes 16-31 the k-high chunk
        const int q = (warp + dy) * TC_HW + mt * 16 + dx + (lane & 15);
        tc_ldsm_x4(s_in + T::in_off(q, ks * 2 + (lane >> 4)), a[mt]);
      }
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt) {
        // B fragment (k16 x n8, "col"): weight rows = output channels nt * 8 + (lane & 7), chunks k-low / k-high
        uint32_t b[2];
        const int row = tap * COUT + nt * 8 + (lane & 7);
        tc_ldsm_x2(s_w + T::w_off(row, ks * 2 + ((lane >> 3) & 1)), b);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) tc_mma(acc[mt][nt], a[mt], b[0], b[1]);
      }
    }
  }
  // ---- bias + SiLU -> bf16 -> smem [pixel][COUT] -> 16-byte stores
  // accumulator (mt, nt): rows (lane >> 2) and (lane >> 2) + 8 of the m-tile, columns nt * 8 + 2 * (lane & 3) + {0, 1}
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < COUT / 8; ++nt) {
      const int col = nt * 8 + 2 * (lane & 3);
      const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        float v0 = acc[mt][nt][hr * 2] + b0, v1 = acc[mt][nt][hr * 2 + 1] + b1;
        if (silu) { v0 = silu_fast(v0); v1 = silu_fast(v1); }
        const int pix = warp * TC_TW + mt * 16 + hr * 8 + (lane >> 2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_out + pix * (COUT * 2) + col * 2), "r"(pack_bf16x2(v0, v1)) : "memory");
      }
    }
  __syncthreads();
  constexpr int OCH = COUT / 8;                      // 16-byte chunks per output pixel
  for (int i = threadIdx.x; i < TC_TH * TC_TW * OCH; i += 128) {
    const int c = i % OCH, pix = i / OCH;
    const int hy = h0 + pix / TC_TW, wx = w0 + pix % TC_TW;
    if (hy < H && wx < W) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(s_out + pix * (COUT * 2) + c * 16));
      *reinterpret_cast<uint4*>(out + (((long long)n * H + hy) * W + wx) * COUT + c * 8) = v;
    }
  }
}

template <int CIN, int COUT>
static int launch_thin(const void* x, const void* w, const float* bias, void* out, int N, int H, int W, int silu,
                       cudaStream_t st) {
  using T = ThinCfg<CIN, COUT>;
  static DeviceOnce attr;
  if (attr.first() && T::SMEM > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(thin_conv3x3_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  const int tiles_w = (W + TC_TW - 1) / TC_TW, tiles_h = (H + TC_TH - 1) / TC_TH;
  const long long blocks = (long long)N * tiles_w * tiles_h;
  if (blocks > 0x7fffffffLL) return LKGD_ESHAPE;
  thin_conv3x3_kernel<CIN, COUT><<<(unsigned)blocks, 128, T::SMEM, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w), bias,
      reinterpret_cast<__nv_bfloat16*>(out), N, H, W, tiles_w, tiles_h, silu);
  return launch_epilogue();
}

}  // namespace lkgd

using namespace lkgd;

extern "C" int lkgd_cond_conv_in(const float* x, int32_t N, int32_t Cc, int32_t H, int32_t W, const float* weight,
                                 const float* bias, void* out, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Cc < 1 || Cc > 4 || x == nullptr || weight == nullptr || bias == nullptr) return LKGD_ESHAPE;
  if (!aligned16(out)) return LKGD_EALIGN;
  const long long total = (long long)N * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (Cc) {
    case 1: cond_conv_in_kernel<1><<<grid, 256, 0, st>>>(x, weight, bias, o, N, H, W); break;
    case 2: cond_conv_in_kernel<2><<<grid, 256, 0, st>>>(x, weight, bias, o, N, H, W); break;
    case 3: cond_conv_in_kernel<3><<<grid, 256, 0, st>>>(x, weight, bias, o, N, H, W); break;
    default: cond_conv_in_kernel<4><<<grid, 256, 0, st>>>(x, weight, bias, o, N, H, W); break;
  }
  return launch_epilogue();
}

extern "C" int lkgd_thin_conv3x3(const void* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                                 const void* weight, const float* bias, int32_t silu, void* out, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || x == nullptr || weight == nullptr || bias == nullptr) return LKGD_ESHAPE;
  if (!aligned16(x) || !aligned16(weight) || !aligned16(out)) return LKGD_EALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (Cin == 16 && Cout == 16) return launch_thin<16, 16>(x, weight, bias, out, N, H, W, silu, st);
  if (Cin == 32 && Cout == 32) return launch_thin<32, 32>(x, weight, bias, out, N, H, W, silu, st);
  if (Cin == 16 && Cout == 32) return launch_thin<16, 32>(x, weight, bias, out, N, H, W, silu, st);
  return LKGD_ESHAPE;
}
