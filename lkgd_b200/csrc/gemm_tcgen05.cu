// Persistent warp-specialised GEMM / implicit-GEMM convolution for sm_100a.
//   warp 0 (one lane): TMA producer      - cp.async.bulk.tensor tiles of A (2-D / 4-D box, zero-filled halo) and B
//   warp 1 (one lane): tcgen05.mma issuer - 128 x BN x 16 UMMAs, fp32 accumulators in TMEM (2 x 256 columns)
//   warps 2..9       : epilogue           - two warps per TMEM lane quarter, alternating 64-byte column chunks:
//                        cp.async prefetch of the residual chunk into per-warp swizzled smem (double-buffered),
//                        tcgen05.ld -> bias / row-vector / act / GEGLU / residual mix in registers ->
//                        per-warp smem transpose -> coalesced 16-byte global stores.
// smem: S stages x (A 128x64 bf16 = 16 KB, B BNx64 bf16 = BN*128 B), SWIZZLE_128B; S = 4 (BN = 256) .. 8;
//       + 8 x 4 KB epilogue staging.
// See include/lkgd_b200.h (lkgd_gemm) for the contract and the reference call sites it replaces.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lkgd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int EPI_BUF_BYTES = 2048;               // one chunk buffer: 32 rows x 64 B
// EG = epilogue warps per TMEM lane quarter.  EG = 4 (16 warps, 96 registers) was measured on the C3 shapes and lost
// 5-25 % everywhere (spills, twice the per-tile set-up): only EG = 2 is instantiated.
constexpr int gemm_threads(int EG) { return 64 + EG * 4 * 32; }
constexpr int BIAS_BYTES = 2 * 256 * 4;           // bias of the current / next tile (double-buffered)
constexpr int BAR_BYTES = 256;
constexpr int SMEM_LIMIT = 232448;                // 227 KB opt-in maximum per CTA
constexpr int MAX_TAPS = 9;

struct GemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmA1;
  CUtensorMap tmB1;
  int mode, M, N, BN;
  int stages, stage_bytes;
  int epi_bufs;                      // 2 KB residual / output staging buffers per epilogue warp: 2 or 4
  int kb0, ntaps, kb1, k0;           // k-blocks per tap, taps, k-blocks of segment 1, channels per tap
  int H, W, TW, TH, tw_shift, tiles_w, tiles_h;  // conv output geometry + tile patch (TW a power of two)
  int px_shift, IPT, nimg;             // conv: log2(TW * TH) pixels per image in a tile, images per tile (128 >> px_shift)
  int HW, F, tiles_p;                  // tconv
  int f_fast, bf_total;                // tconv: tiles ordered (pixel tile, frame) instead of (frame, pixel tile)
  int m_tiles, n_tiles;
  signed char tap_map[MAX_TAPS], tap_dx[MAX_TAPS], tap_dy[MAX_TAPS];
  // epilogue
  const float* bias;
  const float* rowvec;
  int rv_mode, rv_HW, rv_F, rv_B, rv_ld;
  int act;
  float s0, s1, s2;
  const void* res1;
  const void* res2;
  int ldr1, ldr2, res1_f32, res2_f32;
  void* out;
  int ldo, out_f32, n_store;
  int fast_io;                         // 1: every out / residual row segment is 16-byte addressable
  double* gn_stats;                    // fused GroupNorm statistics [frames][N][2] (nullptr = off)
  int gn_rows;                         // rows per frame image
  void* out2;                          // optional bf16 copy of an fp32 output
  int ldo2;
};

__device__ __forceinline__ int rowvec_index(int mode, int m, int HW, int F, int B) {
  switch (mode) {
    case LKGD_RV_FRAME: return m / HW;
    case LKGD_RV_FRAMEPOS: return (m / HW) % F;
    case LKGD_RV_BATCH: return m / (HW * F);
    case LKGD_RV_TCTX_0272: return ((m / (HW * F)) * HW + (m % HW)) % B;
    case LKGD_RV_BATCH_TCTX: return (m / (HW * F)) * B + ((m / (HW * F)) * HW + (m % HW)) % B;
    default: return 0;
  }
}

struct TileCoord {
  int c1, c2, c3;  // TMA coordinates of the tile origin (besides the channel coordinate)
};

__device__ __forceinline__ void tile_origin(const GemmParams& p, int m_tile, TileCoord& t) {
  if (p.mode == LKGD_A_LINEAR) {
    t.c1 = m_tile * BM; t.c2 = 0; t.c3 = 0;
  } else if (p.mode == LKGD_A_CONV3X3) {
    int per_img = p.tiles_w * p.tiles_h;
    int grp = m_tile / per_img, r = m_tile % per_img;
    t.c1 = (r % p.tiles_w) * p.TW;   // w0
    t.c2 = (r / p.tiles_w) * p.TH;   // h0
    t.c3 = grp * p.IPT;              // first image of the tile (small images: several images share a tile)
  } else {  // TCONV3: tile = (bf, tile_p); big frames: (tile_p, bf), so that the three frame taps of a pixel tile are read
            // by tiles that run back to back and meet in L2 instead of being one whole frame of traffic apart
    int bf = p.f_fast ? m_tile % p.bf_total : m_tile / p.tiles_p;
    t.c1 = (p.f_fast ? m_tile / p.bf_total : m_tile % p.tiles_p) * BM;  // p0
    t.c2 = bf % p.F;                   // f
    t.c3 = bf / p.F;                   // b
  }
}

// output row for local row r of a tile; returns -1 when the row is outside the tensor
__device__ __forceinline__ long long tile_row(const GemmParams& p, const TileCoord& t, int r) {
  if (p.mode == LKGD_A_LINEAR) {
    int m = t.c1 + r;
    return m < p.M ? m : -1;
  } else if (p.mode == LKGD_A_CONV3X3) {
    int w = t.c1 + (r & (p.TW - 1)), h = t.c2 + ((r >> p.tw_shift) & (p.TH - 1)), img = t.c3 + (r >> p.px_shift);
    return (w < p.W && h < p.H && img < p.nimg) ? ((long long)img * p.H + h) * p.W + w : -1;
  } else {
    int pp = t.c1 + r;
    return pp < p.HW ? ((long long)(t.c3 * p.F + t.c2)) * p.HW + pp : -1;
  }
}

}  // namespace lkgd

#include "gemm_epilogue.cuh"

namespace lkgd {

// CTA2 = false: one CTA per 128 x BN tile (cta_group::1).
// CTA2 = true : a cluster of two CTAs shares one 256 x BN tile (cta_group::2): CTA r holds A rows r*128.. and HALF of the
//               B tile (rows r*BN/2..); the leader (rank 0) issues 256 x BN x 16 UMMAs that read both halves, so every SM
//               pulls 16 KB + BN*64 B per k-block instead of 16 KB + BN*128 B (the main loop is L2->smem bound).
//               Accumulator rows stay in each CTA's own TMEM; both CTAs run their own epilogue.
template <bool CTA2, int EG>
__global__ void __launch_bounds__(gemm_threads(EG), 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  constexpr int EPI_WARPS = EG * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_epi = smem + p.stages * p.stage_bytes;
  float* sm_bias = reinterpret_cast<float*>(sm_epi + EPI_WARPS * p.epi_bufs * EPI_BUF_BYTES);       // 2 x 256 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_epi + EPI_WARPS * p.epi_bufs * EPI_BUF_BYTES + BIAS_BYTES);
  uint64_t* full = bars;                         // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;           // [MAX_STAGES]
  uint64_t* tfull = bars + 2 * MAX_STAGES;       // [2]
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CTA2 ? (int)cluster_ctarank() : 0;
  const int worker = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // tile stream index (pair or CTA)
  const int n_workers = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = CTA2 ? (p.m_tiles + 1) / 2 : p.m_tiles;                // 256-row (pair) or 128-row units
  const int total_tiles = m_units * p.n_tiles;
  const int kiters = p.ntaps * p.kb0 + p.kb1;
  const int S = p.stages;
  const int bn_load = CTA2 ? p.BN / 2 : p.BN;                                // B rows this CTA loads per stage

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CTA2 ? 2 * EPI_WARPS : EPI_WARPS); }
      fence_barrier_init();
      tma_prefetch_desc(&p.tmA[0]);
      tma_prefetch_desc(&p.tmB);
    }
    __syncwarp();
    if (CTA2) tmem_alloc_2cta(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // peer barriers initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && elect_one()) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of a pair)
    // One thread: every instruction here is on a dependent scalar chain, so the k loop carries NO index arithmetic -
    // taps / k-blocks are nested loops, smem and barrier addresses advance incrementally (measured: ~500 cycles per
    // k-block of divisions, table look-ups and address conversions before this rewrite, vs 320-512 cycles of MMA).
    const uint32_t tx_bytes = (A_STAGE_BYTES + bn_load * BK * 2) * (CTA2 ? 2 : 1);
    const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
    const uint32_t full_remote0 = CTA2 ? mapa_shared(full0, 0) : full0;   // where the load bytes are counted
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
    uint32_t stage = 0, phase = 0, sa = smem0;
    const bool lin = p.mode == LKGD_A_LINEAR;
    auto step = [&](const CUtensorMap* mA, int a0, int a1, int a2, int a3, const CUtensorMap* mB, int b0, int b1) {
      mbar_wait_a(empty0 + stage * 8, phase ^ 1);
      if (!CTA2 || rank == 0) mbar_expect_tx_a(full0 + stage * 8, tx_bytes);
      const uint32_t fb = full_remote0 + stage * 8;
      if (CTA2) {
        if (lin) tma_load_2d_2cta_a(sa, mA, fb, a0, a1); else tma_load_4d_2cta_a(sa, mA, fb, a0, a1, a2, a3);
        tma_load_2d_2cta_a(sa + A_STAGE_BYTES, mB, fb, b0, b1);
      } else {
        if (lin) tma_load_2d_a(sa, mA, fb, a0, a1); else tma_load_4d_a(sa, mA, fb, a0, a1, a2, a3);
        tma_load_2d_a(sa + A_STAGE_BYTES, mB, fb, b0, b1);
      }
      sa += stage_bytes;
      if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; sa = smem0; }
    };
    for (int tile = worker; tile < total_tiles; tile += n_workers) {
      const int m_unit = tile / p.n_tiles;
      const int m_tile = CTA2 ? m_unit * 2 + rank : m_unit;
      const int n0 = (tile - m_unit * p.n_tiles) * p.BN + rank * bn_load;
      TileCoord tc; tile_origin(p, m_tile, tc);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const CUtensorMap* mA = &p.tmA[p.tap_map[tap]];
        const int c1 = tc.c1 + (p.mode == LKGD_A_CONV3X3 ? p.tap_dx[tap] : 0);
        const int c2 = tc.c2 + p.tap_dy[tap];
        int kcol = 0, bcol = tap * p.k0;
#pragma unroll 1
        for (int kb = 0; kb < p.kb0; ++kb, kcol += BK, bcol += BK) step(mA, kcol, c1, c2, tc.c3, &p.tmB, bcol, n0);
      }
      int kcol = 0;
#pragma unroll 1
      for (int kb = 0; kb < p.kb1; ++kb, kcol += BK) step(&p.tmA1, kcol, tc.c1, tc.c2, tc.c3, &p.tmB1, kcol, n0);
    }
  } else if (warp == 1 && rank == 0 && elect_one()) {
    // ------------------------------------------------------------------ MMA issuer (the leader CTA of a pair)
    const uint32_t idesc = umma_idesc_bf16(p.BN, CTA2 ? 256 : 128);
    const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
    const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem));
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem) + A_STAGE_BYTES);
    const uint64_t dstep = (uint64_t)(p.stage_bytes >> 4);      // the address field counts 16-byte units
    uint32_t stage = 0, phase = 0;
    uint64_t ad = adesc0, bd = bdesc0;
    int tile_iter = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++tile_iter) {
      const int as = tile_iter & 1;
      mbar_wait_a(tempty0 + as * 8, ((tile_iter >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * 256;
      uint32_t acc = 0;
#pragma unroll 1
      for (int it = 0; it < kiters; ++it) {
        mbar_wait_a(full0 + stage * 8, phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (CTA2) umma_bf16_2cta(tmem_d, ad + 2 * k, bd + 2 * k, idesc, acc);
          else umma_bf16(tmem_d, ad + 2 * k, bd + 2 * k, idesc, acc);
          acc = 1;
        }
        if (CTA2) umma_commit_2cta_a(empty0 + stage * 8); else umma_commit_a(empty0 + stage * 8);
        ad += dstep; bd += dstep;
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1; ad = adesc0; bd = bdesc0; }
      }
      if (CTA2) umma_commit_2cta_a(tfull0 + as * 8); else umma_commit_a(tfull0 + as * 8);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogue (8 warps, 2 per TMEM lane quarter)
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;                      // 0..255 among the epilogue threads
    const int lane_base = (warp & 3) * 32;                // TMEM lanes this warp may read
    const int half = ew >> 2;                             // which column chunks it takes: half, half + EG, ...
    const uint32_t stg = smem_u32(sm_epi + ew * p.epi_bufs * EPI_BUF_BYTES);
    int tile_iter = 0;
    // The bias of a tile is fetched ONE TILE AHEAD into a register: in the epilogue-bound layers (K = 320 GEGLU / qkv
    // projections) the accumulators are ready before the epilogue gets to them, so a load issued at the top of the tile
    // was an exposed L2 round trip in front of every tile (ncu source view: 5.8 % of the samples on that one line).
    auto bias_of = [&](int tile_) -> float {
      const int n = (tile_ % p.n_tiles) * p.BN + et;
      return (p.bias != nullptr && et < 256 && et < p.BN && n < p.N) ? __ldg(p.bias + n) : 0.f;
    };
    float bias_next = worker < total_tiles ? bias_of(worker) : 0.f;
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++tile_iter) {
      const int as = tile_iter & 1;
      const int m_tile = CTA2 ? (tile / p.n_tiles) * 2 + rank : tile / p.n_tiles;
      const int n_tile = tile % p.n_tiles;
      TileCoord tc; tile_origin(p, m_tile, tc);
      // bias of this tile -> smem; buffer `as` was last read two tiles ago and every epilogue warp has passed the
      // previous tile's barrier since
      float* sb = sm_bias + as * 256;
      if (et < 256) sb[et] = bias_next;
      if (tile + n_workers < total_tiles) bias_next = bias_of(tile + n_workers);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      mbar_wait(&tfull[as], (tile_iter >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * 256 + (static_cast<uint32_t>(lane_base) << 16);
      if (m_tile < p.m_tiles)        // the odd CTA of the last pair may have no rows of its own
        epilogue_dispatch<EG>(p, tc, n_tile, taddr, lane_base, lane, half, stg, sb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_remote(mapa_shared(smem_u32(&tempty[as]), 0));   // the leader's MMA warp waits for both
        else mbar_arrive(&tempty[as]);
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    __syncwarp();
    if (CTA2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------- host side
static int choose_bn(int N, bool geglu) {
  if (geglu) return 256;
  if (const char* e = getenv("LKGD_GEMM_BN")) {       // tuning experiments only
    int bn = atoi(e);
    if (bn >= 16 && bn <= 256 && bn % 16 == 0) return bn;
  }
  int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  for (int bn = 256; bn >= 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;  // tail tile handled by TMA zero-fill + store masking
}

// Few row tiles (the lower UNet levels of a 14-frame 40x64 training clip: 560 / 2240 rows): the widest tile leaves most
// SMs without work (35 tiles for the level-3 convs).  Narrower tiles that still fit ONE wave trade MMA width for SMs.
// Cost model fitted to tools/bench_gemm.py on the train_* shapes (tile time ~ BN + 64, waves = ceil(tiles / 148)):
// level-3 conv 87 -> 44 us with BN = 64, level-2 linear 25.6 -> 21.4 us with BN = 160; the level-2 convs (100 tiles) and
// every inference shape keep their tile.
static void refine_bn(GemmParams& p, int N, bool geglu) {
  if (geglu || getenv("LKGD_GEMM_BN") || getenv("LKGD_GEMM_NO_REFINE")) return;
  if (N % 64 || (long long)p.m_tiles * p.n_tiles >= 148) return;
  static const int cand[] = {160, 128, 64};
  long long best = ((long long)p.m_tiles * p.n_tiles + 147) / 148 * (p.BN + 64);
  for (int bn : cand) {
    if (bn >= p.BN || N % bn) continue;
    const long long tiles = (long long)p.m_tiles * (N / bn);
    const long long cost = (tiles + 147) / 148 * (bn + 64);
    if (cost < best) { best = cost; p.BN = bn; p.n_tiles = N / bn; }
  }
}

static void choose_patch(int H, int W, int nimg, int& TW, int& TH, int& IPT) {
  // TW * TH * IPT = 128 rows per tile (all powers of two); minimise the padded row count, prefer wide patches (longer
  // contiguous runs) and one image per tile.  Small feature maps (9 x 16 at the bottom of the SVD UNet) put several
  // images in a tile instead of padding 9 rows to 16; at least 32 pixels per image, so that the 32 rows of an epilogue
  // warp stay inside one frame (the fused GroupNorm statistics are per frame).
  long best = -1;
  for (int ipt = 1; ipt <= 4; ipt <<= 1)
    for (int tw = 128 / ipt; tw >= 8; tw >>= 1) {
      int th = 128 / ipt / tw;
      long area = (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((nimg + ipt - 1) / ipt) * 128;
      if (best < 0 || area < best) { best = area; TW = tw; TH = th; IPT = ipt; }
    }
}

// CTA pairs (cta_group::2) halve the weight-tile traffic per SM but couple the two CTAs' epilogues to one MMA stream.
// Measured on the C3 shapes (tools/bench_gemm.py, profiles/r01c_gemm_microbench.json): they win when the main loop is
// long and wide (BN = 256 with K >= 1280: +10..30 %, any tile with K >= 2560: +10..25 %), lose on the short-K,
// epilogue-bound layers (-1..-13 %).
static bool use_cta_pairs(int m_tiles, int n_tiles, int BN, long long k_total) {
  if (getenv("LKGD_GEMM_1CTA")) return false;
  if (BN % 16) return false;
  if (getenv("LKGD_GEMM_2CTA")) return m_tiles >= 2;
  if (m_tiles < 8) return false;
  return (BN == 256 && k_total >= 1280) || k_total >= 2560;
}

static int fill_params(const lkgd_gemm_args* a, GemmParams& p, bool& cta2) {
  memset(&p, 0, sizeof(p));
  if (a->M <= 0 || a->N <= 0 || a->K0 <= 0) return LKGD_ESHAPE;
  if (a->K0 % 8 || a->K1 % 8 || a->ldb % 8 || (a->K1 && a->ldb1 % 8)) return LKGD_EALIGN;
  const bool geglu = a->act == LKGD_ACT_GEGLU;
  if (geglu && (a->N % 256)) return LKGD_ESHAPE;
  p.mode = a->a_mode; p.M = a->M; p.N = a->N;
  p.BN = choose_bn(a->N, geglu);
  p.k0 = a->K0;
  p.kb0 = (a->K0 + BK - 1) / BK;
  p.kb1 = a->K1 > 0 ? (a->K1 + BK - 1) / BK : 0;
  p.n_tiles = (a->N + p.BN - 1) / p.BN;
  int rc;
  if (a->a_mode == LKGD_A_LINEAR) {
    if (a->lda % 8 || (a->K1 && a->lda1 % 8)) return LKGD_EALIGN;
    p.ntaps = 1; p.tap_map[0] = 0;
    p.m_tiles = (a->M + BM - 1) / BM;
    uint64_t dims[2] = {(uint64_t)a->K0, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, BM};
    if ((rc = make_tmap(&p.tmA[0], a->A, 2, dims, strides, box))) return rc;
    if (a->K1) {
      uint64_t d1[2] = {(uint64_t)a->K1, (uint64_t)a->M};
      uint64_t s1[1] = {(uint64_t)a->lda1 * 2};
      if ((rc = make_tmap(&p.tmA1, a->A1, 2, d1, s1, box))) return rc;
    }
  } else if (a->a_mode == LKGD_A_CONV3X3) {
    const int s = a->stride;
    if (s != 1 && s != 2) return LKGD_ESHAPE;
    const bool pad_br = s == 2 && a->pad_br;     // F.pad(x, (0, 1, 0, 1)) + Conv2d(stride 2, padding 0)
    if (pad_br && (a->Hin < 2 || a->Win < 2)) return LKGD_ESHAPE;
    const int Ho = pad_br ? (a->Hin - 2) / 2 + 1 : (a->Hin - 1) / s + 1;
    const int Wo = pad_br ? (a->Win - 2) / 2 + 1 : (a->Win - 1) / s + 1;
    if ((long long)a->NIMG * Ho * Wo != a->M) return LKGD_ESHAPE;
    p.H = Ho; p.W = Wo;
    choose_patch(Ho, Wo, a->NIMG, p.TW, p.TH, p.IPT);
    p.nimg = a->NIMG;
    p.tw_shift = 0;
    while ((1 << p.tw_shift) < p.TW) ++p.tw_shift;
    p.px_shift = 0;
    while ((1 << p.px_shift) < p.TW * p.TH) ++p.px_shift;
    p.tiles_w = (Wo + p.TW - 1) / p.TW; p.tiles_h = (Ho + p.TH - 1) / p.TH;
    p.m_tiles = ((a->NIMG + p.IPT - 1) / p.IPT) * p.tiles_w * p.tiles_h;
    p.ntaps = 9;
    uint32_t box[4] = {BK, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.IPT};
    const uint64_t C = a->K0;
    if (s == 1) {
      uint64_t dims[4] = {C, (uint64_t)a->Win, (uint64_t)a->Hin, (uint64_t)a->NIMG};
      uint64_t strides[3] = {C * 2, C * 2 * a->Win, C * 2 * a->Win * a->Hin};
      if ((rc = make_tmap(&p.tmA[0], a->A, 4, dims, strides, box))) return rc;
      for (int t = 0; t < 9; ++t) { p.tap_map[t] = 0; p.tap_dx[t] = t % 3 - 1; p.tap_dy[t] = t / 3 - 1; }
    } else {
      // four parity planes of the input; tap (ky,kx) reads plane ((ky+1)&1, (kx+1)&1) at offset (ky==0?-1:0)
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          uint64_t dims[4] = {C, (uint64_t)((a->Win - px + 1) / 2), (uint64_t)((a->Hin - py + 1) / 2),
                              (uint64_t)a->NIMG};
          if (dims[1] == 0 || dims[2] == 0) return LKGD_ESHAPE;
          uint64_t strides[3] = {C * 2 * 2, C * 2 * a->Win * 2, C * 2 * a->Win * a->Hin};
          const char* base = reinterpret_cast<const char*>(a->A) + ((size_t)py * a->Win + px) * C * 2;
          if ((rc = make_tmap(&p.tmA[py * 2 + px], base, 4, dims, strides, box))) return rc;
        }
      for (int t = 0; t < 9; ++t) {
        int ky = t / 3, kx = t % 3;
        if (pad_br) {   // input row 2 y + ky: plane ky & 1 at plane row y + (ky >> 1)
          p.tap_map[t] = (signed char)(((ky & 1) * 2) + (kx & 1));
          p.tap_dy[t] = ky == 2 ? 1 : 0; p.tap_dx[t] = kx == 2 ? 1 : 0;
        } else {
          p.tap_map[t] = (signed char)((((ky + 1) & 1) * 2) + ((kx + 1) & 1));
          p.tap_dy[t] = ky == 0 ? -1 : 0; p.tap_dx[t] = kx == 0 ? -1 : 0;
        }
      }
    }
    if (a->K1) {  // centre-tap segment over an [NIMG, Ho, Wo, K1] tensor
      uint64_t C1 = a->K1;
      uint64_t dims[4] = {C1, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)a->NIMG};
      uint64_t strides[3] = {C1 * 2, C1 * 2 * Wo, C1 * 2 * Wo * Ho};
      if ((rc = make_tmap(&p.tmA1, a->A1, 4, dims, strides, box))) return rc;
    }
  } else if (a->a_mode == LKGD_A_TCONV3) {
    if ((long long)a->NIMG * a->F * a->HW != a->M) return LKGD_ESHAPE;
    p.HW = a->HW; p.F = a->F;
    p.tiles_p = (a->HW + BM - 1) / BM;
    p.m_tiles = a->NIMG * a->F * p.tiles_p;
    p.bf_total = a->NIMG * a->F;
    // one frame of A above 16 MB (the VAE decoder's 144x256 ... 576x1024 levels; the UNet's largest is 5.9 MB): with the
    // outputs and residuals written in between, the next frame's taps no longer find this frame in the 126 MB L2
    p.f_fast = (long long)a->HW * a->K0 * 2 > (16ll << 20) && !getenv("LKGD_TCONV_P_FAST");
    p.ntaps = 3;
    for (int t = 0; t < 3; ++t) { p.tap_map[t] = 0; p.tap_dx[t] = 0; p.tap_dy[t] = t - 1; }
    uint32_t box[4] = {BK, BM, 1, 1};
    const uint64_t C = a->K0;
    uint64_t dims[4] = {C, (uint64_t)a->HW, (uint64_t)a->F, (uint64_t)a->NIMG};
    uint64_t strides[3] = {C * 2, C * 2 * a->HW, C * 2 * a->HW * a->F};
    if ((rc = make_tmap(&p.tmA[0], a->A, 4, dims, strides, box))) return rc;
    if (a->K1) {
      uint64_t C1 = a->K1;
      uint64_t d1[4] = {C1, (uint64_t)a->HW, (uint64_t)a->F, (uint64_t)a->NIMG};
      uint64_t s1[3] = {C1 * 2, C1 * 2 * a->HW, C1 * 2 * a->HW * a->F};
      if ((rc = make_tmap(&p.tmA1, a->A1, 4, d1, s1, box))) return rc;
    }
  } else {
    return LKGD_ESHAPE;
  }
  refine_bn(p, a->N, geglu);
  cta2 = use_cta_pairs(p.m_tiles, p.n_tiles, p.BN, (long long)p.ntaps * a->K0 + a->K1);
  const int bn_load = cta2 ? p.BN / 2 : p.BN;
  {
    uint64_t dims[2] = {(uint64_t)p.ntaps * a->K0, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->ldb * 2};
    uint32_t box[2] = {BK, (uint32_t)bn_load};
    if ((rc = make_tmap(&p.tmB, a->Bw, 2, dims, strides, box))) return rc;
    if (a->K1) {
      uint64_t d1[2] = {(uint64_t)a->K1, (uint64_t)a->N};
      uint64_t s1[1] = {(uint64_t)a->ldb1 * 2};
      if ((rc = make_tmap(&p.tmB1, a->Bw1, 2, d1, s1, box))) return rc;
    }
  }
  p.bias = a->bias; p.rowvec = a->rowvec;
  p.rv_mode = a->rowvec ? a->rv_mode : LKGD_RV_NONE;
  p.rv_HW = a->rv_HW > 0 ? a->rv_HW : 1; p.rv_F = a->rv_F > 0 ? a->rv_F : 1; p.rv_B = a->rv_B > 0 ? a->rv_B : 1;
  p.rv_ld = a->rv_ld > 0 ? a->rv_ld : (geglu ? a->N / 2 : a->N);
  if (a->rowvec && p.rv_ld % 4) return LKGD_EALIGN;
  p.act = a->act; p.s0 = a->s0; p.s1 = a->s1; p.s2 = a->s2;
  p.res1 = a->res1; p.ldr1 = a->ldr1; p.res1_f32 = a->res1_f32;
  p.res2 = a->res2; p.ldr2 = a->ldr2; p.res2_f32 = a->res2_f32;
  p.out = a->out; p.ldo = a->ldo; p.out_f32 = a->out_f32; p.n_store = a->n_store;
  p.stage_bytes = A_STAGE_BYTES + bn_load * BK * 2;
  // Residual prefetch depth.  A warp keeps (epi_bufs - 1) chunks of 2 KB in flight; with 8 warps and one chunk each the
  // residual stream of the short-K layers is latency-bound (16 KB in flight per SM against ~35 KB needed for a 148th of
  // the HBM bandwidth).  Four buffers per warp where the main loop is short (K <= 1024: +4..9 % measured in-process
  // with tools/bench_gemm.py --ab LKGD_GEMM_SHALLOW); longer loops prefer the extra operand stage (-4..11 %).
  p.epi_bufs = 2;
  if (a->res1 != nullptr && (long long)p.ntaps * a->K0 + a->K1 <= 1024 && !getenv("LKGD_GEMM_SHALLOW")) p.epi_bufs = 4;
  p.stages = (SMEM_LIMIT - 1024 - BAR_BYTES - BIAS_BYTES - 8 * p.epi_bufs * EPI_BUF_BYTES) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  {
    // 16-byte addressable row segments everywhere -> staged, coalesced epilogue I/O
    const int oes = a->out_f32 ? 4 : 2;
    const int n_cols = geglu ? a->N / 2 : a->N;
    const int n_store = a->n_store > 0 ? a->n_store : n_cols;
    bool ok = aligned16(a->out) && ((size_t)a->ldo * oes) % 16 == 0 && (n_store * oes) % 16 == 0;
    if (a->res1) { const int es = a->res1_f32 ? 4 : 2; ok = ok && aligned16(a->res1) && ((size_t)a->ldr1 * es) % 16 == 0 && (n_store * es) % 16 == 0; }
    if (a->res2) { const int es = a->res2_f32 ? 4 : 2; ok = ok && aligned16(a->res2) && ((size_t)a->ldr2 * es) % 16 == 0 && (n_store * es) % 16 == 0; }
    if ((a->bias && !aligned16(a->bias)) || (a->rowvec && !aligned16(a->rowvec))) return LKGD_EALIGN;
    p.fast_io = ok ? 1 : 0;
  }
  p.out2 = a->out2; p.ldo2 = a->ldo2;
  if (a->out2 != nullptr) {
    const int n_cols = geglu ? a->N / 2 : a->N;
    const int n_store = a->n_store > 0 ? a->n_store : n_cols;
    if (!a->out_f32 || geglu || n_store % 8) return LKGD_ESHAPE;
    if (!aligned16(a->out2) || a->ldo2 % 8 || a->ldo2 < n_store) return LKGD_EALIGN;
  }
  p.gn_stats = a->gn_stats; p.gn_rows = a->gn_rows;
  if (a->gn_stats != nullptr) {
    // fp32 output through the staged path, every tile inside one frame image
    if (!a->out_f32 || geglu || !p.fast_io || a->gn_rows <= 0 || a->M % a->gn_rows) return LKGD_ESHAPE;
    if (a->a_mode == LKGD_A_LINEAR && a->gn_rows % BM) return LKGD_ESHAPE;
    if (a->a_mode == LKGD_A_CONV3X3 && (long long)p.H * p.W != a->gn_rows) return LKGD_ESHAPE;
    if (a->a_mode == LKGD_A_TCONV3 && a->HW != a->gn_rows) return LKGD_ESHAPE;
  }
  return LKGD_OK;
}

}  // namespace lkgd

using namespace lkgd;

template <bool CTA2, int EG>
static int launch_gemm(const GemmParams& p, void* stream) {
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<CTA2, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  const int smem_bytes = 1024 + p.stages * p.stage_bytes + EG * 4 * p.epi_bufs * EPI_BUF_BYTES + BIAS_BYTES + BAR_BYTES;
  const int sms = sm_count();
  if (CTA2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(gemm_threads(EG));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // persistent pairs: never launch more clusters than can be co-resident (a GPC with an odd SM count strands one SM)
    static int max_pairs = 0;
    if (max_pairs == 0) {
      cfg.gridDim = dim3(sms);
      cfg.dynamicSmemBytes = SMEM_LIMIT;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<CTA2, EG>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sms / 2 - 4;
      }
      max_pairs = n;
      cfg.dynamicSmemBytes = smem_bytes;
    }
    int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    if (pairs > max_pairs) pairs = max_pairs;
    cfg.gridDim = dim3(2 * pairs);
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<CTA2, EG>, p);
    if (e != cudaSuccess) return set_cuda_error(e);
    return launch_epilogue();
  }
  int grid = p.m_tiles * p.n_tiles;
  if (grid > sms) grid = sms;
  gemm_tcgen05_kernel<CTA2, EG><<<grid, gemm_threads(EG), smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return launch_epilogue();
}

extern "C" int lkgd_gemm(const lkgd_gemm_args* a, void* stream) {
  if (a == nullptr || a->A == nullptr || a->Bw == nullptr || a->out == nullptr) return LKGD_ESHAPE;
  GemmParams p;
  bool cta2 = false;
  int rc = fill_params(a, p, cta2);
  if (rc) return rc;
  if (!p.fast_io && (a->res1 || a->res2)) return LKGD_EALIGN;   // residual rows must be 16-byte addressable
  if (a->res2 && (!a->res1 || (a->res1_f32 != 0) != (a->res2_f32 != 0))) return LKGD_ESHAPE;   // res2 needs res1 of the same dtype
  return cta2 ? launch_gemm<true, 2>(p, stream) : launch_gemm<false, 2>(p, stream);
}
