// Pipe-throughput microbenchmarks for sm_100a (B200): which of MUFU.EX2 / F2FP / FFMA / FFMA2 / FMNMX / PRMT share an
// issue pipe, in lane-ops per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* clks, float seed) {
  float a[8];
  uint32_t u[8];
  unsigned long long v2[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; u[i] = __float_as_uint(a[i]); }
#pragma unroll
  for (int i = 0; i < 4; ++i) v2[i] = (unsigned long long)u[i] << 32 | u[i + 4];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                       asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); }
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v2[i & 3]) : "l"(v2[(i + 1) & 3]), "l"(v2[(i + 2) & 3]));
      if (MODE == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (MODE == 6) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (MODE == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                       asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); }
      if (MODE == 8) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 3) & 7]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7])); }
      if (MODE == 9) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (MODE == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v2[i & 3]) : "l"(v2[(i + 1) & 3]));
      if (MODE == 11) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (MODE == 12) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(u[(i + 2) & 7]));
      if (MODE == 13) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 14) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]) + (float)(v2[i & 3] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clks[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_inner, float* out, long long* clks) {
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<MODE><<<148, warps * 32>>>(out, clks, 0.5f);
    cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(out, clks, 0.5f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clks, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    double inst = (double)ITERS * 8 * ops_per_inner * warps;   // warp-instructions per SM
    printf("%-28s warps/SM %2d: %.2f warp-instr/clk/SM = %.1f lanes/clk/SM\n", name, warps, inst / c, inst * 32 / c);
  }
}

int main() {
  float* out; long long* clks;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clks, 148 * 8);
  run<0>("MUFU.EX2", 1, out, clks);
  run<1>("F2FP.BF16.PACK_AB", 1, out, clks);
  run<2>("EX2 + F2FP (2 instr)", 2, out, clks);
  run<3>("FFMA 3-reg", 1, out, clks);
  run<9>("FFMA imm", 1, out, clks);
  run<4>("FFMA2 (instr)", 1, out, clks);
  run<10>("FADD2 (instr)", 1, out, clks);
  run<5>("FMNMX", 1, out, clks);
  run<6>("PRMT", 1, out, clks);
  run<7>("EX2 + PRMT (2 instr)", 2, out, clks);
  run<8>("EX2 + FFMA (2 instr)", 2, out, clks);
  run<11>("F2FP.F16.PACK_AB", 1, out, clks);
  run<12>("HFMA2", 1, out, clks);
  run<13>("MUFU.EX2.F16x2", 1, out, clks);
  run<14>("MUFU.EX2.BF16x2", 1, out, clks);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
