// Throughput of ex2.approx on one SM-filling grid: f32 vs packed f16x2 / bf16x2 (results per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 8 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(v[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
      if (MODE == 3) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; ex2.approx.f16 hi, hi; mov.b32 %0, {lo, hi};}" : "+r"(v[i]));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  int sms = 148, threads = 1024, iters = 4096;
  uint32_t* out;
  cudaMalloc(&out, sms * threads * 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<sms, threads>>>(out, 16, 1);
  cudaEventRecord(a);
  k<MODE><<<sms, threads>>>(out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double instr = (double)threads * iters * 8;          // per SM
  double ns = ms * 1e6;
  printf("%-28s %8.3f ms  %6.2f instr/ns/SM  = %6.2f results/ns/SM (at 1.9 GHz: %5.1f results/clk)\n", name, ms, instr / ns,
         instr * per_instr / ns, instr * per_instr / ns / 1.9);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("2 x ex2.approx.f16", 2);
  return 0;
}
