// Micro-benchmark: MUFU.EX2 throughput for f32 vs packed f16x2 / bf16x2 operands (results per clock per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 3) {  // f32 pair -> pack -> f16x2 ex2 (the softmax inner step)
        float x = __uint_as_float(a[i]), y = x + 1.f;
        uint32_t p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(y), "f"(x));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p));
        a[i] = p;
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int per_op) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t* out;
  cudaMalloc(&out, sms * 4 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<sms * 4, 512>>>(out, 100, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * 2, 1024>>>(out, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * 2 * 1024 * iters * 8 * per_op;
  printf("%-28s %.3f ms  %.2f T results/s  (%.1f results/clk/SM at %d MHz nominal)\n", name, ms, ops / ms / 1e9,
         ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.f16x2 + ex2.f16x2", 2);
  return 0;
}
