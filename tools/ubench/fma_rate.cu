// FMA-pipe issue rates on sm_100a: FFMA vs mixed-precision FHFMA (fma.rn.f32.f16) vs packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  unsigned short hx = (unsigned short)(0x3c00 + threadIdx.x), hw = 0x3800;
  float fx = 1.0001f, fw = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(fx), "f"(fw));
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(hx), "h"(hw));
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          unsigned long long d, x, w;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a[i]), "f"(a[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(fx), "f"(fx));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(fw), "f"(fw));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(x), "l"(w));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, 16, 1.f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = (double)blocks * 256 * iters * 8 * 16;
  printf("%-8s %.3f ms  %.2f T FMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, ms, fmas / ms / 1e9,
         fmas / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("FFMA", out);
  run<1>("FHFMA", out);
  run<2>("FFMA2", out);
  return 0;
}
