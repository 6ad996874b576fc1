// Microbenchmark: MUFU.TANH throughput, fp32 vs f16 (packed f16x2 issues two MUFU.TANH.F16).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = 0x3c003800u + threadIdx.x, h1 = h0 + 7, h2 = h0 + 13, h3 = h0 + 29;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h3));
    } else {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h3));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
template <int MODE>
void run(const char* name, int results_per_op) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148, 1024>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double results = 148.0 * 1024 * iters * 4 * results_per_op;
  printf("%s: %.3f ms, %.1f G results/s, %.2f results/clk/SM @1.9GHz\n", name, ms, results / ms / 1e6, results / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}
int main() { run<0>("tanh.approx.f32", 1); run<1>("tanh.approx.f16x2", 2); run<2>("tanh.approx.bf16x2", 2); return 0; }
