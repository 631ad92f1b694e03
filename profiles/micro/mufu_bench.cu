// mufu_bench.cu — MUFU.TANH throughput on sm_100a: tanh.approx.f32 vs tanh.approx.f16x2 / bf16x2 (results per clock per SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3f + i * 0.1f; h[i] = 0x3c003800u + threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * 8 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %8.3f ms  %7.2f Gop/s  %6.2f instr/clk/SM (at %d MHz)  -> %6.2f results/clk/SM\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, per_op * ops / (ms * 1e-3) / 148 / (clk * 1e3));
  cudaFree(out);
}
int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.f16x2", 2);
  run<2>("tanh.approx.bf16x2", 2);
  run<3>("ex2.approx.ftz.f32", 1);
  run<4>("ex2.approx.f16x2", 2);
  return 0;
}
