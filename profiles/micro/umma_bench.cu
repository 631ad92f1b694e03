// umma_bench.cu — tcgen05.mma issue/throughput microbenchmark on sm_100a (SS mode, bf16, cta_group::1).
// Each CTA issues `reps` batches of `nk` MMAs (M=128, N, K=16) on garbage smem tiles and measures clock64 from the
// first issue to the commit arrival.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mlx-vae_b200/csrc -o umma_bench.bin umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace arcvae;

__global__ void __launch_bounds__(128, 1) k(int N, int nk, int reps, int b_mn, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* A = smem;            // 4 panels x 16 KB
  uint8_t* B = smem + 65536;    // 4 panels x (N x 128 B)
  for (int i = threadIdx.x; i < (65536 + 4 * 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, false, b_mn != 0);
    long long tot = 0, issue = 0;
    for (int r = 0; r < reps; r++) {
      long long t0 = clock64();
      for (int i = 0; i < nk; i++) {
        const int kp = (i >> 2) & 3, k = i & 3;
        const uint32_t a_addr = tc::smem_u32(A + kp * 16384) + 32 * k;
        const uint64_t db = b_mn ? tc::make_smem_desc(tc::smem_u32(B + kp * N * 128) + k * 2048, 8192, 1024)
                                 : tc::make_smem_desc(tc::smem_u32(B + kp * N * 128) + 32 * k, 16, 1024);
        tc::mma_bf16(tmem, tc::make_smem_desc(a_addr, 16, 1024), db, idesc, i > 0 ? 1u : 0u);
      }
      tc::mma_commit(&bar);
      long long t1 = clock64();
      tc::mbar_wait(&bar, r & 1);
      long long t2 = clock64();
      tot += t2 - t0; issue += t1 - t0;
    }
    out[blockIdx.x * 2] = tot / reps;
    out[blockIdx.x * 2 + 1] = issue / reps;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out; cudaError_t e0 = cudaMallocManaged(&out, 148 * 2 * 8);
  if (e0 != cudaSuccess) { printf("malloc %s\n", cudaGetErrorString(e0)); return 1; }
  e0 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e0 != cudaSuccess) { printf("attr %s\n", cudaGetErrorString(e0)); return 1; }
  const size_t smem = 65536 + 4 * 256 * 128 + 1024;
  int Ns[] = {256, 192, 128, 64, 32};
  for (int b_mn = 0; b_mn < 2; b_mn++)
    for (int N : Ns)
      for (int nk : {16, 64}) {
        if (b_mn && (N % 64)) continue;
        k<<<148, 128, smem>>>(N, nk, 50, b_mn, out);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        double tot = 0, iss = 0;
        for (int i = 0; i < 148; i++) { tot += out[2 * i]; iss += out[2 * i + 1]; }
        printf("b_mn=%d N=%3d nk=%2d: batch %.0f clk (%.1f clk/MMA; floor %d), issue loop %.0f clk\n", b_mn, N, nk, tot / 148,
               tot / 148 / nk, 128 * N / 256, iss / 148);
      }
  return 0;
}
