// ll_latency.cu — one-way latency of a flag-in-data exchange through L2 between two SMs of a B200 (the mechanism of
// lstm_bwd3_kernel / lstm_fwd3_kernel): thread 0 of CTA a stores a flagged 16-byte vector (st.relaxed.gpu), thread 0 of CTA b
// polls it (ld.relaxed.gpu) and answers the same way.  Ping-pong round trip / 2 = one-way latency.
//   mode 0: idle GPU (only the two CTAs)
//   mode 1: the other 7 warps of both CTAs stream coalesced 512-byte stores to HBM at full rate (the tape traffic)
//   mode 2: the sender itself issues a burst of 24 x 16-byte stores (its warp: 24 x 512 B) right before every flag store
//   mode 3: mode 1 + every other SM streams stores too (whole-GPU write traffic)
//   mode 4: like 1 but the background is LOADS (coalesced 512 B, streaming from HBM)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_latency ll_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ll_store(uint4* p, uint32_t f) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(f) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256, 1)
k(uint4* box, uint4* sink, long sink_vecs, int iters, int mode, volatile int* stop, long long* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x >= 2) {
    // background SMs (mode 3): stream stores until told to stop
    if (mode != 3) return;
    uint4* base = sink + ((long)blockIdx.x * 8 + warp) * (sink_vecs / (gridDim.x * 8));
    const long n = sink_vecs / (gridDim.x * 8) / 32;
    for (long i = 0, c = 0; !*stop && c < (1L << 24); i = (i + 1) % n, c++) base[i * 32 + lane] = make_uint4(1, 2, 3, 4);
    return;
  }
  if (warp > 0) {
    if (mode == 1 || mode == 3) {
      uint4* base = sink + ((long)blockIdx.x * 8 + warp) * (sink_vecs / (gridDim.x * 8));
      const long n = sink_vecs / (gridDim.x * 8) / 32;
      for (long i = 0, c = 0; !*stop && c < (1L << 24); i = (i + 1) % n, c++) base[i * 32 + lane] = make_uint4(1, 2, 3, 4);
    } else if (mode == 4) {
      const uint4* base = sink + ((long)blockIdx.x * 8 + warp) * (sink_vecs / (gridDim.x * 8));
      const long n = sink_vecs / (gridDim.x * 8) / 32;
      uint32_t acc = 0;
      for (long i = 0, c = 0; !*stop && c < (1L << 24); i = (i + 1) % n, c++) { uint4 v = __ldcs(base + i * 32 + lane); acc += v.x; }
      if (acc == 0x12345678u) out[8] = acc;
    }
    return;
  }
  // warp 0 of CTA 0 / CTA 1: ping-pong
  uint4* mine = box + blockIdx.x * 64;           // inbox of this CTA
  uint4* peer = box + (1 - blockIdx.x) * 64;
  uint4* burst = sink + (sink_vecs - 2 * 24 * 32) + blockIdx.x * 24 * 32;
  long long t0 = 0, t1 = 0;
  for (int it = 1; it <= iters; it++) {
    if (it == 16) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (blockIdx.x == 0) {
      if (mode == 2)
        for (int j = 0; j < 24; j++) burst[j * 32 + lane] = make_uint4(it, j, 0, 0);
      if (lane == 0) {
        ll_store(peer, (uint32_t)it);
        for (long c = 0; ll_load(mine).w != (uint32_t)it && c < (1L << 22); c++) {}
      }
    } else {
      if (lane == 0) for (long c = 0; ll_load(mine).w != (uint32_t)it && c < (1L << 22); c++) {}
      if (mode == 2)
        for (int j = 0; j < 24; j++) burst[j * 32 + lane] = make_uint4(it, j, 0, 0);
      if (lane == 0) ll_store(peer, (uint32_t)it);
    }
    __syncwarp();
  }
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (lane == 0) {
    out[blockIdx.x] = (t1 - t0) / (iters - 15);
    __threadfence();
    if (blockIdx.x == 0) *stop = 1;
  }
}

int main() {
  uint4 *box, *sink;
  int* stop;
  long long* out;
  const long sink_vecs = (1L << 30) / 16;          // 1 GB: the background traffic misses L2
  cudaMalloc(&box, 128 * 16);
  cudaMalloc(&sink, sink_vecs * 16);
  cudaMalloc(&stop, 4);
  cudaMalloc(&out, 16 * 8);
  const char* names[5] = {"idle", "7 warps of both SMs streaming stores", "24-store burst before each flag store",
                          "every SM streaming stores", "7 warps of both SMs streaming loads"};
  for (int mode = 0; mode < 5; mode++) {
    cudaMemset(box, 0, 128 * 16);
    cudaMemset(stop, 0, 4);
    cudaMemset(out, 0, 16 * 8);
    const int grid = mode == 3 ? 148 : 2;
    k<<<grid, 256>>>(box, sink, sink_vecs, 2000, mode, stop, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): round trip %lld ns -> one-way %.0f ns  [%s]\n", mode, names[mode], h[0], h[0] / 2.0, cudaGetErrorString(e));
  }
  return 0;
}
