// common.cuh — shared host/device helpers for libarcvae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/arcvae_b200.h"

namespace arcvae {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define ARCVAE_CUDA(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::arcvae::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +    \
                          __FILE__ + ":" + std::to_string(__LINE__));                             \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

#define ARCVAE_TRY(expr)                                                                          \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != 0) return _r;                                                                       \
  } while (0)

#define ARCVAE_REQUIRE(cond, msg)                                                                 \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::arcvae::set_error(std::string("requirement failed: ") + #cond + " — " + (msg));           \
      return 2;                                                                                   \
    }                                                                                             \
  } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define ARCVAE_LAUNCHED()                                                                         \
  do {                                                                                            \
    ::arcvae::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    ARCVAE_CUDA(cudaGetLastError());                                                              \
  } while (0)

// ---- per-device one-time state (the library is usable from several devices of one process) ----------------------
// true exactly once per (slot, current device): guards cudaFuncSetAttribute calls and other per-device set-up
enum { ONCE_GEMM_TC = 0, ONCE_CELL0, ONCE_SCATTER_F32, ONCE_SCATTER_BF16, ONCE_SAMPLER,
       ONCE_GEMM_WS, ONCE_FWD3, ONCE_BWD3, ONCE_CELL0B, ONCE_NSLOTS = 16 };
bool first_use_on_device(int slot);
int device_sm_count();            // multiprocessors of the current device (cached per device)
// STICKY per-device error flag raised by the bounded waits of the persistent kernels; the library never clears it except
// through arcvae_device_error_clear().  k_adam refuses to update while it is set.
int* device_error_flag();

// ---- optional per-category device timing (CUDA events on the launch stream; off by default) ------------------
enum { TIME_GEMM_F32 = 0, TIME_RECURRENCE = 1, TIME_LOSS = 2, TIME_ADAM = 3, TIME_GEMM_TC = 4, TIME_POINTWISE = 5,
       TIME_SAMPLER = 6, TIME_NCAT = 8 };
// executed tensor-core FLOP (2*M*N*K of every tcgen05 GEMM / recurrence launch), for bench.py's executed-FLOP fraction
void count_flops(int cat, double flop);
void timing_begin(int cat, cudaStream_t st);
void timing_end(int cat, cudaStream_t st);
struct TimeScope {
  int cat; cudaStream_t st;
  TimeScope(int c, cudaStream_t s) : cat(c), st(s) { timing_begin(cat, st); }
  ~TimeScope() { timing_end(cat, st); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t off, cap;
  Arena(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---- time-list row map: local row i -> global time-major row tlist[i / Bt] * Bt + i % Bt -------
struct RowMap {
  const int* tlist;  // device; nullptr = identity
  int Bt;
  __device__ __forceinline__ long operator()(int i) const {
    if (tlist == nullptr) return i;
    int q = i / Bt;
    return (long)tlist[q] * Bt + (i - q * Bt);
  }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// tanh with full fp32 accuracy (tanhf is ~1 ulp; the fast-math version is not used: parity is fp32)
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

// MUFU.TANH based activations (abs error ~2^-11): bf16 paths only, where every gate is rounded to bf16 (2^-9) before
// it is stored or fed back through the tensor core anyway
__device__ __forceinline__ float tanh_approx_(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_approx_(float x) { return fmaf(tanh_approx_(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- internal entry points shared between translation units ------------------------------------
// C[M,N] (+)= op(A) op(B) + bias.  transA=0: A[m*lda+k]; 1: A[k*lda+m].  transB=0: B[k*ldb+n]; 1: B[n*ldb+k].
// rows of A (transA=0 only) and of C go through `rm`.  splitk>1 -> atomicAdd into C (requires accumulate).
int gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
             int ldc, const float* bias, bool accumulate, RowMap rm, int splitk, cudaStream_t st);

int pick_splitk(int M, int N, int K);

}  // namespace arcvae
