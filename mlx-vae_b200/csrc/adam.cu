// adam.cu — mlx.optimizers.Adam as the reference uses it (trainer.py:75-76, :320-324): betas (0.9, 0.999),
// eps 1e-8, NO bias correction.  One launch over the flat parameter buffer (encoder and decoder optimizers
// have identical hyper-parameters, so two optimizer objects == one pass).
// HBM-bound: 7 x 4 B per parameter (p, g, m, v read; p, m, v written) = 64.7 MB per step at 2,309,200 parameters.
#include "kernels.cuh"

namespace arcvae {

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps,
                                              float gscale, const int* __restrict__ err) {
  // a persistent kernel of this step timed out (sticky device flag): its gradients are garbage, keep the weights
  if (err != nullptr && *err != 0) return;
  size_t n4 = n >> 2;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define ADAM1(c)                                   \
  {                                                \
    float gx = gg.c * gscale;                      \
    mm.c = b1 * mm.c + (1.f - b1) * gx;            \
    vv.c = b2 * vv.c + (1.f - b2) * gx * gx;       \
    pp.c = pp.c - lr * mm.c / (sqrtf(vv.c) + eps); \
  }
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gx = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gx;
    float vi = b2 * v[i] + (1.f - b2) * gx * gx;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr * mi / (sqrtf(vi) + eps);
  }
}

__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float x = g[i];
    acc += (double)x * (double)x;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
    r = warp_sum_d(r);
    if (threadIdx.x == 0) atomicAdd(out, r);
  }
}

}  // namespace arcvae

extern "C" int arcvae_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                                float beta2, float eps, float grad_scale, void* stream) {
  using namespace arcvae;
  if (n == 0) return 0;
  ARCVAE_REQUIRE(((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0,
                 "adam buffers must be 16-byte aligned");
  size_t work = (n >> 2) + 1;
  int grid = (int)((work + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  TimeScope ts(TIME_ADAM, (cudaStream_t)stream);
  k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, grad_scale, device_error_flag());
  ARCVAE_LAUNCHED();
  return 0;
}

extern "C" int arcvae_sumsq(const float* g, size_t n, double* out, void* stream) {
  using namespace arcvae;
  if (n == 0) return 0;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_sumsq<<<grid, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  ARCVAE_LAUNCHED();
  return 0;
}
