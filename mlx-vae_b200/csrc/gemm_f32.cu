// gemm_f32.cu — fp32 FFMA GEMM (reference-precision contractions).
//
// Used for every contraction when precision == ARCVAE_PREC_FP32 (the mode whose results match the fp32
// reference to ~1e-6) and, in every mode, for the small head / table / weight-gradient-of-table products
// whose shapes (K=80, N=1, ld=129 ...) are not tensor-core material.
//
// Tile: 128x128x16 per 256-thread CTA, 8x8 register micro-tile per thread arranged as four 4x4 quadrants so
// that the shared-memory reads are conflict-free float4s.  Split-K (grid.z) with fp32 atomics serves the
// weight-gradient shapes (M,N small; K = B*T).
#include "common.cuh"

namespace arcvae {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2)
gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                float* __restrict__ C, int ldc, const float* __restrict__ bias, int accumulate, RowMap rm, int kchunk) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int ty = tid / 16, tx = tid % 16;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const bool a_al = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
  const bool b_al = ((ldb & 3) == 0) && ((((uintptr_t)B) & 15) == 0);

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile -> As[k][m]
    if (!TA) {
      // K contiguous: 128 rows x 4 float4
#pragma unroll
      for (int it = 0; it < 2; it++) {
        int idx = tid + it * 256;
        int r = idx >> 2, kq = (idx & 3) * 4;
        int m = m0 + r, k = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < M) {
          const float* src = A + rm(m) * (long)lda + k;
          if (a_al && k + 3 < kend && ((k & 3) == 0)) {
            float4 q = *reinterpret_cast<const float4*>(src);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (k + j < kend) v[j] = src[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) As[kq + j][r] = v[j];
      }
    } else {
      // M contiguous: 16 k-rows x 32 float4
#pragma unroll
      for (int it = 0; it < 2; it++) {
        int idx = tid + it * 256;
        int kk = idx >> 5, mq = (idx & 31) * 4;
        int k = k0 + kk, m = m0 + mq;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          const float* src = A + (long)k * lda + m;
          if (a_al && m + 3 < M && ((m & 3) == 0)) {
            q = *reinterpret_cast<const float4*>(src);
          } else {
            if (m + 0 < M) q.x = src[0];
            if (m + 1 < M) q.y = src[1];
            if (m + 2 < M) q.z = src[2];
            if (m + 3 < M) q.w = src[3];
          }
        }
        *reinterpret_cast<float4*>(&As[kk][mq]) = q;
      }
    }
    // ---- stage B tile -> Bs[k][n]
    if (TB) {
      // B[n*ldb + k]: K contiguous
#pragma unroll
      for (int it = 0; it < 2; it++) {
        int idx = tid + it * 256;
        int r = idx >> 2, kq = (idx & 3) * 4;
        int n = n0 + r, k = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < N) {
          const float* src = B + (long)n * ldb + k;
          if (b_al && k + 3 < kend && ((k & 3) == 0)) {
            float4 q = *reinterpret_cast<const float4*>(src);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (k + j < kend) v[j] = src[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) Bs[kq + j][r] = v[j];
      }
    } else {
      // B[k*ldb + n]: N contiguous
#pragma unroll
      for (int it = 0; it < 2; it++) {
        int idx = tid + it * 256;
        int kk = idx >> 5, nq = (idx & 31) * 4;
        int k = k0 + kk, n = n0 + nq;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
          const float* src = B + (long)k * ldb + n;
          if (b_al && n + 3 < N && ((n & 3) == 0)) {
            q = *reinterpret_cast<const float4*>(src);
          } else {
            if (n + 0 < N) q.x = src[0];
            if (n + 1 < N) q.y = src[1];
            if (n + 2 < N) q.z = src[2];
            if (n + 3 < N) q.w = src[3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[kk][nq]) = q;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    float* crow = C + rm(m) * (long)ldc;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[n];
      if (split) {
        atomicAdd(crow + n, v);
      } else if (accumulate) {
        crow[n] += v;
      } else {
        crow[n] = v;
      }
    }
  }
}

// Small problems (the token-table products of layer 0 and their gradients: 80..1024 x 80..1024 x 80..1024).  With 128 x 128
// tiles they occupy 1-8 CTAs that walk K serially with unpipelined loads: ~30 us each, seven of them per training step.
// 32 x 64 tiles, BK = 32: 30-100 CTAs, 3-4 iterations each.  Same summation order per output element (sequential k).
constexpr int SBM = 32, SBN = 64, SBK = 32;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_small_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                      float* __restrict__ C, int ldc, const float* __restrict__ bias, int accumulate, RowMap rm, int kchunk) {
  __shared__ float As[SBK][SBM + 1];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = kbeg; k0 < kend; k0 += SBK) {
#pragma unroll
    for (int it = 0; it < SBM * SBK / 256; it++) {
      const int idx = tid + it * 256;
      const int r = TA ? (idx & (SBM - 1)) : (idx / SBK), kk = TA ? (idx / SBM) : (idx & (SBK - 1));
      const int m = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < kend) v = TA ? A[(long)k * lda + m] : A[rm(m) * (long)lda + k];
      As[kk][r] = v;
    }
#pragma unroll
    for (int it = 0; it < SBN * SBK / 256; it++) {
      const int idx = tid + it * 256;
      const int c = TB ? (idx / SBK) : (idx & (SBN - 1)), kk = TB ? (idx & (SBK - 1)) : (idx / SBN);
      const int n = n0 + c, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < kend) v = TB ? B[(long)n * ldb + k] : B[(long)k * ldb + n];
      Bs[kk][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; kk++) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int m = m0 + ty * 2 + i;
    if (m >= M) continue;
    float* crow = C + rm(m) * (long)ldc;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[n];
      if (split) atomicAdd(crow + n, v);
      else if (accumulate) crow[n] += v;
      else crow[n] = v;
    }
  }
}

int gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
             int ldc, const float* bias, bool accumulate, RowMap rm, int splitk, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  ARCVAE_REQUIRE(!(transA && rm.tlist != nullptr), "row map only with transA=0");
  ARCVAE_REQUIRE(splitk <= 1 || accumulate, "split-K needs accumulate semantics");
  if (splitk < 1) splitk = 1;
  if ((long)cdiv(M, BM) * cdiv(N, BN) <= 16 && (long)M * N <= 1024 * 1024 && K <= 8192 && !(transA && transB)) {
    // few 128 x 128 tiles: the small-tile kernel (more CTAs, short K loops); split K only where the caller allows atomics
    int sk = 1;
    if (accumulate && K >= 256) sk = K / 128 < 16 ? K / 128 : 16;
    int kchunk = cdiv(cdiv(K, sk), SBK) * SBK;
    sk = cdiv(K, kchunk);
    dim3 grid(cdiv(N, SBN), cdiv(M, SBM), sk);
    TimeScope ts(TIME_GEMM_F32, st);
    const int acc = accumulate ? 1 : 0;
    if (!transA && !transB) gemm_f32_small_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
    else if (!transA) gemm_f32_small_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
    else gemm_f32_small_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
    ARCVAE_LAUNCHED();
    return 0;
  }
  int kchunk = cdiv(cdiv(K, splitk), BK) * BK;
  if (kchunk <= 0) kchunk = BK;
  splitk = cdiv(K, kchunk);
  if (splitk < 1) splitk = 1;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splitk);
  dim3 block(256);
  TimeScope ts(TIME_GEMM_F32, st);
  int acc = accumulate ? 1 : 0;
  if (!transA && !transB)
    gemm_f32_kernel<false, false><<<grid, block, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
  else if (!transA && transB)
    gemm_f32_kernel<false, true><<<grid, block, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
  else if (transA && !transB)
    gemm_f32_kernel<true, false><<<grid, block, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
  else
    gemm_f32_kernel<true, true><<<grid, block, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, acc, rm, kchunk);
  ARCVAE_LAUNCHED();
  return 0;
}

// pick a split-K factor so that a weight-gradient shaped GEMM (few output tiles, long K) fills the 148 SMs
int pick_splitk(int M, int N, int K) {
  long tiles = (long)cdiv(M, BM) * cdiv(N, BN);
  if (tiles >= 148 || K <= 4 * BK) return 1;
  long want = (2 * 148 + tiles - 1) / tiles;
  long maxs = K / (8 * BK);
  if (maxs < 1) maxs = 1;
  return (int)(want < maxs ? want : maxs);
}

}  // namespace arcvae

extern "C" int arcvae_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
                               int ldb, float* C, int ldc, const float* bias, int accumulate, void* stream) {
  arcvae::RowMap rm{nullptr, 1};
  int sk = accumulate ? arcvae::pick_splitk(M, N, K) : 1;
  return arcvae::gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate != 0, rm, sk,
                          (cudaStream_t)stream);
}
