// pointwise.cu — element-wise / gather / scatter / reduction kernels of the AR-CVAE step.
// All are HBM-bound: coalesced along the hidden/gate axis, grid-stride loops sized to the 148 SMs.
#include <cstdlib>

#include "kernels.cuh"

namespace arcvae {

static inline int grid_for(long n, int block, int per_sm = 8) {
  long g = (n + block - 1) / block;
  long cap = 148L * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_transpose_tokens(const int32_t* __restrict__ x, int B, int T, int32_t* __restrict__ xT) {
  __shared__ int32_t tile[32][33];
  int b0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int b = b0 + i, t = t0 + threadIdx.x;
    if (b < B && t < T) tile[i][threadIdx.x] = x[(long)b * T + t];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, b = b0 + threadIdx.x;
    if (b < B && t < T) xT[(long)t * B + b] = tile[threadIdx.x][i];
  }
}
int transpose_tokens(const int32_t* x, int B, int T, int32_t* xT, cudaStream_t st) {
  dim3 grid(cdiv(B, 32), cdiv(T, 32)), block(32, 8);
  k_transpose_tokens<<<grid, block, 0, st>>>(x, B, T, xT);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ tok, int R, int N,
                              float* __restrict__ out) {
  long total = (long)R * N;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / N;
    int n = (int)(i - r * N);
    out[i] = table[(long)tok[r] * N + n];
  }
}
int gather_rows(const float* table, const int32_t* tok, int R, int N, float* out, cudaStream_t st) {
  k_gather_rows<<<grid_for((long)R * N, 256, 16), 256, 0, st>>>(table, tok, R, N, out);
  ARCVAE_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_lstm_cell_fwd(float* __restrict__ gates, const float* __restrict__ c_prev, float* __restrict__ c,
                                float* __restrict__ h, __nv_bfloat16* __restrict__ hb, int Bn, int H) {
  long total = (long)Bn * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long b = idx / H;
    int j = (int)(idx - b * H);
    float* g = gates + b * 4L * H;
    float i_ = sigmoidf_(g[j]);
    float f_ = sigmoidf_(g[H + j]);
    float g_ = tanhf_(g[2 * H + j]);
    float o_ = sigmoidf_(g[3 * H + j]);
    float cn = i_ * g_;
    if (c_prev != nullptr) cn = fmaf(f_, c_prev[idx], cn);
    g[j] = i_; g[H + j] = f_; g[2 * H + j] = g_; g[3 * H + j] = o_;
    c[idx] = cn;
    float hv = o_ * tanhf_(cn);
    h[idx] = hv;
    if (hb != nullptr) hb[idx] = __float2bfloat16(hv);
  }
}
int lstm_cell_fwd(float* gates, const float* c_prev, float* c, float* h, __nv_bfloat16* hb, int Bn, int H,
                  cudaStream_t st) {
  k_lstm_cell_fwd<<<grid_for((long)Bn * H, 256), 256, 0, st>>>(gates, c_prev, c, h, hb, Bn, H);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_lstm_cell_bwd(float* __restrict__ gates, const float* __restrict__ c, const float* __restrict__ c_prev,
                                const float* __restrict__ dh_ext, const float* __restrict__ dh_rec,
                                float* __restrict__ dc, __nv_bfloat16* __restrict__ dAb, int Bn, int H) {
  long total = (long)Bn * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long b = idx / H;
    int j = (int)(idx - b * H);
    float* g = gates + b * 4L * H;
    float i_ = g[j], f_ = g[H + j], g_ = g[2 * H + j], o_ = g[3 * H + j];
    float dh = 0.f;
    if (dh_ext != nullptr) dh += dh_ext[idx];
    if (dh_rec != nullptr) dh += dh_rec[idx];
    float tc = tanhf_(c[idx]);
    float dct = dc[idx] + dh * o_ * (1.f - tc * tc);
    float cp = (c_prev != nullptr) ? c_prev[idx] : 0.f;
    float d_o = dh * tc;
    float d_i = dct * g_;
    float d_g = dct * i_;
    float d_f = dct * cp;
    float a_i = d_i * i_ * (1.f - i_), a_f = d_f * f_ * (1.f - f_), a_g = d_g * (1.f - g_ * g_),
          a_o = d_o * o_ * (1.f - o_);
    g[j] = a_i; g[H + j] = a_f; g[2 * H + j] = a_g; g[3 * H + j] = a_o;
    if (dAb != nullptr) {
      __nv_bfloat16* gb = dAb + b * 4L * H;
      gb[j] = __float2bfloat16(a_i); gb[H + j] = __float2bfloat16(a_f);
      gb[2 * H + j] = __float2bfloat16(a_g); gb[3 * H + j] = __float2bfloat16(a_o);
    }
    dc[idx] = dct * f_;
  }
}
int lstm_cell_bwd(float* gates, const float* c, const float* c_prev, const float* dh_ext, const float* dh_rec,
                  float* dc, __nv_bfloat16* dAb, int Bn, int H, cudaStream_t st) {
  k_lstm_cell_bwd<<<grid_for((long)Bn * H, 256), 256, 0, st>>>(gates, c, c_prev, dh_ext, dh_rec, dc, dAb, Bn, H);
  ARCVAE_LAUNCHED();
  return 0;
}

// ---- fused per-step encoder path (hidden sizes without a cluster kernel): bf16 tapes in natural column order ----------
// reverse of one step from the bf16 gate tape: dA (bf16, [Bn,4H]) out; dc carries dL/dc_t in and dL/dc_{t-1} out
__global__ void k_lstm_cell_bwd_b(const __nv_bfloat16* __restrict__ gates, const float* __restrict__ c,
                                  const float* __restrict__ c_prev, const float* __restrict__ dh_ext,
                                  const float* __restrict__ dh_rec, int nsplit, long split_stride,
                                  float* __restrict__ dc, __nv_bfloat16* __restrict__ dAb, int Bn, int H) {
  const long total = (long)Bn * (H >> 1);               // two adjacent hidden units per thread (bf16x2 accesses)
  const int H2 = H >> 1;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long b = idx / H2;
    const int j = (int)(idx - b * H2) * 2;
    const __nv_bfloat16* g = gates + b * 4L * H + j;
    const float2 i_ = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(g));
    const float2 f_ = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(g + H));
    const float2 g_ = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(g + 2L * H));
    const float2 o_ = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(g + 3L * H));
    const long e = b * H + j;
    float2 dh = make_float2(0.f, 0.f);
    if (dh_ext != nullptr) { const float2 t = *reinterpret_cast<const float2*>(dh_ext + e); dh.x += t.x; dh.y += t.y; }
    if (dh_rec != nullptr) {
      for (int s = 0; s < nsplit; s++) {                    // split-K partials of d h_t = dA_{t+1} @ Wh
        const float2 t = *reinterpret_cast<const float2*>(dh_rec + s * split_stride + e);
        dh.x += t.x; dh.y += t.y;
      }
    }
    const float2 cc = *reinterpret_cast<const float2*>(c + e);
    const float2 cp = c_prev != nullptr ? *reinterpret_cast<const float2*>(c_prev + e) : make_float2(0.f, 0.f);
    const float2 dcin = *reinterpret_cast<const float2*>(dc + e);
    float ai[2], af[2], ag[2], ao[2], dco[2];
    const float iv[2] = {i_.x, i_.y}, fv[2] = {f_.x, f_.y}, gv[2] = {g_.x, g_.y}, ov[2] = {o_.x, o_.y};
    const float dhv[2] = {dh.x, dh.y}, cv[2] = {cc.x, cc.y}, cpv[2] = {cp.x, cp.y}, dcv[2] = {dcin.x, dcin.y};
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float tcv = tanh_approx_(cv[k]);
      const float dct = dcv[k] + dhv[k] * ov[k] * (1.f - tcv * tcv);
      ao[k] = dhv[k] * tcv * ov[k] * (1.f - ov[k]);
      ai[k] = dct * gv[k] * iv[k] * (1.f - iv[k]);
      ag[k] = dct * iv[k] * (1.f - gv[k] * gv[k]);
      af[k] = dct * cpv[k] * fv[k] * (1.f - fv[k]);
      dco[k] = dct * fv[k];
    }
    __nv_bfloat16* d = dAb + b * 4L * H + j;
    *reinterpret_cast<__nv_bfloat162*>(d) = __floats2bfloat162_rn(ai[0], ai[1]);
    *reinterpret_cast<__nv_bfloat162*>(d + H) = __floats2bfloat162_rn(af[0], af[1]);
    *reinterpret_cast<__nv_bfloat162*>(d + 2L * H) = __floats2bfloat162_rn(ag[0], ag[1]);
    *reinterpret_cast<__nv_bfloat162*>(d + 3L * H) = __floats2bfloat162_rn(ao[0], ao[1]);
    *reinterpret_cast<float2*>(dc + e) = make_float2(dco[0], dco[1]);
  }
}
int lstm_cell_bwd_b(const __nv_bfloat16* gates_b, const float* c, const float* c_prev, const float* dh_ext,
                    const float* dh_rec, int nsplit, long split_stride, float* dc, __nv_bfloat16* dAb, int Bn, int H,
                    cudaStream_t st) {
  ARCVAE_REQUIRE((H & 1) == 0, "even hidden size");
  k_lstm_cell_bwd_b<<<grid_for((long)Bn * (H >> 1), 256), 256, 0, st>>>(gates_b, c, c_prev, dh_ext, dh_rec, nsplit, split_stride,
                                                                         dc, dAb, Bn, H);
  ARCVAE_LAUNCHED();
  return 0;
}
// tile-permuted bf16 copy of a [4H, D] gate matrix: out row j*256 + gi*64 + u <- W row gi*H + j*64 + u
__global__ void k_perm4_rows_to_bf16(const float* __restrict__ W, int H, int D, __nv_bfloat16* __restrict__ out) {
  const long total = 4L * H * D;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long prow = i / D;
    const int col = (int)(i - prow * D);
    const int j = (int)(prow >> 8), rem = (int)(prow & 255), gi = rem >> 6, u = rem & 63;
    out[i] = __float2bfloat16(W[((long)gi * H + j * 64 + u) * D + col]);
  }
}
int perm4_rows_to_bf16(const float* W, int H, int D, __nv_bfloat16* out, cudaStream_t st) {
  ARCVAE_REQUIRE(H % 64 == 0, "tile-permuted gates need hidden_dim % 64 == 0");
  k_perm4_rows_to_bf16<<<grid_for(4L * H * D, 256), 256, 0, st>>>(W, H, D, out);
  ARCVAE_LAUNCHED();
  return 0;
}
// out[r, :] = table[tok[r], :] for bf16 rows of N elements (N % 8 == 0): 16-byte vectors
__global__ void k_gather_rows_bf16(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ tok, long R, int N8,
                                   __nv_bfloat16* __restrict__ out) {
  const long total = R * N8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / N8;
    const int c = (int)(i - r * N8);
    reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(table) + (long)tok[r] * N8 + c);
  }
}
int gather_rows_bf16(const __nv_bfloat16* table, const int32_t* tok, long R, int N, __nv_bfloat16* out, cudaStream_t st) {
  ARCVAE_REQUIRE((N & 7) == 0, "bf16 row gather needs N % 8 == 0");
  TimeScope ts(TIME_POINTWISE, st);
  k_gather_rows_bf16<<<grid_for(R * (N >> 3), 256, 16), 256, 0, st>>>(table, tok, R, N >> 3, out);
  ARCVAE_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// decoder cells (compact gates i,g,o)
__device__ __forceinline__ void dec0_preact(const float* __restrict__ table, const float* __restrict__ wc,
                                            const float* __restrict__ cond_row, int tokv, int C, int H, int j,
                                            float& ai, float& ag, float& ao) {
  const float* trow = table + (long)tokv * 3 * H;
  ai = trow[j]; ag = trow[H + j]; ao = trow[2 * H + j];
  for (int c = 0; c < C; c++) {
    float cv = cond_row[c];
    ai = fmaf(cv, wc[(long)j * C + c], ai);
    ag = fmaf(cv, wc[(long)(H + j) * C + c], ag);
    ao = fmaf(cv, wc[(long)(2 * H + j) * C + c], ao);
  }
}

__global__ void k_dec_cell0_fwd(const float* __restrict__ table, const float* __restrict__ wc,
                                const int32_t* __restrict__ tok, const float* __restrict__ cond, int B, int C, int H,
                                int R, RowMap rm, float* __restrict__ h, __nv_bfloat16* __restrict__ hb) {
  long total = (long)R * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int i = (int)(idx / H);
    int j = (int)(idx - (long)i * H);
    long r = rm(i);
    int b = (int)(r % B);
    float ai, ag, ao;
    dec0_preact(table, wc, cond + (long)b * C, tok[r], C, H, j, ai, ag, ao);
    float cc = sigmoidf_(ai) * tanhf_(ag);
    float hv = sigmoidf_(ao) * tanhf_(cc);
    if (h != nullptr) h[r * H + j] = hv;
    if (hb != nullptr) hb[r * H + j] = __float2bfloat16(hv);
  }
}
// vectorised form: one thread = 8 consecutive hidden units of one row (H % 8 == 0).  FAST = MUFU.TANH activations
// (bf16 path: h is rounded to bf16 anyway); otherwise full-precision expf / tanhf.
template <bool FAST>
__global__ void k_dec_cell0_fwd_v8(const float* __restrict__ table, const float* __restrict__ wc,
                                   const int32_t* __restrict__ tok, const float* __restrict__ cond, int B, int C, int H,
                                   int R, RowMap rm, float* __restrict__ h, __nv_bfloat16* __restrict__ hb,
                                   __nv_bfloat16* __restrict__ gates_b) {
  const int cpr = H >> 3;                               // 8-unit chunks per row
  const long total = (long)R * cpr;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / cpr);
    const int j = (int)(idx - (long)i * cpr) << 3;
    const long r = rm(i);
    const float* trow = table + (long)__ldg(tok + r) * 3 * H + j;
    float a[3][8];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(trow + g * H));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(trow + g * H + 4));
      a[g][0] = v0.x; a[g][1] = v0.y; a[g][2] = v0.z; a[g][3] = v0.w;
      a[g][4] = v1.x; a[g][5] = v1.y; a[g][6] = v1.z; a[g][7] = v1.w;
    }
    const float* crow = cond + (r % B) * C;
    for (int c = 0; c < C; c++) {
      const float cv = __ldg(crow + c);
#pragma unroll
      for (int g = 0; g < 3; g++)
#pragma unroll
        for (int k = 0; k < 8; k++) a[g][k] = fmaf(cv, __ldg(wc + (long)(g * H + j + k) * C + c), a[g][k]);
    }
    float hv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (FAST) {
        a[0][k] = sigmoid_approx_(a[0][k]); a[1][k] = tanh_approx_(a[1][k]); a[2][k] = sigmoid_approx_(a[2][k]);
        hv[k] = a[2][k] * tanh_approx_(a[0][k] * a[1][k]);
      } else {
        hv[k] = sigmoidf_(a[2][k]) * tanhf_(sigmoidf_(a[0][k]) * tanhf_(a[1][k]));
      }
    }
    if (FAST && gates_b != nullptr) {
      // activated gates for the fused backward, tile-permuted layout [row][64-unit block][i|g|o][64] (needs H % 64 == 0)
      __nv_bfloat16* gb = gates_b + r * 3L * H + (j >> 6) * 192 + (j & 63);
#pragma unroll
      for (int g = 0; g < 3; g++) {
        uint4 o;
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; k++) po[k] = __floats2bfloat162_rn(a[g][2 * k], a[g][2 * k + 1]);
        *reinterpret_cast<uint4*>(gb + g * 64) = o;
      }
    }
    if (h != nullptr) {
      *reinterpret_cast<float4*>(h + r * H + j) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(h + r * H + j + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
    }
    if (hb != nullptr) {
      uint4 o;
      __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int k = 0; k < 4; k++) po[k] = __floats2bfloat162_rn(hv[2 * k], hv[2 * k + 1]);
      *reinterpret_cast<uint4*>(hb + r * H + j) = o;
    }
  }
}
// bf16 path, large row counts: the [V,3H] table is staged ONCE per CTA in shared memory as bf16 (V=80, H=256: 120 KB), so
// the gather never leaves the SM (the global-memory version pulls 1.5-3 KB per row through L2: 1.6 GB at B*T = 524288)
// and the kernel is bound by its h / gate-tape writes.  One persistent CTA per SM, 768 threads, thread = 8 hidden units.
__global__ void __launch_bounds__(768, 1)
k_dec_cell0_fwd_smem(const float* __restrict__ table, const float* __restrict__ wc, const int32_t* __restrict__ tok,
                     const float* __restrict__ cond, int B, int C, int H, int V, int R, RowMap rm,
                     __nv_bfloat16* __restrict__ hb, __nv_bfloat16* __restrict__ gates_b) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(sm_raw);                       // [V,3H]
  float* wcs = reinterpret_cast<float*>(sm_raw + (size_t)V * 3 * H * sizeof(__nv_bfloat16));   // [3H,C]
  const int H3 = 3 * H;
  for (int i = threadIdx.x; i < V * H3 / 2; i += blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(table)[i];
    reinterpret_cast<__nv_bfloat162*>(tb)[i] = __floats2bfloat162_rn(v.x, v.y);
  }
  for (int i = threadIdx.x; i < H3 * C; i += blockDim.x) wcs[i] = wc[i];
  __syncthreads();
  const int cpr = H >> 3;
  const long total = (long)R * cpr;
  const long stride = (long)gridDim.x * blockDim.x;
  // the token (and through it the table row) is a dependent load: fetch the NEXT item's token and condition while this
  // item's cell math runs, otherwise every item pays a full L2 round trip with nothing to overlap it
  long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  // the grid stride is a multiple of the chunks per row, so a thread always owns the SAME 8 hidden units: its 24 cond
  // weights live in registers (C == 1; reading them from shared memory per item was an 8-way bank conflict)
  const bool wreg_ok = (C == 1) && (stride % cpr) == 0;
  float wreg[3][8];
  if (wreg_ok && idx < total) {
    const int j0 = (int)(idx % cpr) << 3;
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
      for (int k = 0; k < 8; k++) wreg[g][k] = wcs[g * H + j0 + k];
  }
  long r_n = 0;
  int tok_n = 0, j_n = 0;
  float c0_n = 0.f;
  // 32-bit index arithmetic (R * H / 8 < 2^31, checked by the launcher): the 64-bit divisions and the 64-bit `r % B` of the
  // first version were a large part of the instruction stream of a kernel that is not memory-bound
  auto split = [&](long id, int& i, int& j) {
    const unsigned u = (unsigned)id;
    i = (int)(u / (unsigned)cpr);
    j = (int)(u - (unsigned)i * (unsigned)cpr) << 3;
  };
  if (idx < total) {
    int i;
    split(idx, i, j_n);
    r_n = rm(i);
    tok_n = __ldg(tok + r_n);
    c0_n = __ldg(cond + (long)((unsigned)r_n % (unsigned)B) * C);
  }
  for (; idx < total; idx += stride) {
    const long r = r_n;
    const int j = j_n;
    const int tokv = tok_n;
    const float cond0 = c0_n;
    if (idx + stride < total) {
      int i;
      split(idx + stride, i, j_n);
      r_n = rm(i);
      tok_n = __ldg(tok + r_n);
      c0_n = __ldg(cond + (long)((unsigned)r_n % (unsigned)B) * C);
    }
    const __nv_bfloat16* trow = tb + (long)tokv * H3 + j;
    float a[3][8];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const uint4 v = *reinterpret_cast<const uint4*>(trow + g * H);
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = __bfloat1622float2(pv[k]);
        a[g][2 * k] = t.x; a[g][2 * k + 1] = t.y;
      }
    }
    if (wreg_ok) {
#pragma unroll
      for (int g = 0; g < 3; g++)
#pragma unroll
        for (int k = 0; k < 8; k++) a[g][k] = fmaf(cond0, wreg[g][k], a[g][k]);
    } else {
      const float* crow = cond + (long)((unsigned)r % (unsigned)B) * C;
      for (int c = 0; c < C; c++) {
        const float cv = (c == 0) ? cond0 : __ldg(crow + c);
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
          for (int k = 0; k < 8; k++) a[g][k] = fmaf(cv, wcs[(g * H + j + k) * C + c], a[g][k]);
      }
    }
    float hv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      a[0][k] = sigmoid_approx_(a[0][k]); a[1][k] = tanh_approx_(a[1][k]); a[2][k] = sigmoid_approx_(a[2][k]);
      hv[k] = a[2][k] * tanh_approx_(a[0][k] * a[1][k]);
    }
    if (gates_b != nullptr) {
      __nv_bfloat16* gb = gates_b + r * 3L * H + (j >> 6) * 192 + (j & 63);
#pragma unroll
      for (int g = 0; g < 3; g++) {
        uint4 o;
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; k++) po[k] = __floats2bfloat162_rn(a[g][2 * k], a[g][2 * k + 1]);
        *reinterpret_cast<uint4*>(gb + g * 64) = o;
      }
    }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; k++) po[k] = __floats2bfloat162_rn(hv[2 * k], hv[2 * k + 1]);
    *reinterpret_cast<uint4*>(hb + r * H + j) = o;
  }
}

// Reverse of the layer-0 decoder cell WITHOUT a gate tape: the gates are a function of (token, cond) only, so they are
// recomputed from the same shared-memory table the forward kernel uses (4 MUFU per unit) instead of being written
// (0.8 GB) by the forward pass and read back (0.8 GB) by the d h GEMM's epilogue.  In: dhb [R,H] bf16 (d h_0 from the plain
// GEMM dG_1 @ Wx_1); out: dG_0 [R,3H] bf16 in the tile-permuted compact layout the table scatter expects.
__global__ void __launch_bounds__(768, 1)
k_dec_cell0_bwd_smem(const float* __restrict__ table, const float* __restrict__ wc, const int32_t* __restrict__ tok,
                     const float* __restrict__ cond, int B, int C, int H, int V, long R,
                     const __nv_bfloat16* __restrict__ dhb, __nv_bfloat16* __restrict__ dg_out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(sm_raw);                       // [V,3H]
  float* wcs = reinterpret_cast<float*>(sm_raw + (size_t)V * 3 * H * sizeof(__nv_bfloat16));   // [3H,C]
  const int H3 = 3 * H;
  for (int i = threadIdx.x; i < V * H3 / 2; i += blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(table)[i];
    reinterpret_cast<__nv_bfloat162*>(tb)[i] = __floats2bfloat162_rn(v.x, v.y);
  }
  for (int i = threadIdx.x; i < H3 * C; i += blockDim.x) wcs[i] = wc[i];
  __syncthreads();
  const int cpr = H >> 3;
  const long total = R * cpr;
  const long stride = (long)gridDim.x * blockDim.x;
  long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  // the grid stride is a multiple of the chunks per row, so a thread always owns the SAME 8 hidden units: its 24 cond
  // weights live in registers (C == 1; from shared memory they are an 8-way bank conflict per item)
  const bool wreg_ok = (C == 1) && (stride % cpr) == 0;
  float wreg[3][8];
  if (wreg_ok && idx < total) {
    const int j0 = (int)(idx % cpr) << 3;
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
      for (int k = 0; k < 8; k++) wreg[g][k] = wcs[g * H + j0 + k];
  }
  long r_n = 0;
  int tok_n = 0, j_n = 0;
  float c0_n = 0.f;
  uint4 dh_n = make_uint4(0, 0, 0, 0);
  auto split = [&](long id, long& r, int& j) {        // 32-bit index arithmetic (R * H / 8 < 2^31, checked by the launcher)
    const unsigned u = (unsigned)id;
    const unsigned i = u / (unsigned)cpr;
    r = (long)i;
    j = (int)(u - i * (unsigned)cpr) << 3;
  };
  if (idx < total) {
    split(idx, r_n, j_n);
    tok_n = __ldg(tok + r_n);
    c0_n = __ldg(cond + (long)((unsigned)r_n % (unsigned)B) * C);
    dh_n = __ldcs(reinterpret_cast<const uint4*>(dhb + r_n * H + j_n));
  }
  for (; idx < total; idx += stride) {
    const long r = r_n;
    const int j = j_n, tokv = tok_n;
    const uint4 dhv = dh_n;
    const float cond0 = c0_n;
    if (idx + stride < total) {                  // next item's dependent loads under this item's math
      split(idx + stride, r_n, j_n);
      tok_n = __ldg(tok + r_n);
      c0_n = __ldg(cond + (long)((unsigned)r_n % (unsigned)B) * C);
      dh_n = __ldcs(reinterpret_cast<const uint4*>(dhb + r_n * H + j_n));
    }
    const __nv_bfloat16* trow = tb + (long)tokv * H3 + j;
    float a[3][8];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const uint4 v = *reinterpret_cast<const uint4*>(trow + g * H);
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = __bfloat1622float2(pv[k]);
        a[g][2 * k] = t.x; a[g][2 * k + 1] = t.y;
      }
    }
    if (wreg_ok) {
#pragma unroll
      for (int g = 0; g < 3; g++)
#pragma unroll
        for (int k = 0; k < 8; k++) a[g][k] = fmaf(cond0, wreg[g][k], a[g][k]);
    } else {
      const float* crow = cond + (long)((unsigned)r % (unsigned)B) * C;
      for (int c = 0; c < C; c++) {
        const float cv = (c == 0) ? cond0 : __ldg(crow + c);
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
          for (int k = 0; k < 8; k++) a[g][k] = fmaf(cv, wcs[(g * H + j + k) * C + c], a[g][k]);
      }
    }
    float dh[8];
    {
      const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&dhv);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = __bfloat1622float2(pd[k]);
        dh[2 * k] = t.x; dh[2 * k + 1] = t.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float i_ = sigmoid_approx_(a[0][k]), g_ = tanh_approx_(a[1][k]), o_ = sigmoid_approx_(a[2][k]);
      const float tcv = tanh_approx_(i_ * g_);
      const float dcc = dh[k] * o_ * (1.f - tcv * tcv);
      a[2][k] = dh[k] * tcv * o_ * (1.f - o_);
      a[0][k] = dcc * g_ * i_ * (1.f - i_);
      a[1][k] = dcc * i_ * (1.f - g_ * g_);
    }
    __nv_bfloat16* gb = dg_out + r * 3L * H + (j >> 6) * 192 + (j & 63);
#pragma unroll
    for (int g = 0; g < 3; g++) {
      uint4 o;
      __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int k = 0; k < 4; k++) po[k] = __floats2bfloat162_rn(a[g][2 * k], a[g][2 * k + 1]);
      *reinterpret_cast<uint4*>(gb + g * 64) = o;
    }
  }
}
bool dec_cell0_recompute_ok(int H, int V, int C) {
  return (H % 64) == 0 && V > 0 && (size_t)V * 3 * H * sizeof(__nv_bfloat16) + (size_t)3 * H * C * sizeof(float) <= 200 * 1024 &&
         std::getenv("ARCVAE_NO_CELL0_RECOMPUTE") == nullptr;
}
int dec_cell0_bwd_recompute(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H, int V,
                            long R, const __nv_bfloat16* dhb, __nv_bfloat16* dg_out, cudaStream_t st) {
  if (R <= 0) return 0;
  ARCVAE_REQUIRE(dec_cell0_recompute_ok(H, V, C), "layer-0 cell recompute: table must fit shared memory");
  ARCVAE_REQUIRE(R * (H >> 3) < (1L << 31) && R < (1L << 31), "layer-0 cell recompute: 32-bit item index");
  TimeScope ts(TIME_POINTWISE, st);
  const size_t smem_tab = (size_t)V * 3 * H * sizeof(__nv_bfloat16) + (size_t)3 * H * C * sizeof(float);
  if (first_use_on_device(ONCE_CELL0B))
    ARCVAE_CUDA(cudaFuncSetAttribute(k_dec_cell0_bwd_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const long items = R * (H >> 3);
  int grid = (int)((items + 767) / 768);
  if (grid > 148) grid = 148;
  k_dec_cell0_bwd_smem<<<grid, 768, smem_tab, st>>>(table, wc, tok, cond, B, C, H, V, R, dhb, dg_out);
  ARCVAE_LAUNCHED();
  return 0;
}

int dec_cell0_fwd(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H, int V,
                  int R, RowMap rm, float* h, __nv_bfloat16* hb, __nv_bfloat16* gates_b, cudaStream_t st) {
  if (R <= 0) return 0;
  TimeScope ts(TIME_POINTWISE, st);
  ARCVAE_REQUIRE(gates_b == nullptr || (h == nullptr && (H % 64) == 0), "layer-0 gate tape: fused bf16 path, H % 64 == 0");
  const size_t smem_tab = (size_t)V * 3 * H * sizeof(__nv_bfloat16) + (size_t)3 * H * C * sizeof(float);
  if (h == nullptr && hb != nullptr && (H & 7) == 0 && V > 0 && smem_tab <= 200 * 1024 && (long)R * H >= (1L << 24) &&
      (long)R * (H >> 3) < (1L << 31)) {
    if (first_use_on_device(ONCE_CELL0))
      ARCVAE_CUDA(cudaFuncSetAttribute(k_dec_cell0_fwd_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_dec_cell0_fwd_smem<<<148, 768, smem_tab, st>>>(table, wc, tok, cond, B, C, H, V, R, rm, hb, gates_b);
    ARCVAE_LAUNCHED();
    return 0;
  }
  if ((H & 7) == 0) {
    const long total = (long)R * (H >> 3);
    if (h == nullptr) k_dec_cell0_fwd_v8<true><<<grid_for(total, 256, 16), 256, 0, st>>>(table, wc, tok, cond, B, C, H, R, rm, h, hb, gates_b);
    else k_dec_cell0_fwd_v8<false><<<grid_for(total, 256, 16), 256, 0, st>>>(table, wc, tok, cond, B, C, H, R, rm, h, hb, nullptr);
  } else {
    k_dec_cell0_fwd<<<grid_for((long)R * H, 256), 256, 0, st>>>(table, wc, tok, cond, B, C, H, R, rm, h, hb);
  }
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_dec_cell_fwd(float* __restrict__ G, float* __restrict__ h, __nv_bfloat16* __restrict__ hb, int H,
                               int R, RowMap rm) {
  long total = (long)R * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int i = (int)(idx / H);
    int j = (int)(idx - (long)i * H);
    long r = rm(i);
    float* g = G + r * 3L * H;
    float i_ = sigmoidf_(g[j]), g_ = tanhf_(g[H + j]), o_ = sigmoidf_(g[2 * H + j]);
    g[j] = i_; g[H + j] = g_; g[2 * H + j] = o_;
    float hv = o_ * tanhf_(i_ * g_);
    h[r * H + j] = hv;
    if (hb != nullptr) hb[r * H + j] = __float2bfloat16(hv);
  }
}
int dec_cell_fwd(float* G, float* h, __nv_bfloat16* hb, int H, int R, RowMap rm, cudaStream_t st) {
  if (R <= 0) return 0;
  k_dec_cell_fwd<<<grid_for((long)R * H, 256), 256, 0, st>>>(G, h, hb, H, R, rm);
  ARCVAE_LAUNCHED();
  return 0;
}

__device__ __forceinline__ void dec_cell_grads(float i_, float g_, float o_, float dh, float& dai, float& dag,
                                               float& dao) {
  float cc = i_ * g_;
  float tc = tanhf_(cc);
  float dcc = dh * o_ * (1.f - tc * tc);
  dao = dh * tc * o_ * (1.f - o_);
  dai = dcc * g_ * i_ * (1.f - i_);
  dag = dcc * i_ * (1.f - g_ * g_);
}

__global__ void k_dec_cell_bwd(float* __restrict__ G, const float* __restrict__ dh, __nv_bfloat16* __restrict__ dGb,
                               int H, long R) {
  long total = R * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long r = idx / H;
    int j = (int)(idx - r * H);
    float* g = G + r * 3L * H;
    float dai, dag, dao;
    dec_cell_grads(g[j], g[H + j], g[2 * H + j], dh[idx], dai, dag, dao);
    g[j] = dai; g[H + j] = dag; g[2 * H + j] = dao;
    if (dGb != nullptr) {
      __nv_bfloat16* gb = dGb + r * 3L * H;
      gb[j] = __float2bfloat16(dai); gb[H + j] = __float2bfloat16(dag); gb[2 * H + j] = __float2bfloat16(dao);
    }
  }
}
int dec_cell_bwd(float* G, const float* dh, __nv_bfloat16* dGb, int H, long R, cudaStream_t st) {
  k_dec_cell_bwd<<<grid_for(R * H, 256), 256, 0, st>>>(G, dh, dGb, H, R);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_dec_cell0_bwd(const float* __restrict__ table, const float* __restrict__ wc,
                                const int32_t* __restrict__ tok, const float* __restrict__ cond, int B, int C, int H,
                                long R, const float* __restrict__ dh, float* __restrict__ dG) {
  long total = R * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long r = idx / H;
    int j = (int)(idx - r * H);
    int b = (int)(r % B);
    float ai, ag, ao;
    dec0_preact(table, wc, cond + (long)b * C, tok[r], C, H, j, ai, ag, ao);
    float dai, dag, dao;
    dec_cell_grads(sigmoidf_(ai), tanhf_(ag), sigmoidf_(ao), dh[idx], dai, dag, dao);
    float* g = dG + r * 3L * H;
    g[j] = dai; g[H + j] = dag; g[2 * H + j] = dao;
  }
}
int dec_cell0_bwd(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H,
                  long R, const float* dh, float* dG, cudaStream_t st) {
  k_dec_cell0_bwd<<<grid_for(R * H, 256), 256, 0, st>>>(table, wc, tok, cond, B, C, H, R, dh, dG);
  ARCVAE_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_head_build_u(const float* __restrict__ h_last, const float* __restrict__ cond,
                               const float* __restrict__ Wc, const float* __restrict__ bc, int B, int H, int C,
                               float* __restrict__ u, __nv_bfloat16* __restrict__ ub) {
  long total = (long)B * 2 * H;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    long b = idx / (2 * H);
    int j = (int)(idx - b * 2 * H);
    float v;
    if (j < H) {
      v = h_last[b * H + j];
    } else {
      int jj = j - H;
      v = bc[jj];
      for (int c = 0; c < C; c++) v = fmaf(cond[b * C + c], Wc[(long)jj * C + c], v);
    }
    u[idx] = v;
    if (ub != nullptr) ub[idx] = __float2bfloat16(v);
  }
}
int head_build_u(const float* h_last, const float* cond, const float* Wc, const float* bc, int B, int H, int C,
                 float* u, __nv_bfloat16* ub, cudaStream_t st) {
  k_head_build_u<<<grid_for((long)B * 2 * H, 256), 256, 0, st>>>(h_last, cond, Wc, bc, B, H, C, u, ub);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_tanh_inplace(float* x, long n, __nv_bfloat16* __restrict__ xb) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float y = tanhf_(x[i]);
    x[i] = y;
    if (xb != nullptr) xb[i] = __float2bfloat16(y);
  }
}
int tanh_inplace(float* x, long n, __nv_bfloat16* xb, cudaStream_t st) {
  k_tanh_inplace<<<grid_for(n, 256), 256, 0, st>>>(x, n, xb);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_tanh_bwd_inplace(float* d, const float* __restrict__ y, long n, __nv_bfloat16* __restrict__ db) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float yy = y[i];
    const float v = d[i] * (1.f - yy * yy);
    d[i] = v;
    if (db != nullptr) db[i] = __float2bfloat16(v);
  }
}
int tanh_bwd_inplace(float* d, const float* y, long n, __nv_bfloat16* db, cudaStream_t st) {
  k_tanh_bwd_inplace<<<grid_for(n, 256), 256, 0, st>>>(d, y, n, db);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_head_bound(const float* __restrict__ mu_raw, const float* __restrict__ lv_raw, long n,
                             float* __restrict__ mu, float* __restrict__ logvar) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    mu[i] = tanhf_(mu_raw[i] * 0.5f) * 2.0f;
    logvar[i] = tanhf_(lv_raw[i] * 0.5f) - 1.0f;
  }
}
int head_bound(const float* mu_raw, const float* lv_raw, long n, float* mu, float* logvar, cudaStream_t st) {
  k_head_bound<<<grid_for(n, 256), 256, 0, st>>>(mu_raw, lv_raw, n, mu, logvar);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_head_bound_bwd(const float* __restrict__ mu, const float* __restrict__ logvar,
                                 const float* __restrict__ dmu, const float* __restrict__ dlogvar, long n,
                                 float* __restrict__ dmu_raw, float* __restrict__ dlv_raw,
                                 __nv_bfloat16* __restrict__ dmu_b, __nv_bfloat16* __restrict__ dlv_b) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float tm = mu[i] * 0.5f;          // tanh(mu_raw/2)
    float tl = logvar[i] + 1.0f;      // tanh(lv_raw/2)
    const float a = dmu[i] * (1.f - tm * tm);            // d/dx 2*tanh(x/2) = 1 - tanh^2
    const float b = dlogvar[i] * 0.5f * (1.f - tl * tl);  // d/dx tanh(x/2) = (1 - tanh^2)/2
    dmu_raw[i] = a;
    dlv_raw[i] = b;
    if (dmu_b != nullptr) { dmu_b[i] = __float2bfloat16(a); dlv_b[i] = __float2bfloat16(b); }
  }
}
int head_bound_bwd(const float* mu, const float* logvar, const float* dmu, const float* dlogvar, long n,
                   float* dmu_raw, float* dlv_raw, __nv_bfloat16* dmu_b, __nv_bfloat16* dlv_b, cudaStream_t st) {
  k_head_bound_bwd<<<grid_for(n, 256), 256, 0, st>>>(mu, logvar, dmu, dlogvar, n, dmu_raw, dlv_raw, dmu_b, dlv_b);
  ARCVAE_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// column sums: block = 32 x 8 (32 columns, 8 row lanes), rows strided over grid.y
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void k_colsum(const TIn* __restrict__ X, long R, int N, int ldx, float* __restrict__ out,
                         long rows_per_block) {
  __shared__ float red[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  long r0 = blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float acc = 0.f;
  if (n < N)
    for (long r = r0 + threadIdx.y; r < r1; r += 8) acc += ldf(X + r * ldx + n);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}
template <typename TIn>
static int colsum_t(const TIn* X, long R, int N, int ldx, float* out, cudaStream_t st) {
  if (R <= 0 || N <= 0) return 0;
  int gx = cdiv(N, 32);
  long want = (148L * 8 + gx - 1) / gx;
  long rpb = (R + want - 1) / want;
  if (rpb < 64) rpb = 64;
  int gy = cdiv(R, rpb);
  k_colsum<TIn><<<dim3(gx, gy), dim3(32, 8), 0, st>>>(X, R, N, ldx, out, rpb);
  ARCVAE_LAUNCHED();
  return 0;
}
// bf16, 8 columns (16 bytes) per thread: a warp reads 512 contiguous bytes of a row; block = 32 x 8 covers 256 columns
__global__ void k_colsum_bf16_v8(const __nv_bfloat16* __restrict__ X, long R, int N, int ldx, float* __restrict__ out,
                                 long rows_per_block) {
  __shared__ float red[8][32][9];
  const int n = (blockIdx.x * 32 + threadIdx.x) * 8;
  const long r0 = blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * ldx + n));
      const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = __bfloat1622float2(pv[k]);
        acc[2 * k] += t.x; acc[2 * k + 1] += t.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) red[threadIdx.y][threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) s += red[i][threadIdx.x][k];
      atomicAdd(out + n + k, s);
    }
  }
}
int colsum(const float* X, long R, int N, int ldx, float* out, cudaStream_t st) { return colsum_t(X, R, N, ldx, out, st); }
int colsum_bf16(const __nv_bfloat16* X, long R, int N, int ldx, float* out, cudaStream_t st) {
  if (R <= 0 || N <= 0) return 0;
  if ((N % 8) == 0 && (ldx % 8) == 0 && ((reinterpret_cast<uintptr_t>(X) & 15) == 0)) {
    const int gx = cdiv(N, 256);
    long want = (148L * 8 + gx - 1) / gx;
    long rpb = (R + want - 1) / want;
    if (rpb < 64) rpb = 64;
    TimeScope ts(TIME_POINTWISE, st);
    k_colsum_bf16_v8<<<dim3(gx, cdiv(R, rpb)), dim3(32, 8), 0, st>>>(X, R, N, ldx, out, rpb);
    ARCVAE_LAUNCHED();
    return 0;
  }
  return colsum_t(X, R, N, ldx, out, st);
}

// scatter rows by token with an smem accumulator: block = 128 threads (one per column of a 128-wide slab),
// each block walks `rows_per_block` rows; thread j owns column j so the smem adds are conflict-free and
// un-contended.  V*128 floats of smem (V=80: 40 KB).
template <typename TIn>
__global__ void k_scatter_rows_by_token(const TIn* __restrict__ X, const int32_t* __restrict__ tok, long R, int N,
                                        int V, float* __restrict__ dtable, const float* __restrict__ cond, int B, int C,
                                        float* __restrict__ dwc, long rows_per_block) {
  extern __shared__ float acc[];  // [V][128]
  int j = threadIdx.x;
  int n = blockIdx.x * 128 + j;
  for (int v = 0; v < V; v++) acc[v * 128 + j] = 0.f;
  long r0 = blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float wacc[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    for (long r = r0; r < r1; r++) {
      float x = ldf(X + r * N + n);
      acc[tok[r] * 128 + j] += x;
      if (dwc != nullptr) {
        const float* cr = cond + (r % B) * C;
        for (int c = 0; c < C && c < 4; c++) wacc[c] = fmaf(x, cr[c], wacc[c]);
      }
    }
    for (int v = 0; v < V; v++) {
      float s = acc[v * 128 + j];
      if (s != 0.f) atomicAdd(dtable + (long)v * N + n, s);
    }
    if (dwc != nullptr)
      for (int c = 0; c < C && c < 4; c++) atomicAdd(dwc + (long)n * C + c, wacc[c]);
  }
}
template <typename TIn>
static int scatter_t(const TIn* X, const int32_t* tok, long R, int N, int V, float* dtable, const float* cond,
                     int B, int C, float* dwc, cudaStream_t st) {
  if (R <= 0) return 0;
  ARCVAE_REQUIRE(dwc == nullptr || C <= 4, "num_conditions > 4 not supported by the fused cond-weight reduction");
  size_t smem = (size_t)V * 128 * sizeof(float);
  ARCVAE_REQUIRE(smem <= 200 * 1024, "vocab too large for the smem scatter accumulator");
  if (first_use_on_device(sizeof(TIn) == 4 ? ONCE_SCATTER_F32 : ONCE_SCATTER_BF16))
    ARCVAE_CUDA(cudaFuncSetAttribute(k_scatter_rows_by_token<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int gx = cdiv(N, 128);
  long want = (148L * 4 + gx - 1) / gx;
  long rpb = (R + want - 1) / want;
  if (rpb < 256) rpb = 256;
  int gy = cdiv(R, rpb);
  k_scatter_rows_by_token<TIn><<<dim3(gx, gy), 128, smem, st>>>(X, tok, R, N, V, dtable, cond, B, C, dwc, rpb);
  ARCVAE_LAUNCHED();
  return 0;
}
int scatter_rows_by_token(const float* X, const int32_t* tok, long R, int N, int V, float* dtable, const float* cond,
                          int B, int C, float* dwc, cudaStream_t st) {
  return scatter_t(X, tok, R, N, V, dtable, cond, B, C, dwc, st);
}
int scatter_rows_by_token_bf16(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, float* dtable,
                               cudaStream_t st) {
  return scatter_t(X, tok, R, N, V, dtable, nullptr, 1, 1, nullptr, st);
}
int scatter_rows_by_token_bf16_w(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, float* dtable,
                                 const float* cond, int B, int C, float* dwc, cudaStream_t st) {
  return scatter_t(X, tok, R, N, V, dtable, cond, B, C, dwc, st);
}

// ------------------------------------------------------------------------------------------------
// Token scatter on the tensor cores: dtable = onehot(tok)^T @ X is a GEMM with K = R.  The one-hot operand is
// materialised as bf16 [R, NW] (NW = 128): columns [0,V) one-hot of the token (exact in bf16), columns V+2c, V+2c+1 the
// bf16 hi / lo split of cond[r % B, c] (so that dwc = X^T cond comes out of the same GEMM with ~2^-17 relative error),
// the rest zero.  Products with 0/1 are exact and accumulation is fp32, so the result equals the fp32 scatter of the
// bf16 input up to summation order.
// One thread per 16-byte chunk (8 columns) of a row; NW = SCATTER_NW = 128 -> 16 chunks per row, shifts instead of 64-bit
// divisions (the first version spent 65 us on 134 MB: 3x the HBM time).  Only the chunk that holds the token's column and
// the chunks that overlap the cond columns [V, V + 2C) are non-zero.
__global__ void k_build_onehot(const int32_t* __restrict__ tok, long R, int V, const float* __restrict__ cond, int B, int C,
                               __nv_bfloat16* __restrict__ out) {
  constexpr int CPR = SCATTER_NW >> 3;
  const long total = R * CPR;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / CPR;
    const int ch = (int)(idx & (CPR - 1));
    const int c0 = ch << 3;
    const int t = __ldg(tok + r);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if ((t >> 3) == ch) w[(t & 7) >> 1] = (t & 1) ? 0x3F800000u : 0x00003F80u;      // bf16 1.0 in the low / high half
    if (cond != nullptr && c0 + 8 > V && c0 < V + 2 * C) {
      __nv_bfloat16* v = reinterpret_cast<__nv_bfloat16*>(w);
      const int b = (int)(r % B);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int col = c0 + k;
        if (col >= V && col < V + 2 * C) {
          const int c = (col - V) >> 1;
          const float cv = __ldg(cond + (long)b * C + c);
          const float hi = __bfloat162float(__float2bfloat16(cv));
          v[k] = __float2bfloat16(((col - V) & 1) ? (cv - hi) : hi);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + r * SCATTER_NW + c0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// dwc[n*C + c] += D[(V+2c)*N + n] + D[(V+2c+1)*N + n]
__global__ void k_onehot_dwc(const float* __restrict__ D, int N, int V, int C, float* __restrict__ dwc) {
  const int total = N * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / C, c = i - n * C;
    dwc[i] += D[(long)(V + 2 * c) * N + n] + D[(long)(V + 2 * c + 1) * N + n];
  }
}
// dst[r, nat(p)] = src[r, p]: columns in tile-permuted compact order (p = j*192 + gi*64 + u) -> natural (gi*H + j*64 + u)
__global__ void k_unpermute_cols(const float* __restrict__ src, int rows, int H, float* __restrict__ dst) {
  const int H3 = 3 * H;
  const long total = (long)rows * H3;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / H3;
    const int pc = (int)(i - r * H3);
    const int j = pc / 192, rem = pc % 192;
    dst[r * H3 + (rem / 64) * H + j * 64 + (rem % 64)] = src[i];
  }
}
__global__ void k_rowsum_add(const float* __restrict__ X, int M, int ldx, int V, float* __restrict__ out) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += X[(long)m * ldx + v];
  s = warp_sum(s);
  if (lane == 0) out[m] += s;
}
int rowsum_add(const float* X, int M, int ldx, int V, float* out, cudaStream_t st) {
  k_rowsum_add<<<cdiv(M, 8), 256, 0, st>>>(X, M, ldx, V, out);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_transpose_f32(const float* __restrict__ src, int R, int ldx, int Cn, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cn) tile[i][threadIdx.x] = src[(long)r * ldx + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cn) dst[(long)c * R + r] = tile[threadIdx.x][i];
  }
}
int transpose_f32(const float* src, int R, int ldx, int Cn, float* dst, cudaStream_t st) {
  k_transpose_f32<<<dim3(cdiv(Cn, 32), cdiv(R, 32)), dim3(32, 8), 0, st>>>(src, R, ldx, Cn, dst);
  ARCVAE_LAUNCHED();
  return 0;
}
int build_onehot(const int32_t* tok, long R, int V, const float* cond, int B, int C, __nv_bfloat16* onehot, cudaStream_t st) {
  TimeScope ts(TIME_POINTWISE, st);
  k_build_onehot<<<grid_for(R * (SCATTER_NW >> 3), 256, 16), 256, 0, st>>>(tok, R, V, cond, B, C, onehot);
  ARCVAE_LAUNCHED();
  return 0;
}
bool scatter_onehot_supported(int N, int V, int C) { return (N % 64) == 0 && V + 2 * C <= SCATTER_NW; }
// perm_H > 0: X's columns are in tile-permuted compact order (N = 3*perm_H); dtable_ext comes out in natural order and
// `tmp` ([SCATTER_NW, N] floats) receives the permuted product first
int scatter_rows_onehot_tc(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, __nv_bfloat16* onehot,
                           float* dtable_ext, const float* cond, int B, int C, float* dwc, int perm_H, float* tmp,
                           cudaStream_t st) {
  float* natural = dtable_ext;
  if (perm_H > 0) {
    ARCVAE_REQUIRE(tmp != nullptr && N == 3 * perm_H && (perm_H % 64) == 0, "permuted scatter needs a scratch table, N = 3H");
    dtable_ext = tmp;
  }
  if (R <= 0) return 0;
  ARCVAE_REQUIRE(scatter_onehot_supported(N, V, cond != nullptr ? C : 0), "one-hot scatter: N % 64 == 0, V + 2C <= 128");
  const int NW = SCATTER_NW;
  if (tok != nullptr) ARCVAE_TRY(build_onehot(tok, R, V, cond, B, C, onehot, st));   // tok == nullptr: `onehot` is already built
  ARCVAE_CUDA(cudaMemsetAsync(dtable_ext, 0, (size_t)NW * N * sizeof(float), st));
  TcGemm g{};
  g.M = NW; g.N = N; g.K = (int)R;
  g.A = onehot; g.lda = NW; g.a_mn = true;
  g.B = X; g.ldb = N; g.b_mn = true;
  g.C = dtable_ext; g.ldc = N; g.Cb = nullptr; g.ldcb = 0; g.bias = nullptr; g.accumulate = true;
  g.splitk = pick_splitk_tc(NW, N, (int)R);
  g.rm = RowMap{nullptr, 1}; g.a_rows_total = R;
  ARCVAE_TRY(gemm_tc(g, st));
  if (perm_H > 0) {
    k_unpermute_cols<<<grid_for((long)NW * N, 256), 256, 0, st>>>(dtable_ext, NW, perm_H, natural);
    ARCVAE_LAUNCHED();
    dtable_ext = natural;
  }
  if (cond != nullptr && dwc != nullptr) {
    k_onehot_dwc<<<cdiv((long)N * C, 256), 256, 0, st>>>(dtable_ext, N, V, C, dwc);
    ARCVAE_LAUNCHED();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void k_compact_gates(const float* __restrict__ full, int H, int D, float* __restrict__ compact) {
  long total = 3L * H * D;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long row = i / D;
    int c = (int)(i - row * D);
    long src = row < H ? row : row + H;  // i rows [0,H) ; g,o rows [2H,4H)
    compact[i] = full[src * D + c];
  }
}
int compact_gates(const float* full, int H, int D, float* compact, cudaStream_t st) {
  k_compact_gates<<<grid_for(3L * H * D, 256), 256, 0, st>>>(full, H, D, compact);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_expand_gates_add(const float* __restrict__ compact, int H, int D, float* __restrict__ full) {
  long total = 3L * H * D;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long row = i / D;
    int c = (int)(i - row * D);
    long dst = row < H ? row : row + H;
    full[dst * D + c] += compact[i];
  }
}
int expand_gates_add(const float* compact, int H, int D, float* full, cudaStream_t st) {
  k_expand_gates_add<<<grid_for(3L * H * D, 256), 256, 0, st>>>(compact, H, D, full);
  ARCVAE_LAUNCHED();
  return 0;
}
__device__ __forceinline__ long perm_to_full_row(long prow, int H) {
  int j = (int)(prow / 192), rem = (int)(prow % 192);
  int gi = rem / 64, u = rem % 64;
  int gate = gi == 0 ? 0 : (gi == 1 ? 2 : 3);
  return (long)gate * H + j * 64 + u;
}
__global__ void k_compact_perm_gates(const float* __restrict__ full, int H, int D, float* __restrict__ perm) {
  long total = 3L * H * D;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long row = i / D;
    int c = (int)(i - row * D);
    perm[i] = full[perm_to_full_row(row, H) * D + c];
  }
}
int compact_perm_gates(const float* full, int H, int D, float* perm, cudaStream_t st) {
  k_compact_perm_gates<<<grid_for(3L * H * D, 256), 256, 0, st>>>(full, H, D, perm);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_expand_perm_gates_add(const float* __restrict__ perm, int H, int D, float* __restrict__ full) {
  long total = 3L * H * D;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long row = i / D;
    int c = (int)(i - row * D);
    full[perm_to_full_row(row, H) * D + c] += perm[i];
  }
}
int expand_perm_gates_add(const float* perm, int H, int D, float* full, cudaStream_t st) {
  k_expand_perm_gates_add<<<grid_for(3L * H * D, 256), 256, 0, st>>>(perm, H, D, full);
  ARCVAE_LAUNCHED();
  return 0;
}
__global__ void k_add_strided(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int R, int Cn) {
  long total = (long)R * Cn;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / Cn;
    int c = (int)(i - r * Cn);
    dst[r * ldd + c] += src[r * lds + c];
  }
}
int add_strided(const float* src, int lds, float* dst, int ldd, int R, int Cn, cudaStream_t st) {
  if (R <= 0 || Cn <= 0) return 0;
  k_add_strided<<<grid_for((long)R * Cn, 256), 256, 0, st>>>(src, lds, dst, ldd, R, Cn);
  ARCVAE_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one warp per row: argmax with lowest-index tie break (mx.argmax)
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

__global__ void k_argmax_feedback(const float* __restrict__ logits, const int* __restrict__ tlist, int ntl, int B,
                                  int V, int32_t* __restrict__ tok) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  long total = (long)ntl * B;
  for (long w = warp; w < total; w += nwarps) {
    int q = (int)(w / B);
    int b = (int)(w - (long)q * B);
    int t = tlist[q];
    const float* row = logits + ((long)t * B + b) * V;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      float x = row[v];
      if (x > bv) { bv = x; bi = v; }   // strictly greater keeps the lowest index within a lane
    }
    warp_argmax(bv, bi);
    if (lane == 0) tok[(long)(t + 1) * B + b] = bi;
  }
}
int argmax_feedback(const float* logits, const int* tlist, int ntl, int B, int V, int32_t* tok, cudaStream_t st) {
  if (ntl <= 0) return 0;
  long total = (long)ntl * B;
  k_argmax_feedback<<<grid_for(total * 32, 256), 256, 0, st>>>(logits, tlist, ntl, B, V, tok);
  ARCVAE_LAUNCHED();
  return 0;
}

// sampler selection: one warp per row.  multinomial: inverse-CDF over softmax(logits/T) with one Philox uniform.
__global__ void k_select_token(const float* __restrict__ logits, int B, int V, float temperature, int multinomial,
                               uint64_t seed, int step, int max_length, int end_token, int32_t* __restrict__ tokens_out,
                               int32_t* __restrict__ cur, int32_t* __restrict__ ended,
                               int32_t* __restrict__ ended_count) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < B; b += nwarps) {
    const float* row = logits + (long)b * V;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      float x = row[v] / temperature;
      if (x > bv) { bv = x; bi = v; }
    }
    warp_argmax(bv, bi);
    int choice = bi;
    if (multinomial) {
      float s = 0.f;
      for (int v = lane; v < V; v += 32) s += expf(row[v] / temperature - bv);
      s = warp_sum(s);
      uint32_t r4[4];
      philox4x32(seed, (uint64_t)b, (uint64_t)step, r4);
      float target = u01(r4[0]) * s;
      // sequential inverse CDF in index order, 32 entries at a time
      float run = 0.f;
      choice = -1;
      for (int v0 = 0; v0 < V && choice < 0; v0 += 32) {
        int v = v0 + lane;
        float p = (v < V) ? expf(row[v] / temperature - bv) : 0.f;
        float incl = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          float y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        unsigned hit = __ballot_sync(0xffffffffu, (v < V) && (run + incl >= target));
        if (hit) choice = v0 + (__ffs(hit) - 1);
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (choice < 0) choice = bi;  // rounding left the target just above the total mass
    }
    if (lane == 0) {
      tokens_out[(long)b * max_length + step] = choice;
      cur[b] = choice;
      if (choice == end_token && !ended[b]) {
        ended[b] = 1;
        atomicAdd(ended_count, 1);
      }
    }
  }
}
int select_token(const float* logits, int B, int V, float temperature, int multinomial, uint64_t seed, int step,
                 int max_length, int end_token, int32_t* tokens_out, int32_t* cur, int32_t* ended, int32_t* ended_count,
                 cudaStream_t st) {
  k_select_token<<<grid_for((long)B * 32, 256), 256, 0, st>>>(logits, B, V, temperature, multinomial, seed, step,
                                                              max_length, end_token, tokens_out, cur, ended,
                                                              ended_count);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_sampler_check_stop(const int32_t* ended_count, int B, int step, int32_t* t_stop) {
  if (*ended_count >= B && *t_stop > step) *t_stop = step;
}
int sampler_check_stop(const int32_t* ended_count, int B, int step, int32_t* t_stop, cudaStream_t st) {
  k_sampler_check_stop<<<1, 1, 0, st>>>(ended_count, B, step, t_stop);
  ARCVAE_LAUNCHED();
  return 0;
}

__global__ void k_set_int(int32_t* p, int32_t v) { *p = v; }
int set_int(int32_t* p, int32_t v, cudaStream_t st) {
  k_set_int<<<1, 1, 0, st>>>(p, v);
  ARCVAE_LAUNCHED();
  return 0;
}

}  // namespace arcvae
