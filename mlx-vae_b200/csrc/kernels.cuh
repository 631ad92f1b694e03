// kernels.cuh — internal launch wrappers (pointwise.cu, loss.cu, adam.cu) used by the orchestration in
// encoder.cu / decoder.cu / sampler.cu.  All of them enqueue on `st` and return 0 / error code.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace arcvae {

// x[B,T] -> xT[T,B]
int transpose_tokens(const int32_t* x, int B, int T, int32_t* xT, cudaStream_t st);
// out[r,0:N] = table[tok[r],0:N] for r < R
int gather_rows(const float* table, const int32_t* tok, int R, int N, float* out, cudaStream_t st);

// LSTM cell, forward, in place: gates [Bn,4H] holds pre-activations (i,f,g,o) and receives the activated gates
// (sigmoid(i), sigmoid(f), tanh(g), sigmoid(o)); c = f*c_prev + i*g (c_prev == nullptr: c = i*g); h = o*tanh(c)
int lstm_cell_fwd(float* gates, const float* c_prev, float* c, float* h, __nv_bfloat16* hb, int Bn, int H, cudaStream_t st);
// reverse of one step: gates (activated, in) -> dA (pre-activation grads, out, in place).
// dh = dh_ext + dh_rec (either may be null); dc (in/out, [Bn,H]) carries dL/dc_t in and dL/dc_{t-1} out.
int lstm_cell_bwd(float* gates, const float* c, const float* c_prev, const float* dh_ext, const float* dh_rec,
                  float* dc, __nv_bfloat16* dAb, int Bn, int H, cudaStream_t st);

// fused per-step encoder path: bf16 gate tape -> bf16 dA (natural column order), see encoder.cu PATH_STEP_FUSED
// dh_rec: nsplit partial [Bn,H] buffers, split_stride floats apart (split-K GEMM with TcGemm::split_stride)
int lstm_cell_bwd_b(const __nv_bfloat16* gates_b, const float* c, const float* c_prev, const float* dh_ext,
                    const float* dh_rec, int nsplit, long split_stride, float* dc, __nv_bfloat16* dAb, int Bn, int H,
                    cudaStream_t st);
int perm4_rows_to_bf16(const float* W, int H, int D, __nv_bfloat16* out, cudaStream_t st);
int gather_rows_bf16(const __nv_bfloat16* table, const int32_t* tok, long R, int N, __nv_bfloat16* out, cudaStream_t st);

// decoder cells (zero state: c = i*g, h = o*tanh(c); forget gate unused) on COMPACT gate layout [.., 3H] = (i,g,o)
// layer 0: a = table[tok[r]] + cond[r % B] @ wc^T   (table [V,3H], wc [3H,C]); rows r = rm(i), i < R
// gates_b (optional, fused bf16 path): activated (i,g,o) in the tile-permuted layout of the fused GEMM epilogues
// V = rows of the table (sizes the shared-memory copy of the large-batch bf16 kernel)
// layer-0 decoder cell reverse with the gates recomputed from the token table (no gate tape); see pointwise.cu
bool dec_cell0_recompute_ok(int H, int V, int C);
int dec_cell0_bwd_recompute(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H, int V,
                            long R, const __nv_bfloat16* dhb, __nv_bfloat16* dg_out, cudaStream_t st);
int dec_cell0_fwd(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H, int V,
                  int R, RowMap rm, float* h, __nv_bfloat16* hb, __nv_bfloat16* gates_b, cudaStream_t st);
// layers >= 1: G [.,3H] pre-activation -> activated in place; h out
int dec_cell_fwd(float* G, float* h, __nv_bfloat16* hb, int H, int R, RowMap rm, cudaStream_t st);
// backward: G activated (in) -> dG (out, in place) given dh [.,H]
int dec_cell_bwd(float* G, const float* dh, __nv_bfloat16* dGb, int H, long R, cudaStream_t st);
// layer 0 backward with recompute of the gates from the table: dG0 [R,3H] out
int dec_cell0_bwd(const float* table, const float* wc, const int32_t* tok, const float* cond, int B, int C, int H,
                  long R, const float* dh, float* dG, cudaStream_t st);

// head: u[b,0:H] = h_last[b,:]; u[b,H:2H] = cond[b,:] @ Wc^T + bc
// the bf16 pointers (nullable) receive a rounded copy of the fp32 result: operands of the tensor-core head products
int head_build_u(const float* h_last, const float* cond, const float* Wc, const float* bc, int B, int H, int C,
                 float* u, __nv_bfloat16* ub, cudaStream_t st);
int tanh_inplace(float* x, long n, __nv_bfloat16* xb, cudaStream_t st);
// d <- d * (1 - y*y)
int tanh_bwd_inplace(float* d, const float* y, long n, __nv_bfloat16* db, cudaStream_t st);
// mu = 2*tanh(mu_raw/2); logvar = tanh(lv_raw/2) - 1   (encoder.py:126,:130)
int head_bound(const float* mu_raw, const float* lv_raw, long n, float* mu, float* logvar, cudaStream_t st);
// dmu_raw = dmu*(1-(mu/2)^2); dlv_raw = dlogvar*0.5*(1-(logvar+1)^2)
int head_bound_bwd(const float* mu, const float* logvar, const float* dmu, const float* dlogvar, long n,
                   float* dmu_raw, float* dlv_raw, __nv_bfloat16* dmu_b, __nv_bfloat16* dlv_b, cudaStream_t st);

// out[n] += sum_r X[r*ldx + n]
int colsum(const float* X, long R, int N, int ldx, float* out, cudaStream_t st);
// dtable[tok[r], n] += X[r,n];  optional: dwc[n*C + c] += X[r,n] * cond[(r % B)*C + c]
int scatter_rows_by_token(const float* X, const int32_t* tok, long R, int N, int V, float* dtable, const float* cond,
                          int B, int C, float* dwc, cudaStream_t st);
// copy rows [0,H) U [2H,4H) of a [4H,D] matrix (or [4H] vector, D=1) to a compact [3H,D]; and the adjoint (+=)
int compact_gates(const float* full, int H, int D, float* compact, cudaStream_t st);
int expand_gates_add(const float* compact, int H, int D, float* full, cudaStream_t st);
// strided add: dst[r*ldd + c] += src[r*lds + c], r<R, c<Cn
// tile-permuted compact gates: row p = j*192 + gi*64 + u  <-  full row gate(gi)*H + j*64 + u, gate(0,1,2) = (i,g,o) = (0,2,3)
int compact_perm_gates(const float* full, int H, int D, float* perm, cudaStream_t st);
int expand_perm_gates_add(const float* perm, int H, int D, float* full, cudaStream_t st);
int add_strided(const float* src, int lds, float* dst, int ldd, int R, int Cn, cudaStream_t st);

// greedy feedback (decoder.py:185): tok_next[t+1 rows] = argmax(logits[t rows]) for t in tlist (device list, n entries),
// lowest index wins ties.  logits time-major [T,B,V]; tok time-major [T,B]
int argmax_feedback(const float* logits, const int* tlist, int ntl, int B, int V, int32_t* tok, cudaStream_t st);

// sampler token selection for one step (decoder_sampling.py:110-123)
int select_token(const float* logits, int B, int V, float temperature, int multinomial, uint64_t seed, int step,
                 int max_length, int end_token, int32_t* tokens_out, int32_t* cur, int32_t* ended, int32_t* ended_count,
                 cudaStream_t st);
int sampler_check_stop(const int32_t* ended_count, int B, int step, int32_t* t_stop, cudaStream_t st);
int set_int(int32_t* p, int32_t v, cudaStream_t st);

// ---- bf16 tensor-core GEMM (gemm_tc.cu): D[M,N] (+)= A*B (+bias); fp32 accumulate --------------------------------
struct TcGemm {
  int M, N, K;
  const __nv_bfloat16* A; int lda; bool a_mn;   // a_mn=false: A[m*lda + k] (K-major); true: A[k*lda + m] (MN-major)
  const __nv_bfloat16* B; int ldb; bool b_mn;   // b_mn=false: B[n*ldb + k] (K-major); true: B[k*ldb + n] (MN-major)
  float* C; int ldc;                            // fp32 output (nullable)
  __nv_bfloat16* Cb; int ldcb;                  // bf16 output (nullable)
  const float* bias;                            // [N] or null
  bool accumulate;                              // C += ...   (mandatory with splitk > 1: fp32 atomics)
  int splitk;
  long split_stride = 0;                        // > 0 with splitk > 1: split ks STORES its partial to C + ks*split_stride
                                                // (no atomics, no pre-zeroed C; the consumer adds the partials)
  RowMap rm;                                    // rows of A (K-major) and C; rm.Bt must be a multiple of 128
  long a_rows_total;                            // rows of the allocation behind A when rm is used
  // fused decoder epilogues (zero-state LSTM cell on compact, tile-permuted gates; see decoder.cu)
  int epi = 0;                                  // TC_EPI_*
  __nv_bfloat16* gates_b = nullptr;             // [rows,3H] activated gates (written by CELL_FWD, read by CELL_BWD)
  __nv_bfloat16* hb_out = nullptr;              // [rows,H]  CELL_FWD: h
  __nv_bfloat16* dg_out = nullptr;              // [rows,3H] CELL_BWD: pre-activation gradients (same layout as gates_b)
  int Hh = 0;
  // fused encoder LSTM step (TC_EPI_LSTM_FWD): one launch = h_{t-1} @ Wh^T + P_t -> gates -> (c_t, h_t).  B is the
  // TILE-PERMUTED Wh (row j*256 + gi*64 + u <- Wh row gi*H + j*64 + u: a 256-wide N tile = the 4 gates of 64 units);
  // every tensor below keeps the NATURAL column order (i | f | g | o, each H wide)
  const __nv_bfloat16* pre_b = nullptr;         // [rows,4H] input projection of this step incl. bias
  const float* c_prev = nullptr;                // [rows,H] cell state of the previous step (null: zero)
  float* c_out = nullptr;                       // [rows,H]
  float* hf_out = nullptr;                      // [rows,H] fp32 copy of h (nullable; the encoder head reads h_{T-1})
  // fc_out with the cross-entropy in its epilogue (TC_EPI_CE; N = V <= 128, one N tile holds a whole row of logits):
  // CE sum, d logits (bf16, scaled by ce_scale = 1 / global token count) and the greedy feedback token — the fp32 logits
  // never reach HBM (losses/recon.py:29-64 + models/decoder.py:175-185 in one pass)
  const int32_t* ce_target = nullptr;           // [T*B] time-major targets
  const uint8_t* ce_fb = nullptr;               // [T] != 0: position t feeds argmax(logits_t) to position t+1
  int32_t* ce_tok = nullptr;                    // [T*B] decoder inputs (time-major), written at (t+1, b)
  int ce_B = 0;                                 // rows per timestep
  __nv_bfloat16* ce_dl = nullptr;               // [rows, ce_ldl] d logits
  int ce_ldl = 0;
  double* ce_sum = nullptr;                     // += sum of CE over the rows of this launch
  float ce_scale = 0.f;
  // multi-segment B (nseg = 2 or 3): weight gradients that share the MN-major A operand (dA^T) run as ONE GEMM whose
  // column tile i multiplies with seg[i].B (row k of A pairs with row k - k_shift of B; rows before 0 count as zero)
  // and accumulates into seg[i].C — A is read from HBM once.  Requires a_mn, b_mn, accumulate.
  // TC_EPI_LSTM_P (weight-stationary kernel only): Cb receives the layer's input projection in the THREAD-FRIENDLY layout
  // of the cluster recurrence (lstm_cluster.cu: [t][tile][cta][lane quarter][unit half][gate*4 + chunk][lane] x 16 B), so
  // that kernel's epilogue threads read their operands as 512-byte warp accesses.  N = 4*Hh, rows = t*lp_B + b.
  int lp_B = 0;
  int lp_vecs = 16;                             // 16-byte vectors per thread slot of that layout (gate g, chunk c -> vector g*4 + c)
  int nseg = 1;
  struct Seg { const __nv_bfloat16* B; int ldb; int N; int k_shift; float* C; int ldc; } seg[3] = {};
};
enum { TC_EPI_PLAIN = 0, TC_EPI_DEC_CELL_FWD = 1, TC_EPI_DEC_CELL_BWD = 2, TC_EPI_LSTM_FWD = 3, TC_EPI_CE = 4, TC_EPI_LSTM_P = 5, TC_EPI_LSTM_DH = 6 };
int gemm_tc(const TcGemm& g, cudaStream_t st);
// weight-stationary variant (gemm_ws.cu) for K <= 256, bf16 output / fused decoder cell; gemm_tc dispatches to it
bool gemm_ws_supported(const TcGemm& g);
int gemm_ws(const TcGemm& g, cudaStream_t st);
int pick_splitk_tc(int M, int N, int K);
int f32_to_bf16(const float* src, __nv_bfloat16* dst, long n, cudaStream_t st);
int f32_to_bf16_pitched(const float* src, long R, int V, __nv_bfloat16* dst, int Vp, cudaStream_t st);   // pad columns zero
int transpose_to_bf16(const float* src, int R, int C, __nv_bfloat16* dst, cudaStream_t st);   // dst[c*R+r] = src[r*C+c]

// persistent cluster LSTM recurrence (lstm_cluster.cu), H == 256
bool lstm_cluster_supported(int H);
size_t lstm_cluster_xh_bytes(int B);      // forward exchange buffer (flag-in-data vectors of h_t)
size_t lstm_cluster_xch_bytes(int B);     // backward exchange buffer (flag-in-data vectors of the partial d h)
size_t lstm_cluster_ktape_elems(int B, int T, int H);   // bf16 elements of one layer's coefficient tape
// table0b != null: layer 0 (operands gathered from the token table); Pb != null: row-major input projection; both null: the
// projection sits in the tape slots (gemm_ws TC_EPI_LSTM_P, lp_vecs = 24)
int lstm_cluster_forward(int B, int T, int H, const __nv_bfloat16* Whb, const int32_t* xT, const __nv_bfloat16* table0b,
                         const __nv_bfloat16* Pb, __nv_bfloat16* hb, __nv_bfloat16* ktape, float* h_last, void* xh,
                         int* err_flag, cudaStream_t st);
// K-split backward: every CTA multiplies its own dA slice, partial d h reduce-scattered through `xch`.  Gradient from the
// layer above: dh_ext (fp32 [T*B,H], row-major) or dh_tf (bf16, thread-friendly layout written by gemm_tc TC_EPI_LSTM_DH)
size_t lstm_cluster_dh_tf_elems(int B, int T, int H);
int lstm_cluster_backward(int B, int T, int H, const __nv_bfloat16* Whb, const __nv_bfloat16* ktape, const float* dh_ext,
                          const __nv_bfloat16* dh_tf, const float* dh_last, int dh_last_ld, __nv_bfloat16* dAb, void* xch, int* err_flag,
                          cudaStream_t st);
int colsum_bf16(const __nv_bfloat16* X, long R, int N, int ldx, float* out, cudaStream_t st);
// out[m] += sum_{v < V} X[m*ldx + v]   (bias gradient from the one-hot segment of a multi-segment weight-gradient GEMM)
int rowsum_add(const float* X, int M, int ldx, int V, float* out, cudaStream_t st);
// dst[c*R + r] = src[r*ldx + c], r < R, c < Cn   (small fp32 transposes of table gradients)
int transpose_f32(const float* src, int R, int ldx, int Cn, float* dst, cudaStream_t st);
// one-hot operand of the tensor-core token scatter (see scatter_rows_onehot_tc); cond == nullptr: tokens only
int build_onehot(const int32_t* tok, long R, int V, const float* cond, int B, int C, __nv_bfloat16* onehot, cudaStream_t st);
int scatter_rows_by_token_bf16(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, float* dtable,
                               cudaStream_t st);

// the same scatter as a tcgen05 GEMM against a materialised one-hot operand (pointwise.cu): dtable_ext is
// [SCATTER_NW, N] fp32 (rows [0,V) = dtable; rows V+2c, V+2c+1 = hi/lo parts of dwc[:,c]); onehot is bf16 [R, SCATTER_NW]
constexpr int SCATTER_NW = 128;
bool scatter_onehot_supported(int N, int V, int C);
int scatter_rows_onehot_tc(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, __nv_bfloat16* onehot,
                           float* dtable_ext, const float* cond, int B, int C, float* dwc, int perm_H, float* tmp,
                           cudaStream_t st);
int scatter_rows_by_token_bf16_w(const __nv_bfloat16* X, const int32_t* tok, long R, int N, int V, float* dtable,
                                 const float* cond, int B, int C, float* dwc, cudaStream_t st);

// fused persistent sampler (sampler_fused.cu): one kernel per batch, activations resident in shared memory
constexpr int SF_MAX_LAYERS = 4;
bool sampler_fused_supported(const arcvae_dims& d, int precision);
size_t sampler_fused_prep_bytes(const arcvae_dims& d);
// table [V,3H] fp32 (i|g|o), wc [3H,C]; Wxpb[l] / bp[l] tile-permuted compact weights / bias of layers >= 1;
// ended_count: 4 ints of scratch; prep: sampler_fused_prep_bytes() of scratch
int sampler_fused_run(const arcvae_dims& d, const float* table, const float* wc, __nv_bfloat16* const* Wxpb,
                      float* const* bp, const __nv_bfloat16* Woutb, const float* bout, const float* cond, int B,
                      int max_length, float temperature, int early_stopping, int multinomial, uint64_t seed,
                      int32_t* tokens, int32_t* t_stop, int32_t* ended_count, void* prep, cudaStream_t st);

// the same logical matrix in fp32 and (optionally) bf16; gemm_any picks the tensor-core kernel when precision is bf16
// and the operands satisfy the TMA constraints, else the fp32 FFMA kernel on the fp32 copies
struct Mat {
  const float* f;
  const __nv_bfloat16* b;
  int ld;
};
// C[M,N] (+)= op(A) op(B) + bias; transA/transB as gemm_f32.  (1,1) is not supported.
int gemm_any(int precision, int transA, int transB, int M, int N, int K, Mat A, Mat B, float* C, int ldc,
             const float* bias, bool accumulate, RowMap rm, long a_rows_total, cudaStream_t st);

// Philox4x32-10: 4 x uint32 for (seed, counter = (offset + idx))
__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in (0,1]: (x + 1) * 2^-32 computed in fp32 after dropping to 24 bits -> never 0
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }

}  // namespace arcvae
