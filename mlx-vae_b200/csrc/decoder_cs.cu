// decoder_cs.cu — the decoder with `carry_state=True`.
//
// NOT the reference's behaviour: models/decoder.py builds (hidden, cell) from z and the conditions
// (initialize_hidden_state, :76-111, called at :143) and then never passes them to the LSTM layers (:165-168), so the
// reference decoder is stateless (F1).  This file implements the evidently intended model: every layer l is a real LSTM
// over the T positions with h_{-1} = (z_to_hidden(z) + condition_to_hidden(cond)) / 2 (the same vector in every layer,
// :95-106) and c_{-1} = 0 (:109), state carried from position to position, and its reverse pass (BPTT), which also
// returns d total / d z so that the reconstruction loss reaches the encoder.
//
// Execution: step-major (t outer, layers inner), because the token fed at t+1 may be the greedy output of step t
// (decoder.py:180-185).  Built from the library's existing kernels: fp32 FFMA or tcgen05 GEMMs (gemm_any), the fp32 LSTM
// cell kernels of the per-step encoder path, the token scatter and the greedy feedback.  Layer 0's input projection is a
// table gather: table4 = Emb @ Wx0[:, :E]^T + b0  ([V,4H]) plus the rank-C term cond (x) Wx0[:, E:].
// This is an extension mode, not the benchmarked path: 3-4 launches per layer and position.
#include <vector>

#include "kernels.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

struct CsTape {
  int32_t* in_tok;                     // [T,B]
  int* tl;                             // [T] identity timestep list (argmax_feedback takes device lists)
  float* table4;                       // [V,4H]
  float* wc4;                          // [4H,C]
  float* h0;                           // [B,H]
  float* tmp;                          // [B,H]
  bf16* h0b;
  float* gates[ARCVAE_MAX_LAYERS];     // [R,4H] activated gates after forward, dA after backward
  float* c[ARCVAE_MAX_LAYERS];         // [R,H]
  float* h[ARCVAE_MAX_LAYERS];         // [R,H]
  bf16* hb[ARCVAE_MAX_LAYERS];         // [R,H]
  bf16* Whb[ARCVAE_MAX_LAYERS];        // [4H,H]
  bf16* Wxb[ARCVAE_MAX_LAYERS];        // [4H,H], l >= 1
  bf16* Woutb;                         // [V,H]
};

static size_t cs_tape_layout(const arcvae_dims& d, int B, int T, void* base, size_t cap, CsTape* t) {
  Arena a(base, cap);
  CsTape tt{};
  const size_t R = (size_t)T * B, H = d.H;
  tt.in_tok = a.take<int32_t>(R);
  tt.tl = a.take<int>((size_t)T + 1);
  tt.table4 = a.take<float>((size_t)d.V * 4 * H);
  tt.wc4 = a.take<float>((size_t)4 * H * d.C);
  tt.h0 = a.take<float>((size_t)B * H);
  tt.tmp = a.take<float>((size_t)B * H);
  tt.h0b = a.take<bf16>((size_t)B * H);
  for (int l = 0; l < d.NL; l++) {
    tt.gates[l] = a.take<float>(R * 4 * H);
    tt.c[l] = a.take<float>(R * H);
    tt.h[l] = a.take<float>(R * H);
    tt.hb[l] = a.take<bf16>(R * H);
    tt.Whb[l] = a.take<bf16>(4 * H * H);
    tt.Wxb[l] = l >= 1 ? a.take<bf16>(4 * H * H) : nullptr;
  }
  tt.Woutb = a.take<bf16>((size_t)d.V * H);
  if (t) *t = tt;
  return align_up(a.off, 256);
}

struct CsScratch {
  float* dX_top;                       // [R,H] d h of the top layer from the logits
  float* dxbuf;                        // [B,H] d h handed to the layer below at the current position
  float* dh_carry[ARCVAE_MAX_LAYERS];  // [B,H] recurrent d h_{t-1}
  float* dc[ARCVAE_MAX_LAYERS];        // [B,H]
  bf16* dAb[ARCVAE_MAX_LAYERS];        // [R,4H] bf16 copies of dA (tensor-core operands)
  float* dtable4;                      // [V,4H]
  float* dwc4;                         // [4H,C]
  float* dhinit;                       // [B,H]
  bf16* dlb;                           // [R,V] (V % 8 == 0 only)
};

static size_t cs_scratch_layout(const arcvae_dims& d, int B, int T, void* base, size_t cap, CsScratch* s) {
  Arena a(base, cap);
  CsScratch ss{};
  const size_t R = (size_t)T * B, H = d.H;
  ss.dX_top = a.take<float>(R * H);
  ss.dxbuf = a.take<float>((size_t)B * H);
  for (int l = 0; l < d.NL; l++) {
    ss.dh_carry[l] = a.take<float>((size_t)B * H);
    ss.dc[l] = a.take<float>((size_t)B * H);
    ss.dAb[l] = a.take<bf16>(R * 4 * H);
  }
  ss.dtable4 = a.take<float>((size_t)d.V * 4 * H);
  ss.dwc4 = a.take<float>((size_t)4 * H * d.C);
  ss.dhinit = a.take<float>((size_t)B * H);
  ss.dlb = a.take<bf16>(R * d.V);
  if (s) *s = ss;
  return align_up(a.off, 256);
}

// gates0[b, n] = table4[tok[b], n] + sum_c cond[b,c] * wc4[n,c]      (layer-0 input projection of one position)
__global__ void k_cs_pre0(const float* __restrict__ table4, const float* __restrict__ wc4, const int32_t* __restrict__ tok,
                          const float* __restrict__ cond, int B, int C, int G4, float* __restrict__ gates) {
  const long total = (long)B * G4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / G4), n = (int)(i - (long)b * G4);
    float v = table4[(long)tok[b] * G4 + n];
    for (int c = 0; c < C; c++) v = fmaf(cond[b * C + c], wc4[n * C + c], v);
    gates[i] = v;
  }
}
// out = (a + b) / 2 (+ bf16 copy)
__global__ void k_cs_avg2(const float* __restrict__ a, const float* __restrict__ b, long n, float* __restrict__ out,
                          bf16* __restrict__ outb) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = 0.5f * (a[i] + b[i]);
    out[i] = v;
    outb[i] = __float2bfloat16(v);
  }
}
struct PtrList { const float* p[ARCVAE_MAX_LAYERS]; };
// out = 0.5 * sum_l in[l]
__global__ void k_cs_half_sum(PtrList in, int nl, long n, float* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int l = 0; l < nl; l++) v += in.p[l][i];
    out[i] = 0.5f * v;
  }
}
__global__ void k_cs_iota(int* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = i;
}
// z = mu + eps * exp(logvar / 2)  =>  d mu += dz ; d logvar += dz * (z - mu) / 2
__global__ void k_reparam_bwd(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ mu,
                              long n, float* __restrict__ dmu, float* __restrict__ dlogvar) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float g = dz[i];
    dmu[i] += g;
    dlogvar[i] += g * 0.5f * (z[i] - mu[i]);
  }
}

static inline int grid1d(long n) {
  long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return (int)(g < 1 ? 1 : g);
}

static int check_dims_cs(const arcvae_dims* d) {
  ARCVAE_REQUIRE(d != nullptr, "dims");
  ARCVAE_REQUIRE(d->V > 0 && d->E > 0 && d->H > 0 && d->C > 0 && d->L > 0, "positive dims");
  ARCVAE_REQUIRE(d->NL >= 1 && d->NL <= ARCVAE_MAX_LAYERS, "num_layers in [1, 8]");
  return 0;
}

}  // namespace arcvae

using namespace arcvae;

extern "C" size_t arcvae_decoder_cs_tape_bytes(const arcvae_dims* d, int B, int T) {
  return d ? cs_tape_layout(*d, B, T, nullptr, 0, nullptr) : 0;
}
extern "C" size_t arcvae_decoder_cs_scratch_bytes(const arcvae_dims* d, int B, int T) {
  return d ? cs_scratch_layout(*d, B, T, nullptr, 0, nullptr) : 0;
}

extern "C" int arcvae_decoder_cs_forward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* z,
                                         const float* cond, const int32_t* target, const uint8_t* tf_mask_host, int B,
                                         int T, float* logits_tm, int32_t* dec_inputs_tm, void* tape, size_t tape_bytes,
                                         int precision, void* stream) {
  ARCVAE_TRY(check_dims_cs(d));
  ARCVAE_REQUIRE(B > 0 && T > 0, "empty batch / sequence");
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(z != nullptr && cond != nullptr && logits_tm != nullptr, "z, cond and the logits output are mandatory");
  cudaStream_t st = (cudaStream_t)stream;
  CsTape tp;
  const size_t need = cs_tape_layout(*d, B, T, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "carry-state decoder tape too small");
  const int H = d->H, G4 = 4 * d->H, V = d->V, E = d->E, C = d->C, L = d->L, top = d->NL - 1;
  const long R = (long)T * B;
  const bool bf = precision == ARCVAE_PREC_BF16;
  RowMap id{nullptr, 1};

  std::vector<uint8_t> coin(T, 0);       // decoder.py:180, one coin per position for the whole batch
  for (int t = 0; t < T; t++) coin[t] = (target != nullptr && tf_mask_host != nullptr && tf_mask_host[t]) ? 1 : 0;

  // parameters-only preparation
  ARCVAE_TRY(gemm_f32(0, 1, V, G4, E, p->embedding, E, p->Wx[0], E + C, tp.table4, G4, p->bias[0], false, id, 1, st));
  ARCVAE_CUDA(cudaMemcpy2DAsync(tp.wc4, (size_t)C * sizeof(float), p->Wx[0] + E, (size_t)(E + C) * sizeof(float),
                                (size_t)C * sizeof(float), G4, cudaMemcpyDeviceToDevice, st));
  if (bf) {
    for (int l = 0; l < d->NL; l++) {
      ARCVAE_TRY(f32_to_bf16(p->Wh[l], tp.Whb[l], (long)G4 * H, st));
      if (l >= 1) ARCVAE_TRY(f32_to_bf16(p->Wx[l], tp.Wxb[l], (long)G4 * H, st));
    }
    ARCVAE_TRY(f32_to_bf16(p->fc_out_w, tp.Woutb, (long)V * H, st));
  }
  k_cs_iota<<<1, 256, 0, st>>>(tp.tl, T);
  ARCVAE_LAUNCHED();
  // initialize_hidden_state (decoder.py:95-111): h_{-1} = (z Wz^T + bz + cond Wc^T + bc) / 2 for every layer, c_{-1} = 0
  ARCVAE_TRY(gemm_f32(0, 1, B, H, L, z, L, p->z_to_hidden_w, L, tp.h0, H, p->z_to_hidden_b, false, id, 1, st));
  ARCVAE_TRY(gemm_f32(0, 1, B, H, C, cond, C, p->condition_to_hidden_w, C, tp.tmp, H, p->condition_to_hidden_b, false, id, 1, st));
  k_cs_avg2<<<grid1d((long)B * H), 256, 0, st>>>(tp.h0, tp.tmp, (long)B * H, tp.h0, tp.h0b);
  ARCVAE_LAUNCHED();
  // token fed at position t: 0 at t = 0 (decoder.py:146), x[:, t-1] where coin[t-1], else the greedy output of t-1
  ARCVAE_CUDA(cudaMemsetAsync(tp.in_tok, 0, (size_t)R * sizeof(int32_t), st));
  if (target != nullptr) {
    for (int t = 1; t < T; t++)
      if (coin[t - 1])
        ARCVAE_CUDA(cudaMemcpy2DAsync(tp.in_tok + (long)t * B, sizeof(int32_t), target + (t - 1), (size_t)T * sizeof(int32_t),
                                      sizeof(int32_t), B, cudaMemcpyDeviceToDevice, st));
  }

  for (int t = 0; t < T; t++) {
    const long r0 = (long)t * B;
    for (int l = 0; l <= top; l++) {
      float* g_t = tp.gates[l] + r0 * G4;
      const float* hprev = t > 0 ? tp.h[l] + (r0 - B) * H : tp.h0;
      const bf16* hprevb = t > 0 ? tp.hb[l] + (r0 - B) * H : tp.h0b;
      if (l == 0) {
        k_cs_pre0<<<grid1d((long)B * G4), 256, 0, st>>>(tp.table4, tp.wc4, tp.in_tok + r0, cond, B, C, G4, g_t);
        ARCVAE_LAUNCHED();
      } else {
        ARCVAE_TRY(gemm_any(precision, 0, 1, B, G4, H, Mat{tp.h[l - 1] + r0 * H, bf ? tp.hb[l - 1] + r0 * H : nullptr, H},
                            Mat{p->Wx[l], bf ? tp.Wxb[l] : nullptr, H}, g_t, G4, p->bias[l], false, id, B, st));
      }
      // nn.LSTM with hidden given: ifgo += hidden @ Wh^T — at EVERY position here, including the first (h_{-1} != 0)
      ARCVAE_TRY(gemm_any(precision, 0, 1, B, G4, H, Mat{hprev, bf ? hprevb : nullptr, H}, Mat{p->Wh[l], bf ? tp.Whb[l] : nullptr, H},
                          g_t, G4, nullptr, true, id, B, st));
      ARCVAE_TRY(lstm_cell_fwd(g_t, t > 0 ? tp.c[l] + (r0 - B) * H : nullptr, tp.c[l] + r0 * H, tp.h[l] + r0 * H,
                               tp.hb[l] + r0 * H, B, H, st));
    }
    if (t + 1 < T && !coin[t]) {          // decoder.py:185: the next input is argmax(logits_t)
      ARCVAE_TRY(gemm_any(precision, 0, 1, B, V, H, Mat{tp.h[top] + r0 * H, bf ? tp.hb[top] + r0 * H : nullptr, H},
                          Mat{p->fc_out_w, bf ? tp.Woutb : nullptr, H}, logits_tm + r0 * V, V, p->fc_out_b, false, id, B, st));
      ARCVAE_TRY(argmax_feedback(logits_tm, tp.tl + t, 1, B, V, tp.in_tok, st));
    }
  }
  // fc_out for all positions at once (decoder.py:175)
  ARCVAE_TRY(gemm_any(precision, 0, 1, (int)R, V, H, Mat{tp.h[top], bf ? tp.hb[top] : nullptr, H},
                      Mat{p->fc_out_w, bf ? tp.Woutb : nullptr, H}, logits_tm, V, p->fc_out_b, false, id, R, st));
  if (dec_inputs_tm != nullptr)
    ARCVAE_CUDA(cudaMemcpyAsync(dec_inputs_tm, tp.in_tok, (size_t)R * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int arcvae_decoder_cs_backward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* z,
                                          const float* cond, int B, int T, float* dlogits_tm, void* tape, size_t tape_bytes,
                                          const arcvae_decoder_params* g, float* dz, void* scratch, size_t scratch_bytes,
                                          int precision, void* stream) {
  ARCVAE_TRY(check_dims_cs(d));
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(g != nullptr && dlogits_tm != nullptr && z != nullptr && cond != nullptr, "grad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  CsTape tp;
  const size_t need = cs_tape_layout(*d, B, T, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "carry-state decoder tape too small");
  CsScratch sc;
  const size_t need_s = cs_scratch_layout(*d, B, T, scratch, scratch_bytes, &sc);
  ARCVAE_REQUIRE(scratch != nullptr && need_s <= scratch_bytes, "carry-state decoder scratch too small");
  const int H = d->H, G4 = 4 * d->H, V = d->V, E = d->E, C = d->C, L = d->L, top = d->NL - 1;
  const long R = (long)T * B;
  const bool bf = precision == ARCVAE_PREC_BF16;
  const bool bfl = bf && (V % 8) == 0;     // bf16 copy of dlogits only when its pitch is TMA-legal
  RowMap id{nullptr, 1};

  if (bfl) ARCVAE_TRY(f32_to_bf16(dlogits_tm, sc.dlb, R * V, st));
  // fc_out: logits = h_top @ Wout^T + b
  ARCVAE_TRY(gemm_any(precision, 1, 0, V, H, (int)R, Mat{dlogits_tm, bfl ? sc.dlb : nullptr, V},
                      Mat{tp.h[top], bfl ? tp.hb[top] : nullptr, H}, g->fc_out_w, H, nullptr, true, id, R, st));
  ARCVAE_TRY(colsum(dlogits_tm, R, V, V, g->fc_out_b, st));
  ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, V, Mat{dlogits_tm, bfl ? sc.dlb : nullptr, V},
                      Mat{p->fc_out_w, bfl ? tp.Woutb : nullptr, H}, sc.dX_top, H, nullptr, false, id, R, st));
  for (int l = 0; l <= top; l++) {
    ARCVAE_CUDA(cudaMemsetAsync(sc.dh_carry[l], 0, (size_t)B * H * sizeof(float), st));
    ARCVAE_CUDA(cudaMemsetAsync(sc.dc[l], 0, (size_t)B * H * sizeof(float), st));
  }
  // BPTT, position by position, top layer first (the greedy feedback carries no gradient: argmax)
  for (int t = T - 1; t >= 0; t--) {
    const long r0 = (long)t * B;
    for (int l = top; l >= 0; l--) {
      float* g_t = tp.gates[l] + r0 * G4;
      bf16* dAb_t = bf ? sc.dAb[l] + r0 * G4 : nullptr;
      ARCVAE_TRY(lstm_cell_bwd(g_t, tp.c[l] + r0 * H, t > 0 ? tp.c[l] + (r0 - B) * H : nullptr,
                               l == top ? sc.dX_top + r0 * H : sc.dxbuf, sc.dh_carry[l], sc.dc[l], dAb_t, B, H, st));
      // d h_{t-1} of this layer (at t = 0: the gradient of the initial state)
      ARCVAE_TRY(gemm_any(precision, 0, 0, B, H, G4, Mat{g_t, dAb_t, G4}, Mat{p->Wh[l], bf ? tp.Whb[l] : nullptr, H},
                          sc.dh_carry[l], H, nullptr, false, id, B, st));
      if (l > 0)
        ARCVAE_TRY(gemm_any(precision, 0, 0, B, H, G4, Mat{g_t, dAb_t, G4}, Mat{p->Wx[l], bf ? tp.Wxb[l] : nullptr, H},
                            sc.dxbuf, H, nullptr, false, id, B, st));
    }
  }
  // initial state: h_{-1} = (z Wz^T + bz + cond Wc^T + bc) / 2, shared by all layers
  {
    PtrList pl{};
    for (int l = 0; l <= top; l++) pl.p[l] = sc.dh_carry[l];
    k_cs_half_sum<<<grid1d((long)B * H), 256, 0, st>>>(pl, d->NL, (long)B * H, sc.dhinit);
    ARCVAE_LAUNCHED();
  }
  ARCVAE_TRY(gemm_f32(1, 0, H, L, B, sc.dhinit, H, z, L, g->z_to_hidden_w, L, nullptr, true, id, pick_splitk(H, L, B), st));
  ARCVAE_TRY(colsum(sc.dhinit, B, H, H, g->z_to_hidden_b, st));
  ARCVAE_TRY(gemm_f32(1, 0, H, C, B, sc.dhinit, H, cond, C, g->condition_to_hidden_w, C, nullptr, true, id, pick_splitk(H, C, B), st));
  ARCVAE_TRY(colsum(sc.dhinit, B, H, H, g->condition_to_hidden_b, st));
  if (dz != nullptr)
    ARCVAE_TRY(gemm_f32(0, 0, B, L, H, sc.dhinit, H, p->z_to_hidden_w, L, dz, L, nullptr, false, id, 1, st));

  // weight gradients from the dA tapes
  for (int l = top; l >= 0; l--) {
    float* dA = tp.gates[l];
    const bf16* dAb = bf ? sc.dAb[l] : nullptr;
    if (T > 1) {                            // dWh += dA[1:]^T h[:-1]
      const long K = R - B;
      ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)K, Mat{dA + (long)B * G4, bf ? dAb + (long)B * G4 : nullptr, G4},
                          Mat{tp.h[l], bf ? tp.hb[l] : nullptr, H}, g->Wh[l], H, nullptr, true, id, K, st));
    }
    ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, B, Mat{dA, dAb, G4}, Mat{tp.h0, bf ? tp.h0b : nullptr, H}, g->Wh[l], H,
                        nullptr, true, id, B, st));                     // ... + dA_0^T h_{-1}
    if (l > 0) {
      ARCVAE_TRY(colsum(dA, R, G4, G4, g->bias[l], st));
      ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)R, Mat{dA, dAb, G4}, Mat{tp.h[l - 1], bf ? tp.hb[l - 1] : nullptr, H},
                          g->Wx[l], H, nullptr, true, id, R, st));
    } else {
      // a0 = table4[tok] + cond @ wc4^T ; table4 = Emb @ Wx0[:, :E]^T + b0
      ARCVAE_CUDA(cudaMemsetAsync(sc.dtable4, 0, (size_t)V * G4 * sizeof(float), st));
      ARCVAE_CUDA(cudaMemsetAsync(sc.dwc4, 0, (size_t)G4 * C * sizeof(float), st));
      ARCVAE_TRY(scatter_rows_by_token(dA, tp.in_tok, R, G4, V, sc.dtable4, cond, B, C, sc.dwc4, st));
      ARCVAE_TRY(colsum(sc.dtable4, V, G4, G4, g->bias[0], st));
      ARCVAE_TRY(gemm_f32(0, 0, V, E, G4, sc.dtable4, G4, p->Wx[0], E + C, g->embedding, E, nullptr, true, id,
                          pick_splitk(V, E, G4), st));
      ARCVAE_TRY(gemm_f32(1, 0, G4, E, V, sc.dtable4, G4, p->embedding, E, g->Wx[0], E + C, nullptr, true, id, 1, st));
      ARCVAE_TRY(add_strided(sc.dwc4, C, g->Wx[0] + E, E + C, G4, C, st));
    }
  }
  return 0;
}

// d z -> (d mu, d logvar) through z = mu + eps * exp(logvar / 2) (models/encoder.py:147-153): accumulates
extern "C" int arcvae_reparam_backward(const float* dz, const float* z, const float* mu, int B, int L, float* dmu,
                                       float* dlogvar, void* stream) {
  ARCVAE_REQUIRE(dz && z && mu && dmu && dlogvar, "pointers");
  const long n = (long)B * L;
  if (n <= 0) return 0;
  k_reparam_bwd<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(dz, z, mu, n, dmu, dlogvar);
  ARCVAE_LAUNCHED();
  return 0;
}
