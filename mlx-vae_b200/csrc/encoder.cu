// encoder.cu — MLXEncoder.__call__ (models/encoder.py:76-132) forward and its reverse pass (BPTT).
//
// Internal layout is TIME-MAJOR: row r = t*B + b, so the rows of one timestep are contiguous.
//   layer 0 input projection  = gather of table0 = Emb @ Wx0^T + b0  ([V,4H]; V=80 rows make the product a table)
//   layer l>0 input projection = H_{l-1} @ Wx_l^T + b_l               (time-parallel GEMM over all T*B rows)
//   recurrence                 = MLX nn.LSTM semantics: zero initial state, gate order i,f,g,o, c_0 = i*g
//   head                       = [h_T ; Linear(cond)] -> fc_mu / fc_logvar(_hidden) -> tanh bounds
// Execution paths, same math:
//   ARCVAE_PREC_FP32              per-step fp32 FFMA GEMM + cell kernels (reference precision)
//   ARCVAE_PREC_BF16, H == 256    persistent cluster kernel per layer and direction (lstm_cluster.cu): W_hh resident in
//                                 shared memory for all T steps, bf16 tapes (cell-reverse coefficients, h, dA), cell
//                                 state in registers, flag-in-data exchange between the cluster's CTAs
//   ARCVAE_PREC_BF16, H % 64 == 0 (e.g. the scaled config, H = 1024: W_hh = 8 MB bf16 does not fit any cluster's shared
//                                 memory) ONE launch per timestep: tcgen05 GEMM h_{t-1} @ Wh^T with the whole LSTM cell in
//                                 its epilogue (gemm_tc.cu TC_EPI_LSTM_FWD); W_hh split over the N tiles of the grid
//                                 and L2-resident across steps; bf16 tapes; reverse = cell kernel + split-K GEMM per step
//   ARCVAE_PREC_BF16, otherwise   per-step tcgen05 GEMM + fp32 cell kernels
#include <cstdlib>

#include "kernels.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

enum { PATH_STEP_F32 = 0, PATH_STEP_BF16 = 1, PATH_CLUSTER = 2, PATH_STEP_FUSED = 3, PATH_COUNT = 4 };

static int pick_path(const arcvae_dims& d, int precision) {
  if (precision != ARCVAE_PREC_BF16) return PATH_STEP_F32;
  const bool no_cluster = std::getenv("ARCVAE_NO_CLUSTER") != nullptr;   // read per call: tests flip it
  if (lstm_cluster_supported(d.H) && !no_cluster) return PATH_CLUSTER;
  if ((d.H % 64) == 0 && std::getenv("ARCVAE_NO_FUSED_STEP") == nullptr) return PATH_STEP_FUSED;
  return PATH_STEP_BF16;
}

struct EncTape {
  int32_t* xT;       // [T,B]
  float* table0;     // [V,4H]
  float* u;          // [B,2H]
  float* lvh;        // [B,2H]  tanh(fc_logvar_hidden(u))
  float* mu_raw;     // [B,L]
  float* lv_raw;     // [B,L]
  float* mu;         // [B,L]
  float* logvar;     // [B,L]
  float* h_last;     // [B,H]   h_{T-1} of the top layer (cluster path)
  float* c[ARCVAE_MAX_LAYERS];      // [T*B,H] fp32 cell state (all paths)
  // per-step paths
  float* gates[ARCVAE_MAX_LAYERS];  // [T*B,4H]  activated gates after forward, dA after backward
  float* h[ARCVAE_MAX_LAYERS];      // [T*B,H]
  // bf16 paths
  bf16* hb[ARCVAE_MAX_LAYERS];      // [T*B,H]
  bf16* Whb[ARCVAE_MAX_LAYERS];     // [4H,H]
  bf16* Wxb[ARCVAE_MAX_LAYERS];     // [4H,H], l >= 1
  // cluster path
  bf16* gates_b[ARCVAE_MAX_LAYERS]; // fused per-step path: [T*B,4H] activated gates
  bf16* ktape[ARCVAE_MAX_LAYERS];   // cluster path: coefficient tape of the cell reverse (lstm_cluster.cu, RC_KV)
  void* xh;                         // forward exchange buffer of the cluster kernel (flag-in-data vectors of h_t)
  // bf16 operands of the tensor-core head products (bf16 paths)
  bf16* ub;          // [B,2H]
  bf16* lvhb;        // [B,2H]
  bf16* Wmub;        // [L,2H]
  bf16* Wlhb;        // [2H,2H]
  bf16* Wlvb;        // [L,2H]
  bf16* Pb;                         // [T*B,4H] input projection of the layer being run (reused)
  bf16* table0b;                    // [V,4H] bf16 copy of table0
  // fused per-step path
  bf16* Whp[ARCVAE_MAX_LAYERS];     // [4H,H] tile-permuted Wh (B operand of the fused LSTM-step GEMM)
  bf16* hzero;                      // [B,H] zeros: h_{-1}
};

static size_t enc_tape_layout(const arcvae_dims& d, int B, int T, int path, void* base, size_t cap, EncTape* t) {
  Arena a(base, cap);
  const size_t R = (size_t)T * B, H = d.H;
  EncTape tt{};
  tt.xT = a.take<int32_t>(R);
  tt.table0 = a.take<float>((size_t)d.V * 4 * H);
  tt.u = a.take<float>((size_t)B * 2 * H);
  tt.lvh = a.take<float>((size_t)B * 2 * H);
  tt.mu_raw = a.take<float>((size_t)B * d.L);
  tt.lv_raw = a.take<float>((size_t)B * d.L);
  tt.mu = a.take<float>((size_t)B * d.L);
  tt.logvar = a.take<float>((size_t)B * d.L);
  tt.h_last = a.take<float>((size_t)B * H);
  if (path != PATH_STEP_F32) {
    tt.ub = a.take<bf16>((size_t)B * 2 * H);
    tt.lvhb = a.take<bf16>((size_t)B * 2 * H);
    tt.Wmub = a.take<bf16>((size_t)d.L * 2 * H);
    tt.Wlhb = a.take<bf16>((size_t)4 * H * H);
    tt.Wlvb = a.take<bf16>((size_t)d.L * 2 * H);
  }
  for (int l = 0; l < d.NL; l++) {
    if (path != PATH_CLUSTER) tt.c[l] = a.take<float>(R * H);   // the cluster kernels keep c_t in registers; their tape holds coefficients
    if (path == PATH_STEP_FUSED) {
      tt.gates_b[l] = a.take<bf16>(R * 4 * H);
      tt.Whp[l] = a.take<bf16>(4 * H * H);
    }
    if (path != PATH_CLUSTER && path != PATH_STEP_FUSED) {
      tt.gates[l] = a.take<float>(R * 4 * H);
      tt.h[l] = a.take<float>(R * H);
    }
    if (path != PATH_STEP_F32) {
      tt.hb[l] = a.take<bf16>(R * H);
      tt.Whb[l] = a.take<bf16>(4 * H * H);
      if (l >= 1) tt.Wxb[l] = a.take<bf16>(4 * H * H);
    }
    if (path == PATH_CLUSTER) {
      tt.ktape[l] = a.take<bf16>(lstm_cluster_ktape_elems(B, T, (int)H));
      if (l == 0) tt.xh = a.take<char>(lstm_cluster_xh_bytes(B));
    }
  }
  if (path == PATH_CLUSTER || path == PATH_STEP_FUSED) {
    tt.Pb = a.take<bf16>(R * 4 * H);
    tt.table0b = a.take<bf16>((size_t)d.V * 4 * H);
  }
  if (path == PATH_STEP_FUSED) tt.hzero = a.take<bf16>((size_t)B * H);
  if (t) *t = tt;
  return align_up(a.off, 256);
}

struct EncScratch {
  float* dmu_raw;   // [B,L]
  float* dlv_raw;   // [B,L]
  float* dlvh;      // [B,2H]
  float* du;        // [B,2H]
  float* dX;        // [T*B,H]  gradient w.r.t. a layer's input sequence (from the layer above)
  float* dh_rec[2]; // [B,H]
  float* dc;        // [B,H]
  float* dtable0;   // [max(V, SCATTER_NW),4H]  (rows >= V: scratch of the one-hot GEMM)
  bf16* dAb;        // [T*B,4H] bf16 pre-activation gradients (tensor-core operand / cluster exchange)
  bf16* dmub;       // [B,L]   bf16 copies for the tensor-core head products
  bf16* dlvb;       // [B,L]
  bf16* dlvhb;      // [B,2H]
  bf16* onehot;     // [T*B,SCATTER_NW] one-hot tokens (tensor-core scatter)
  void* xch;        // cluster backward: exchange buffers of the partial d h
  bf16* dXtf;       // cluster backward: d h for the layer below, bf16, thread-friendly layout (gemm_tc TC_EPI_LSTM_DH)
  float* segtmp;    // [4H,SCATTER_NW] one-hot segment of the fused weight-gradient GEMM (dtable0^T / bias row sums)
  float* dh_part;   // [4][B,H] split-K partials of d h_{t-1} (fused per-step path)
};

static size_t enc_scratch_layout(const arcvae_dims& d, int B, int T, int path, void* base, size_t cap, EncScratch* s) {
  Arena a(base, cap);
  EncScratch ss{};
  ss.dmu_raw = a.take<float>((size_t)B * d.L);
  ss.dlv_raw = a.take<float>((size_t)B * d.L);
  ss.dlvh = a.take<float>((size_t)B * 2 * d.H);
  ss.du = a.take<float>((size_t)B * 2 * d.H);
  ss.dX = a.take<float>(d.NL > 1 ? (size_t)T * B * d.H : 1);
  ss.dh_rec[0] = a.take<float>((size_t)B * d.H);
  ss.dh_rec[1] = a.take<float>((size_t)B * d.H);
  ss.dc = a.take<float>((size_t)B * d.H);
  ss.dtable0 = a.take<float>((size_t)(d.V > SCATTER_NW ? d.V : SCATTER_NW) * 4 * d.H);
  ss.dAb = (path != PATH_STEP_F32) ? a.take<bf16>((size_t)T * B * 4 * d.H) : nullptr;
  if (path != PATH_STEP_F32) {
    ss.dmub = a.take<bf16>((size_t)B * d.L);
    ss.dlvb = a.take<bf16>((size_t)B * d.L);
    ss.dlvhb = a.take<bf16>((size_t)B * 2 * d.H);
  }
  ss.onehot = (path != PATH_STEP_F32) ? a.take<bf16>((size_t)T * B * SCATTER_NW) : nullptr;
  ss.xch = (path != PATH_STEP_F32) ? a.take<char>(lstm_cluster_xch_bytes(B)) : nullptr;
  ss.dXtf = (path != PATH_STEP_F32 && d.NL > 1) ? a.take<bf16>(lstm_cluster_dh_tf_elems(B, T, d.H)) : nullptr;
  ss.segtmp = (path != PATH_STEP_F32) ? a.take<float>((size_t)4 * d.H * SCATTER_NW) : nullptr;
  ss.dh_part = (path != PATH_STEP_F32) ? a.take<float>((size_t)4 * B * d.H) : nullptr;
  if (s) *s = ss;
  return align_up(a.off, 256);
}

static int check_dims(const arcvae_dims* d) {
  ARCVAE_REQUIRE(d != nullptr, "dims");
  ARCVAE_REQUIRE(d->V > 0 && d->E > 0 && d->H > 0 && d->L > 0 && d->C > 0, "positive dims");
  ARCVAE_REQUIRE(d->NL >= 1 && d->NL <= ARCVAE_MAX_LAYERS, "num_layers in [1, 8]");
  return 0;
}

// ---- head: encoder.py:106-130 and its reverse ----------------------------------------------------------------------
static int head_forward(const arcvae_dims& d, const arcvae_encoder_params* p, const EncTape& tp, const float* h_last,
                        const float* cond, int B, float* mu, float* logvar, int precision, cudaStream_t st) {
  const int H = d.H, H2 = 2 * d.H;
  RowMap id{nullptr, 1};
  const bool bf = precision == ARCVAE_PREC_BF16;      // bf16 mode: the head products run on the tensor cores as well
  if (bf) {
    ARCVAE_TRY(f32_to_bf16(p->fc_mu_w, tp.Wmub, (long)d.L * H2, st));
    ARCVAE_TRY(f32_to_bf16(p->fc_logvar_hidden_w, tp.Wlhb, (long)H2 * H2, st));
    ARCVAE_TRY(f32_to_bf16(p->fc_logvar_w, tp.Wlvb, (long)d.L * H2, st));
  }
  ARCVAE_TRY(head_build_u(h_last, cond, p->condition_fc_w, p->condition_fc_b, B, H, d.C, tp.u, bf ? tp.ub : nullptr, st));
  ARCVAE_TRY(gemm_any(precision, 0, 1, B, d.L, H2, Mat{tp.u, bf ? tp.ub : nullptr, H2}, Mat{p->fc_mu_w, bf ? tp.Wmub : nullptr, H2},
                      tp.mu_raw, d.L, p->fc_mu_b, false, id, B, st));
  ARCVAE_TRY(gemm_any(precision, 0, 1, B, H2, H2, Mat{tp.u, bf ? tp.ub : nullptr, H2},
                      Mat{p->fc_logvar_hidden_w, bf ? tp.Wlhb : nullptr, H2}, tp.lvh, H2, p->fc_logvar_hidden_b, false, id, B, st));
  ARCVAE_TRY(tanh_inplace(tp.lvh, (long)B * H2, bf ? tp.lvhb : nullptr, st));
  ARCVAE_TRY(gemm_any(precision, 0, 1, B, d.L, H2, Mat{tp.lvh, bf ? tp.lvhb : nullptr, H2},
                      Mat{p->fc_logvar_w, bf ? tp.Wlvb : nullptr, H2}, tp.lv_raw, d.L, p->fc_logvar_b, false, id, B, st));
  ARCVAE_TRY(head_bound(tp.mu_raw, tp.lv_raw, (long)B * d.L, tp.mu, tp.logvar, st));
  if (mu) ARCVAE_CUDA(cudaMemcpyAsync(mu, tp.mu, (size_t)B * d.L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (logvar) ARCVAE_CUDA(cudaMemcpyAsync(logvar, tp.logvar, (size_t)B * d.L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// leaves d h_T in sc.du[:, 0:H] (row pitch 2H)
static int head_backward(const arcvae_dims& d, const arcvae_encoder_params* p, const arcvae_encoder_params* g,
                         const EncTape& tp, const EncScratch& sc, const float* cond, int B, const float* dmu,
                         const float* dlogvar, int precision, cudaStream_t st) {
  const int H = d.H, L = d.L, H2 = 2 * d.H;
  RowMap id{nullptr, 1};
  const bool bf = precision == ARCVAE_PREC_BF16;
  auto M2 = [&](const float* f, const bf16* b, int ld) { return Mat{f, bf ? b : nullptr, ld}; };
  ARCVAE_TRY(head_bound_bwd(tp.mu, tp.logvar, dmu, dlogvar, (long)B * L, sc.dmu_raw, sc.dlv_raw, bf ? sc.dmub : nullptr,
                            bf ? sc.dlvb : nullptr, st));
  // fc_logvar: lv_raw = lvh @ Wlv^T + b
  ARCVAE_TRY(gemm_any(precision, 1, 0, L, H2, B, M2(sc.dlv_raw, sc.dlvb, L), M2(tp.lvh, tp.lvhb, H2), g->fc_logvar_w, H2, nullptr,
                      true, id, B, st));
  ARCVAE_TRY(colsum(sc.dlv_raw, B, L, L, g->fc_logvar_b, st));
  ARCVAE_TRY(gemm_any(precision, 0, 0, B, H2, L, M2(sc.dlv_raw, sc.dlvb, L), M2(p->fc_logvar_w, tp.Wlvb, H2), sc.dlvh, H2, nullptr,
                      false, id, B, st));
  ARCVAE_TRY(tanh_bwd_inplace(sc.dlvh, tp.lvh, (long)B * H2, bf ? sc.dlvhb : nullptr, st));
  // fc_logvar_hidden: pre = u @ Wlh^T + b
  ARCVAE_TRY(gemm_any(precision, 1, 0, H2, H2, B, M2(sc.dlvh, sc.dlvhb, H2), M2(tp.u, tp.ub, H2), g->fc_logvar_hidden_w, H2, nullptr,
                      true, id, B, st));
  ARCVAE_TRY(colsum(sc.dlvh, B, H2, H2, g->fc_logvar_hidden_b, st));
  ARCVAE_TRY(gemm_any(precision, 0, 0, B, H2, H2, M2(sc.dlvh, sc.dlvhb, H2), M2(p->fc_logvar_hidden_w, tp.Wlhb, H2), sc.du, H2,
                      nullptr, false, id, B, st));
  // fc_mu
  ARCVAE_TRY(gemm_any(precision, 1, 0, L, H2, B, M2(sc.dmu_raw, sc.dmub, L), M2(tp.u, tp.ub, H2), g->fc_mu_w, H2, nullptr, true, id,
                      B, st));
  ARCVAE_TRY(colsum(sc.dmu_raw, B, L, L, g->fc_mu_b, st));
  ARCVAE_TRY(gemm_any(precision, 0, 0, B, H2, L, M2(sc.dmu_raw, sc.dmub, L), M2(p->fc_mu_w, tp.Wmub, H2), sc.du, H2, nullptr, true,
                      id, B, st));
  // condition_fc: cproj = cond @ Wc^T + bc ; d cproj = du[:, H:2H]   (N = num_conditions: not tensor-core shaped)
  ARCVAE_TRY(gemm_f32(1, 0, H, d.C, B, sc.du + H, H2, cond, d.C, g->condition_fc_w, d.C, nullptr, true, id, pick_splitk(H, d.C, B), st));
  ARCVAE_TRY(colsum(sc.du + H, B, H, H2, g->condition_fc_b, st));
  return 0;
}

// layer 0: P0 = table0[x]  ->  dtable0 = onehot(x)^T @ dA ; table0 = Emb @ Wx0^T + b0
static int layer0_input_backward(const arcvae_dims& d, const arcvae_encoder_params* p, const arcvae_encoder_params* g,
                                 const EncScratch& sc, cudaStream_t st) {
  const int G4 = 4 * d.H;
  RowMap id{nullptr, 1};
  ARCVAE_TRY(colsum(sc.dtable0, d.V, G4, G4, g->bias[0], st));
  ARCVAE_TRY(gemm_f32(0, 0, d.V, d.E, G4, sc.dtable0, G4, p->Wx[0], d.E, g->embedding, d.E, nullptr, true, id,
                      pick_splitk(d.V, d.E, G4), st));
  ARCVAE_TRY(gemm_f32(1, 0, G4, d.E, d.V, sc.dtable0, G4, p->embedding, d.E, g->Wx[0], d.E, nullptr, true, id, 1, st));
  return 0;
}

}  // namespace arcvae

using namespace arcvae;

// sized for the largest layout so one buffer serves every precision / path
extern "C" size_t arcvae_encoder_tape_bytes(const arcvae_dims* d, int B, int T) {
  if (!d) return 0;
  size_t m = 0;
  for (int path = 0; path < PATH_COUNT; path++) {
    size_t s = enc_tape_layout(*d, B, T, path, nullptr, 0, nullptr);
    if (s > m) m = s;
  }
  return m;
}
extern "C" size_t arcvae_encoder_scratch_bytes(const arcvae_dims* d, int B, int T) {
  if (!d) return 0;
  return enc_scratch_layout(*d, B, T, PATH_STEP_BF16, nullptr, 0, nullptr);
}

extern "C" int arcvae_encoder_forward(const arcvae_dims* d, const arcvae_encoder_params* p, const int32_t* x,
                                      const float* cond, int B, int T, float* mu, float* logvar, void* tape,
                                      size_t tape_bytes, int precision, void* stream) {
  ARCVAE_TRY(check_dims(d));
  ARCVAE_REQUIRE(B > 0 && T > 0, "empty batch / sequence");
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  cudaStream_t st = (cudaStream_t)stream;
  const int path = pick_path(*d, precision);
  EncTape tp;
  size_t need = enc_tape_layout(*d, B, T, path, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "encoder tape too small");
  const int H = d->H, G4 = 4 * d->H;
  const long R = (long)T * B;
  RowMap id{nullptr, 1};
  const bool bf = path != PATH_STEP_F32;

  ARCVAE_TRY(transpose_tokens(x, B, T, tp.xT, st));
  // table0[v,:] = Emb[v,:] @ Wx0^T + b0   (encoder.py:93 gather + nn.LSTM's addmm folded: rows are tokens)
  ARCVAE_TRY(gemm_f32(0, 1, d->V, G4, d->E, p->embedding, d->E, p->Wx[0], d->E, tp.table0, G4, p->bias[0], false, id, 1, st));
  if (bf) {
    for (int l = 0; l < d->NL; l++) {
      ARCVAE_TRY(f32_to_bf16(p->Wh[l], tp.Whb[l], (long)G4 * H, st));
      if (l >= 1) ARCVAE_TRY(f32_to_bf16(p->Wx[l], tp.Wxb[l], (long)G4 * H, st));
    }
  }

  if (path == PATH_CLUSTER) {
    int* const errf = device_error_flag();      // sticky: raised by a bounded wait that ran out, never cleared here
    ARCVAE_REQUIRE(errf != nullptr, "device error flag allocation failed");
    ARCVAE_TRY(f32_to_bf16(tp.table0, tp.table0b, (long)d->V * G4, st));
    for (int l = 0; l < d->NL; l++) {
      bool p_in_tape = false;
      if (l >= 1) {
        // time-parallel input projection, bf16 out (bias included): P = hb_{l-1} @ Wx_l^T + b_l.  Preferred: the
        // weight-stationary GEMM writes it into THIS layer's gate tape in the tape's thread-friendly layout (the recurrence
        // reads it with 512-byte warp accesses and overwrites it with the activated gates); else row-major into Pb.
        TcGemm g{};
        g.M = (int)R; g.N = G4; g.K = H;
        g.A = tp.hb[l - 1]; g.lda = H; g.a_mn = false;
        g.B = tp.Wxb[l]; g.ldb = H; g.b_mn = false;
        g.C = nullptr; g.ldc = 0; g.bias = p->bias[l]; g.accumulate = false; g.splitk = 1;
        g.rm = id; g.a_rows_total = R;
        g.Cb = tp.ktape[l]; g.ldcb = G4; g.epi = TC_EPI_LSTM_P; g.Hh = H; g.lp_B = B; g.lp_vecs = 24;
        p_in_tape = gemm_ws_supported(g) && std::getenv("ARCVAE_NO_WS") == nullptr;
        if (p_in_tape) {
          ARCVAE_TRY(gemm_ws(g, st));
        } else {
          g.Cb = tp.Pb; g.epi = TC_EPI_PLAIN; g.Hh = 0; g.lp_B = 0; g.lp_vecs = 16;
          ARCVAE_TRY(gemm_tc(g, st));
        }
      }
      ARCVAE_TRY(lstm_cluster_forward(B, T, H, tp.Whb[l], tp.xT, l == 0 ? tp.table0b : nullptr,
                                      (l == 0 || p_in_tape) ? nullptr : tp.Pb, tp.hb[l], tp.ktape[l],
                                      l == d->NL - 1 ? tp.h_last : nullptr, tp.xh, errf, st));
    }
    return head_forward(*d, p, tp, tp.h_last, cond, B, mu, logvar, precision, st);
  }

  if (path == PATH_STEP_FUSED) {
    ARCVAE_TRY(f32_to_bf16(tp.table0, tp.table0b, (long)d->V * G4, st));
    ARCVAE_CUDA(cudaMemsetAsync(tp.hzero, 0, (size_t)B * H * sizeof(bf16), st));
    for (int l = 0; l < d->NL; l++) {
      ARCVAE_TRY(perm4_rows_to_bf16(p->Wh[l], H, H, tp.Whp[l], st));
      if (l == 0) {
        ARCVAE_TRY(gather_rows_bf16(tp.table0b, tp.xT, R, G4, tp.Pb, st));        // P0 = table0[x] (bias folded in)
      } else {
        TcGemm g{};                                                               // P_l = h_{l-1} @ Wx_l^T + b_l, all T*B rows
        g.M = (int)R; g.N = G4; g.K = H;
        g.A = tp.hb[l - 1]; g.lda = H; g.a_mn = false;
        g.B = tp.Wxb[l]; g.ldb = H; g.b_mn = false;
        g.C = nullptr; g.ldc = 0; g.Cb = tp.Pb; g.ldcb = G4; g.bias = p->bias[l]; g.accumulate = false; g.splitk = 1;
        g.rm = id; g.a_rows_total = R;
        ARCVAE_TRY(gemm_tc(g, st));
      }
      TimeScope ts(TIME_RECURRENCE, st);
      for (int t = 0; t < T; t++) {
        // one launch per step: gates = P_t + h_{t-1} @ Wh^T, cell, h_t — nn.LSTM's loop body in the GEMM epilogue
        const long r0 = (long)t * B;
        TcGemm g{};
        g.M = B; g.N = G4; g.K = H;
        g.A = t > 0 ? tp.hb[l] + (r0 - B) * H : tp.hzero; g.lda = H; g.a_mn = false;
        g.B = tp.Whp[l]; g.ldb = H; g.b_mn = false;
        g.accumulate = false; g.splitk = 1; g.rm = id; g.a_rows_total = B;
        g.epi = TC_EPI_LSTM_FWD; g.Hh = H;
        g.pre_b = tp.Pb + r0 * G4;
        g.c_prev = t > 0 ? tp.c[l] + (r0 - B) * H : nullptr;
        g.c_out = tp.c[l] + r0 * H;
        g.gates_b = tp.gates_b[l] + r0 * G4;
        g.hb_out = tp.hb[l] + r0 * H;
        g.hf_out = (l == d->NL - 1 && t == T - 1) ? tp.h_last : nullptr;
        ARCVAE_TRY(gemm_tc(g, st));
      }
    }
    return head_forward(*d, p, tp, tp.h_last, cond, B, mu, logvar, precision, st);
  }

  for (int l = 0; l < d->NL; l++) {
    if (l == 0) {
      ARCVAE_TRY(gather_rows(tp.table0, tp.xT, (int)R, G4, tp.gates[0], st));
    } else {
      // time-parallel input projection of layer l: all T*B rows at once
      ARCVAE_TRY(gemm_any(precision, 0, 1, (int)R, G4, H, Mat{tp.h[l - 1], tp.hb[l - 1], H}, Mat{p->Wx[l], tp.Wxb[l], H},
                          tp.gates[l], G4, p->bias[l], false, id, R, st));
    }
    TimeScope ts(TIME_RECURRENCE, st);
    for (int t = 0; t < T; t++) {
      float* g_t = tp.gates[l] + (long)t * B * G4;
      float* c_t = tp.c[l] + (long)t * B * H;
      float* h_t = tp.h[l] + (long)t * B * H;
      bf16* hb_t = bf ? tp.hb[l] + (long)t * B * H : nullptr;
      if (t > 0) {
        // gates_t += h_{t-1} @ Wh^T   (nn.LSTM: `ifgo = ifgo + hidden @ Wh.T` once hidden is not None)
        ARCVAE_TRY(gemm_any(precision, 0, 1, B, G4, H, Mat{h_t - (long)B * H, bf ? hb_t - (long)B * H : nullptr, H},
                            Mat{p->Wh[l], tp.Whb[l], H}, g_t, G4, nullptr, true, id, B, st));
      }
      ARCVAE_TRY(lstm_cell_fwd(g_t, t > 0 ? c_t - (long)B * H : nullptr, c_t, h_t, hb_t, B, H, st));
    }
  }
  const float* h_last = tp.h[d->NL - 1] + (long)(T - 1) * B * H;
  return head_forward(*d, p, tp, h_last, cond, B, mu, logvar, precision, st);
}

extern "C" int arcvae_encoder_backward(const arcvae_dims* d, const arcvae_encoder_params* p, const float* cond, int B,
                                       int T, const float* dmu, const float* dlogvar, void* tape, size_t tape_bytes,
                                       const arcvae_encoder_params* g, void* scratch, size_t scratch_bytes,
                                       int precision, void* stream) {
  ARCVAE_TRY(check_dims(d));
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(g != nullptr && dmu != nullptr && dlogvar != nullptr, "grad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  const int path = pick_path(*d, precision);
  EncTape tp;
  size_t need = enc_tape_layout(*d, B, T, path, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "encoder tape too small");
  EncScratch sc;
  size_t need_s = enc_scratch_layout(*d, B, T, path, scratch, scratch_bytes, &sc);
  ARCVAE_REQUIRE(scratch != nullptr && need_s <= scratch_bytes, "encoder scratch too small");
  const int H = d->H, G4 = 4 * d->H, H2 = 2 * d->H;
  const long R = (long)T * B;
  RowMap id{nullptr, 1};
  const bool bf = path != PATH_STEP_F32;

  ARCVAE_TRY(head_backward(*d, p, g, tp, sc, cond, B, dmu, dlogvar, precision, st));

  if (path == PATH_CLUSTER || path == PATH_STEP_FUSED) {
    // weight gradients that share dA^T (dWh, dWx, bias / table scatter) run as ONE multi-segment GEMM per layer: dA is read
    // from HBM once.  The one-hot token operand serves the layer-0 table scatter AND, through its row sums, the bias
    // gradients of the upper layers.
    int* const errf = device_error_flag();
    ARCVAE_REQUIRE(errf != nullptr, "device error flag allocation failed");
    const bool fuse_dw = path == PATH_CLUSTER && scatter_onehot_supported(G4, d->V, 0) && (H % 64) == 0 && H <= 256;
    const bool dx_tf = fuse_dw && H == 256;     // the d X GEMM feeds the next cluster kernel directly
    if (fuse_dw) ARCVAE_TRY(build_onehot(tp.xT, R, d->V, nullptr, B, 0, sc.onehot, st));
    for (int l = d->NL - 1; l >= 0; l--) {
      const bool top = (l == d->NL - 1);
      if (path == PATH_CLUSTER) {
        ARCVAE_TRY(lstm_cluster_backward(B, T, H, tp.Whb[l], tp.ktape[l], (top || dx_tf) ? nullptr : sc.dX, (top || !dx_tf) ? nullptr : sc.dXtf,
                                          top ? sc.du : nullptr, H2, sc.dAb, sc.xch, errf, st));
      } else {
        // BPTT one step per launch pair: cell reverse from the bf16 tape, then d h_{t-1} = dA_t @ Wh (split-K, fp32 atomics)
        ARCVAE_CUDA(cudaMemsetAsync(sc.dc, 0, (size_t)B * H * sizeof(float), st));
        // split-K factor of the d h GEMM (few output tiles, K = 4H): partials are STORED, the cell kernel adds them
        const long tiles = (long)cdiv(B, 128) * cdiv(H, 256), kblocks = G4 / 64;
        int S = 1;
        while (S < 4 && tiles * S * 2 <= 148 && kblocks / (S * 2) >= 8) S *= 2;
        const long pstride = (long)B * H;
        timing_begin(TIME_RECURRENCE, st);
        for (int t = T - 1; t >= 0; t--) {
          const long r0 = (long)t * B;
          const float* dh_ext = top ? nullptr : sc.dX + r0 * H;
          if (top && t == T - 1) {                              // d h_T = du[:, 0:H] (row pitch 2H) -> dense
            ARCVAE_CUDA(cudaMemcpy2DAsync(sc.dh_rec[1], (size_t)H * sizeof(float), sc.du, (size_t)H2 * sizeof(float),
                                          (size_t)H * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
            dh_ext = sc.dh_rec[1];
          }
          ARCVAE_TRY(lstm_cell_bwd_b(tp.gates_b[l] + r0 * G4, tp.c[l] + r0 * H, t > 0 ? tp.c[l] + (r0 - B) * H : nullptr,
                                     dh_ext, t == T - 1 ? nullptr : sc.dh_part, S, pstride, sc.dc, sc.dAb + r0 * G4, B, H, st));
          if (t > 0) {
            TcGemm q{};
            q.M = B; q.N = H; q.K = G4;
            q.A = sc.dAb + r0 * G4; q.lda = G4; q.a_mn = false;
            q.B = tp.Whb[l]; q.ldb = H; q.b_mn = true;
            q.C = sc.dh_part; q.ldc = H; q.accumulate = false; q.splitk = S; q.split_stride = S > 1 ? pstride : 0;
            q.rm = id; q.a_rows_total = B;
            ARCVAE_TRY(gemm_tc(q, st));
          }
        }
        timing_end(TIME_RECURRENCE, st);
      }
      if (fuse_dw) {
        ARCVAE_CUDA(cudaMemsetAsync(sc.segtmp, 0, (size_t)G4 * SCATTER_NW * sizeof(float), st));
        TcGemm q{};
        q.M = G4; q.N = H; q.K = (int)R;
        q.A = sc.dAb; q.lda = G4; q.a_mn = true; q.b_mn = true;
        q.accumulate = true; q.rm = id; q.a_rows_total = R;
        int ns = 0;
        q.seg[ns++] = {tp.hb[l], H, H, B, g->Wh[l], H};                       // dWh += dA[t]^T h[t-1]  (rows shifted by B)
        if (l > 0) q.seg[ns++] = {tp.hb[l - 1], H, H, 0, g->Wx[l], H};         // dWx += dA^T h_{l-1}
        q.seg[ns++] = {sc.onehot, SCATTER_NW, SCATTER_NW, 0, sc.segtmp, SCATTER_NW};   // dA^T onehot(x)
        q.nseg = ns;
        const long tiles = (long)cdiv(G4, 128) * ns;
        long sk = (2 * 148) / tiles;
        const long maxs = cdiv(R, 64) / 8;
        if (sk > maxs) sk = maxs;
        q.splitk = sk < 1 ? 1 : (int)sk;
        ARCVAE_TRY(gemm_tc(q, st));
        if (l > 0) {
          ARCVAE_TRY(rowsum_add(sc.segtmp, G4, SCATTER_NW, d->V, g->bias[l], st));
          if (dx_tf) {
            // d h_{l-1} = dA_l @ Wx_l, bf16 in the layout the cluster BPTT of layer l-1 reads (no fp32 [T*B,H] round trip)
            TcGemm q2{};
            q2.M = (int)R; q2.N = H; q2.K = G4;
            q2.A = sc.dAb; q2.lda = G4; q2.a_mn = false;
            q2.B = tp.Wxb[l]; q2.ldb = H; q2.b_mn = true;
            q2.Cb = sc.dXtf; q2.ldcb = H; q2.accumulate = false; q2.splitk = 1;
            q2.rm = id; q2.a_rows_total = R;
            q2.epi = TC_EPI_LSTM_DH; q2.lp_B = B;
            ARCVAE_TRY(gemm_tc(q2, st));
          } else {
            ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, G4, Mat{nullptr, sc.dAb, G4}, Mat{nullptr, tp.Wxb[l], H}, sc.dX, H,
                                nullptr, false, id, R, st));
          }
        } else {
          ARCVAE_TRY(transpose_f32(sc.segtmp, G4, SCATTER_NW, SCATTER_NW, sc.dtable0, st));   // -> [SCATTER_NW, 4H]
          ARCVAE_TRY(layer0_input_backward(*d, p, g, sc, st));
        }
        continue;
      }
      // dWh += dA[1:]^T @ h[:-1]
      if (T > 1) {
        long K = (long)(T - 1) * B;
        ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)K, Mat{nullptr, sc.dAb + (long)B * G4, G4}, Mat{nullptr, tp.hb[l], H},
                            g->Wh[l], H, nullptr, true, id, K, st));
      }
      if (l > 0) {
        ARCVAE_TRY(colsum_bf16(sc.dAb, R, G4, G4, g->bias[l], st));
        ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)R, Mat{nullptr, sc.dAb, G4}, Mat{nullptr, tp.hb[l - 1], H},
                            g->Wx[l], H, nullptr, true, id, R, st));
        ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, G4, Mat{nullptr, sc.dAb, G4}, Mat{nullptr, tp.Wxb[l], H}, sc.dX, H,
                            nullptr, false, id, R, st));
      } else {
        if (scatter_onehot_supported(G4, d->V, 0)) {
          ARCVAE_TRY(scatter_rows_onehot_tc(sc.dAb, tp.xT, R, G4, d->V, sc.onehot, sc.dtable0, nullptr, B, 0, nullptr, 0, nullptr, st));
        } else {
          ARCVAE_CUDA(cudaMemsetAsync(sc.dtable0, 0, (size_t)d->V * G4 * sizeof(float), st));
          ARCVAE_TRY(scatter_rows_by_token_bf16(sc.dAb, tp.xT, R, G4, d->V, sc.dtable0, st));
        }
        ARCVAE_TRY(layer0_input_backward(*d, p, g, sc, st));
      }
    }
    return 0;
  }

  // ---- BPTT, per-step paths, top layer first
  for (int l = d->NL - 1; l >= 0; l--) {
    ARCVAE_CUDA(cudaMemsetAsync(sc.dc, 0, (size_t)B * H * sizeof(float), st));
    const bool top = (l == d->NL - 1);
    timing_begin(TIME_RECURRENCE, st);
    for (int t = T - 1; t >= 0; t--) {
      float* g_t = tp.gates[l] + (long)t * B * G4;
      const float* c_t = tp.c[l] + (long)t * B * H;
      const float* dh_ext;
      if (top) {
        dh_ext = nullptr;  // only t = T-1 receives d h_T = du[:, 0:H] (strided)
        if (t == T - 1) {
          // copy du[:, 0:H] into a dense [B,H] buffer (dh_rec[1] is free at the first step)
          ARCVAE_CUDA(cudaMemcpy2DAsync(sc.dh_rec[1], (size_t)H * sizeof(float), sc.du, (size_t)H2 * sizeof(float),
                                        (size_t)H * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
          dh_ext = sc.dh_rec[1];
        }
      } else {
        dh_ext = sc.dX + (long)t * B * H;
      }
      const float* dh_rec = (t == T - 1) ? nullptr : sc.dh_rec[0];
      bf16* dAb_t = bf ? sc.dAb + (long)t * B * G4 : nullptr;
      ARCVAE_TRY(lstm_cell_bwd(g_t, c_t, t > 0 ? c_t - (long)B * H : nullptr, dh_ext, dh_rec, sc.dc, dAb_t, B, H, st));
      if (t > 0) {
        // d h_{t-1} (recurrent part) = dA_t @ Wh
        ARCVAE_TRY(gemm_any(precision, 0, 0, B, H, G4, Mat{g_t, dAb_t, G4}, Mat{p->Wh[l], tp.Whb[l], H}, sc.dh_rec[0], H,
                            nullptr, false, id, B, st));
      }
    }
    timing_end(TIME_RECURRENCE, st);
    float* dA = tp.gates[l];
    // dWh += dA[1:]^T @ h[:-1]
    if (T > 1) {
      long K = (long)(T - 1) * B;
      ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)K, Mat{dA + (long)B * G4, bf ? sc.dAb + (long)B * G4 : nullptr, G4},
                          Mat{tp.h[l], tp.hb[l], H}, g->Wh[l], H, nullptr, true, id, K, st));
    }
    if (l > 0) {
      ARCVAE_TRY(colsum(dA, R, G4, G4, g->bias[l], st));
      // dWx += dA^T @ h_{l-1} ; dX = dA @ Wx
      ARCVAE_TRY(gemm_any(precision, 1, 0, G4, H, (int)R, Mat{dA, sc.dAb, G4}, Mat{tp.h[l - 1], tp.hb[l - 1], H}, g->Wx[l], H,
                          nullptr, true, id, R, st));
      ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, G4, Mat{dA, sc.dAb, G4}, Mat{p->Wx[l], tp.Wxb[l], H}, sc.dX, H, nullptr,
                          false, id, R, st));
    } else {
      ARCVAE_CUDA(cudaMemsetAsync(sc.dtable0, 0, (size_t)d->V * G4 * sizeof(float), st));
      ARCVAE_TRY(scatter_rows_by_token(dA, tp.xT, R, G4, d->V, sc.dtable0, nullptr, B, d->C, nullptr, st));
      ARCVAE_TRY(layer0_input_backward(*d, p, g, sc, st));
    }
  }
  return 0;
}

extern "C" int arcvae_recurrence_is_persistent(int H) { return lstm_cluster_supported(H) ? 1 : 0; }

// raised by the cluster kernels' bounded waits (0 = healthy).  Synchronises the stream.  Since ABI v2 the flag is the
// sticky per-device one (arcvae_device_error_read); the tape arguments are kept for source compatibility.
extern "C" int arcvae_encoder_check(const arcvae_dims* d, int B, int T, void* tape, size_t tape_bytes, int precision,
                                    void* stream) {
  ARCVAE_TRY(check_dims(d));
  (void)B; (void)T; (void)tape; (void)tape_bytes; (void)precision;
  int flag = 0;
  ARCVAE_TRY(arcvae_device_error_read(&flag, stream));
  ARCVAE_REQUIRE(flag == 0, "cluster recurrence kernel reported a barrier time-out (sticky device flag; "
                            "arcvae_device_error_clear() resets it)");
  return 0;
}
