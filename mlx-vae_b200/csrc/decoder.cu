// decoder.cu — MLXAutoregressiveDecoder.__call__ (models/decoder.py:113-190), its reverse pass, and the sampler
// loop of MLXAutoregressiveDecoderSampling.generate_with_temperature (models/decoder_sampling.py:48-128).
//
// Reference semantics that shape the design:
//   * every position calls the LSTM layers on a length-1 sequence WITHOUT state, so
//       logits[:,t] = fc_out(cell(cell([Emb[tok_{t-1}] ; cond] Wx0^T + b0) Wx1^T + b1)),  cell(a) = s(a_o)*tanh(s(a_i)*tanh(a_g))
//     (forget gate, Wh, z, z_to_hidden, condition_to_hidden never reach the output);
//   * the token fed at position t is x[:,t-1] when the host coin of position t-1 came up "teacher forcing",
//     else argmax(logits[:,t-1]) — one coin per position for the whole batch.
// So positions whose input is known are time-parallel.  Positions are grouped in LEVELS (level 0: input known;
// level k: input is the greedy output of a level k-1 position); each level is one batched pass over the rows of
// its timesteps (time-major rows r = t*B + b, reached through a RowMap).
//   layer 0  : gather of table = Emb @ Wx0[:, :E]^T + b0 (rows i,g,o only) + cond (x) Wx0[:, E:]  — no GEMM needed
//   layer l>0: GEMM [rows,H] x [H,3H] on the compacted (i,g,o) weight rows
#include <cstdlib>
#include <vector>

#include "kernels.cuh"

namespace arcvae {

struct DecPrep {
  float* Wxc[ARCVAE_MAX_LAYERS];  // [3H, D_l]
  float* bc[ARCVAE_MAX_LAYERS];   // [3H]
  float* table;                   // [V,3H]
  float* wc;                      // [3H,C]
  __nv_bfloat16* Wxcb[ARCVAE_MAX_LAYERS];  // bf16 [3H,H], l >= 1
  __nv_bfloat16* Woutb;                    // bf16 [V,H]
  // fused path: tile-permuted compact gates (row p = j*192 + gi*64 + u), so that one 192-wide GEMM tile holds
  // i, g, o of the same 64 hidden units and the cell can run in the GEMM epilogue
  float* Wxp[ARCVAE_MAX_LAYERS];           // fp32 [3H,H], l >= 1
  float* bp[ARCVAE_MAX_LAYERS];            // fp32 [3H]
  __nv_bfloat16* Wxpb[ARCVAE_MAX_LAYERS];  // bf16 [3H,H]
};

static bool dec_fused_ok(const arcvae_dims& d, int precision, int B, bool row_mapped) {
  if (precision != ARCVAE_PREC_BF16 || d.H % 64 != 0 || d.NL < 1 || d.V + 2 * d.C > SCATTER_NW) return false;
  if (row_mapped && (B % 128) != 0) return false;
  return std::getenv("ARCVAE_NO_FUSED_DEC") == nullptr;
}

static void dec_prep_layout(const arcvae_dims& d, Arena& a, DecPrep* p) {
  for (int l = 0; l < ARCVAE_MAX_LAYERS; l++) p->Wxcb[l] = nullptr;
  for (int l = 1; l < d.NL; l++) p->Wxcb[l] = a.take<__nv_bfloat16>((size_t)3 * d.H * d.H);
  p->Woutb = a.take<__nv_bfloat16>((size_t)d.V * d.H);
  for (int l = 0; l < ARCVAE_MAX_LAYERS; l++) { p->Wxp[l] = nullptr; p->bp[l] = nullptr; p->Wxpb[l] = nullptr; }
  for (int l = 1; l < d.NL; l++) {
    p->Wxp[l] = a.take<float>((size_t)3 * d.H * d.H);
    p->bp[l] = a.take<float>((size_t)3 * d.H);
    p->Wxpb[l] = a.take<__nv_bfloat16>((size_t)3 * d.H * d.H);
  }
  for (int l = 0; l < d.NL; l++) {
    int D = (l == 0) ? d.E + d.C : d.H;
    p->Wxc[l] = a.take<float>((size_t)3 * d.H * D);
    p->bc[l] = a.take<float>((size_t)3 * d.H);
  }
  p->table = a.take<float>((size_t)d.V * 3 * d.H);
  p->wc = a.take<float>((size_t)3 * d.H * d.C);
}

static int dec_prepare(const arcvae_dims& d, const arcvae_decoder_params* p, const DecPrep& pr, int precision,
                       cudaStream_t st) {
  RowMap id{nullptr, 1};
  const int H = d.H, H3 = 3 * d.H;
  for (int l = 0; l < d.NL; l++) {
    int D = (l == 0) ? d.E + d.C : H;
    ARCVAE_TRY(compact_gates(p->Wx[l], H, D, pr.Wxc[l], st));
    ARCVAE_TRY(compact_gates(p->bias[l], H, 1, pr.bc[l], st));
  }
  ARCVAE_TRY(gemm_f32(0, 1, d.V, H3, d.E, p->embedding, d.E, pr.Wxc[0], d.E + d.C, pr.table, H3, pr.bc[0], false, id, 1, st));
  ARCVAE_CUDA(cudaMemcpy2DAsync(pr.wc, (size_t)d.C * sizeof(float), pr.Wxc[0] + d.E, (size_t)(d.E + d.C) * sizeof(float),
                                (size_t)d.C * sizeof(float), H3, cudaMemcpyDeviceToDevice, st));
  if (precision == ARCVAE_PREC_BF16) {
    for (int l = 1; l < d.NL; l++) ARCVAE_TRY(f32_to_bf16(pr.Wxc[l], pr.Wxcb[l], (long)H3 * H, st));
    ARCVAE_TRY(f32_to_bf16(p->fc_out_w, pr.Woutb, (long)d.V * H, st));
    if (d.H % 64 == 0) {
      for (int l = 1; l < d.NL; l++) {
        ARCVAE_TRY(compact_perm_gates(p->Wx[l], H, H, pr.Wxp[l], st));
        ARCVAE_TRY(compact_perm_gates(p->bias[l], H, 1, pr.bp[l], st));
        ARCVAE_TRY(f32_to_bf16(pr.Wxp[l], pr.Wxpb[l], (long)H3 * H, st));
      }
    }
  }
  return 0;
}

struct DecTape {
  DecPrep prep;
  int32_t* in_tok;                   // [T,B]
  float* hd[ARCVAE_MAX_LAYERS];      // [R,H]
  float* G[ARCVAE_MAX_LAYERS];       // [R,3H], l >= 1
  int* tlists;                       // [2T] level lists + feedback lists
  uint8_t* mask;                     // [T]
  __nv_bfloat16* hdb[ARCVAE_MAX_LAYERS];   // bf16 [R,H] copies (tensor-core operands)
  __nv_bfloat16* gates_b[ARCVAE_MAX_LAYERS];   // fused path: bf16 [R,3H] activated gates, tile-permuted
  // fc_out with the cross-entropy in its epilogue (arcvae_decoder_forward_ce)
  __nv_bfloat16* dlb;                // bf16 [R,Vp] d logits
  int32_t* xT;                       // [T,B] targets, time-major
  uint8_t* fb;                       // [T] position t feeds argmax(logits_t) to t+1
};

static size_t dec_tape_layout(const arcvae_dims& d, int B, int T, void* base, size_t cap, DecTape* t) {
  Arena a(base, cap);
  DecTape tt;
  size_t R = (size_t)T * B;
  dec_prep_layout(d, a, &tt.prep);
  tt.in_tok = a.take<int32_t>(R);
  for (int l = 0; l < d.NL; l++) {
    tt.hd[l] = a.take<float>(R * d.H);
    tt.G[l] = (l >= 1) ? a.take<float>(R * 3 * d.H) : nullptr;
  }
  tt.tlists = a.take<int>((size_t)2 * T + 2);
  tt.mask = a.take<uint8_t>((size_t)T + 16);
  for (int l = 0; l < d.NL; l++) tt.hdb[l] = a.take<__nv_bfloat16>(R * d.H);
  for (int l = 0; l < d.NL; l++) tt.gates_b[l] = a.take<__nv_bfloat16>(R * 3 * d.H);   // layer 0: unused when its gates are recomputed
  tt.dlb = a.take<__nv_bfloat16>(R * (size_t)((d.V + 7) / 8 * 8));
  tt.xT = a.take<int32_t>(R);
  tt.fb = a.take<uint8_t>((size_t)T + 16);
  if (t) *t = tt;
  return align_up(a.off, 256);
}

struct DecScratch {
  float* dh[2];                       // [R,H]
  float* dG0;                         // [R,3H]
  float* dtable;                      // [max(V, SCATTER_NW),3H]  (rows >= V: scratch of the one-hot GEMM)
  float* dWxc[ARCVAE_MAX_LAYERS];     // compact weight grads
  float* dbc[ARCVAE_MAX_LAYERS];
  float* dwc;                         // [3H,C]
  __nv_bfloat16* dlb;                 // bf16 [R,Vp] copy of dlogits, row pitch Vp = V rounded up to 8 (TMA needs 16-byte pitches)
  __nv_bfloat16* dGb;                 // bf16 [R,3H] copy of the pre-activation gradients
  __nv_bfloat16* dGb2;                // fused path: second [R,3H] buffer (ping-pong between layers, layer-0 dG)
  __nv_bfloat16* onehot;              // fused path: [R,SCATTER_NW] one-hot of the fed tokens + cond hi/lo columns
  float* dtable_tmp;                  // fused path: [SCATTER_NW,3H] table gradient in tile-permuted column order
  float* segtmp;                      // fused path: [3H,SCATTER_NW] one-hot segment of the weight-gradient GEMM (bias row sums)
};

static size_t dec_scratch_layout(const arcvae_dims& d, int B, int T, void* base, size_t cap, DecScratch* s) {
  Arena a(base, cap);
  DecScratch ss;
  size_t R = (size_t)T * B;
  ss.dh[0] = a.take<float>(R * d.H);
  ss.dh[1] = a.take<float>(R * d.H);
  ss.dG0 = a.take<float>(R * 3 * d.H);
  ss.dtable = a.take<float>((size_t)(d.V > SCATTER_NW ? d.V : SCATTER_NW) * 3 * d.H);
  for (int l = 0; l < d.NL; l++) {
    int D = (l == 0) ? d.E + d.C : d.H;
    ss.dWxc[l] = a.take<float>((size_t)3 * d.H * D);
    ss.dbc[l] = a.take<float>((size_t)3 * d.H);
  }
  ss.dwc = a.take<float>((size_t)3 * d.H * d.C);
  ss.dlb = a.take<__nv_bfloat16>(R * (size_t)((d.V + 7) / 8 * 8));
  ss.dGb = a.take<__nv_bfloat16>(R * 3 * d.H);
  ss.dGb2 = a.take<__nv_bfloat16>(R * 3 * d.H);
  ss.onehot = a.take<__nv_bfloat16>(R * SCATTER_NW);
  ss.dtable_tmp = a.take<float>((size_t)SCATTER_NW * 3 * d.H);
  ss.segtmp = a.take<float>((size_t)3 * d.H * SCATTER_NW);
  if (s) *s = ss;
  return align_up(a.off, 256);
}

// in_tok[t,b] = 0 for t == 0; x[b,t-1] where the coin of t-1 is teacher forcing; 0 placeholder otherwise
__global__ void k_init_dec_inputs(const int32_t* __restrict__ target, const uint8_t* __restrict__ mask, int B, int T,
                                  int32_t* __restrict__ in_tok) {
  long total = (long)T * B;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int t = (int)(i / B), b = (int)(i - (long)t * B);
    int v = 0;
    if (t > 0 && target != nullptr && mask[t - 1]) v = target[(long)b * T + (t - 1)];
    in_tok[i] = v;
  }
}

struct IntChunk { int v[512]; };
__global__ void k_write_ints(int* __restrict__ dst, IntChunk c, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = c.v[i];
}
// small host arrays -> device through kernel parameters (2 KB per launch)
static int upload_ints(int* dst, const int* src, int n, cudaStream_t st) {
  for (int off = 0; off < n; off += 512) {
    IntChunk c;
    const int m = n - off < 512 ? n - off : 512;
    for (int i = 0; i < m; i++) c.v[i] = src[off + i];
    k_write_ints<<<1, 128, 0, st>>>(dst + off, c, m);
    ARCVAE_LAUNCHED();
  }
  return 0;
}

struct CeArgs {                        // fc_out + cross-entropy + d logits + greedy feedback in one GEMM epilogue
  const int32_t* xT; const uint8_t* fb; int32_t* in_tok; __nv_bfloat16* dlb; int Vp; double* ce_sum; float scale;
};

// one batched pass of the decoder stack over the rows selected by `rm`
static int dec_stack_forward(const arcvae_dims& d, const arcvae_decoder_params* p, const DecPrep& pr, const float* cond,
                             const int32_t* in_tok, int B, int nrows, RowMap rm, long rows_total, float* const* hd,
                             __nv_bfloat16* const* hdb, float* const* G, __nv_bfloat16* const* gates_b, float* logits,
                             int precision, bool fused, cudaStream_t st, const CeArgs* ce = nullptr) {
  const int H = d.H, H3 = 3 * d.H;
  const bool bf = precision == ARCVAE_PREC_BF16;
  ARCVAE_TRY(dec_cell0_fwd(pr.table, pr.wc, in_tok, cond, B, d.C, H, d.V, nrows, rm, fused ? nullptr : hd[0],
                           bf ? hdb[0] : nullptr, (fused && !dec_cell0_recompute_ok(H, d.V, d.C)) ? gates_b[0] : nullptr, st));
  for (int l = 1; l < d.NL; l++) {
    if (fused) {
      // GEMM + zero-state cell in the epilogue: h_{l-1} @ Wxp_l^T + bp_l -> (i,g,o) -> h_l ; gates saved as bf16
      TcGemm g{};
      g.M = nrows; g.N = H3; g.K = H;
      g.A = hdb[l - 1]; g.lda = H; g.a_mn = false;
      g.B = pr.Wxpb[l]; g.ldb = H; g.b_mn = false;
      g.bias = pr.bp[l]; g.accumulate = false; g.splitk = 1; g.rm = rm; g.a_rows_total = rows_total;
      g.epi = TC_EPI_DEC_CELL_FWD; g.gates_b = gates_b[l]; g.hb_out = hdb[l]; g.Hh = H;
      ARCVAE_TRY(gemm_tc(g, st));
    } else {
      ARCVAE_TRY(gemm_any(precision, 0, 1, nrows, H3, H, Mat{hd[l - 1], bf ? hdb[l - 1] : nullptr, H},
                          Mat{pr.Wxc[l], pr.Wxcb[l], H}, G[l], H3, pr.bc[l], false, rm, rows_total, st));
      ARCVAE_TRY(dec_cell_fwd(G[l], hd[l], bf ? hdb[l] : nullptr, H, nrows, rm, st));
    }
  }
  if (ce != nullptr) {
    TcGemm g{};
    g.M = nrows; g.N = d.V; g.K = H;
    g.A = hdb[d.NL - 1]; g.lda = H; g.a_mn = false;
    g.B = pr.Woutb; g.ldb = H; g.b_mn = false;
    g.bias = p->fc_out_b; g.accumulate = false; g.splitk = 1; g.rm = rm; g.a_rows_total = rows_total;
    g.epi = TC_EPI_CE;
    g.ce_target = ce->xT; g.ce_fb = ce->fb; g.ce_tok = ce->in_tok; g.ce_B = B; g.ce_dl = ce->dlb; g.ce_ldl = ce->Vp;
    g.ce_sum = ce->ce_sum; g.ce_scale = ce->scale;
    return gemm_tc(g, st);
  }
  ARCVAE_TRY(gemm_any(precision, 0, 1, nrows, d.V, H, Mat{fused ? nullptr : hd[d.NL - 1], bf ? hdb[d.NL - 1] : nullptr, H},
                      Mat{p->fc_out_w, pr.Woutb, H}, logits, d.V, p->fc_out_b, false, rm, rows_total, st));
  return 0;
}

static int check_dims_dec(const arcvae_dims* d) {
  ARCVAE_REQUIRE(d != nullptr, "dims");
  ARCVAE_REQUIRE(d->V > 0 && d->E > 0 && d->H > 0 && d->C > 0, "positive dims");
  ARCVAE_REQUIRE(d->NL >= 1 && d->NL <= ARCVAE_MAX_LAYERS, "num_layers in [1, 8]");
  return 0;
}

}  // namespace arcvae

using namespace arcvae;

extern "C" size_t arcvae_decoder_tape_bytes(const arcvae_dims* d, int B, int T) {
  return d ? dec_tape_layout(*d, B, T, nullptr, 0, nullptr) : 0;
}
extern "C" size_t arcvae_decoder_scratch_bytes(const arcvae_dims* d, int B, int T) {
  return d ? dec_scratch_layout(*d, B, T, nullptr, 0, nullptr) : 0;
}

// ce_sum != nullptr: fc_out runs with the cross-entropy in its epilogue (no logits tensor); see arcvae_decoder_forward_ce
static int decoder_forward_impl(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond,
                                const int32_t* target, const uint8_t* tf_mask_host, int B, int T, float* logits_tm,
                                int32_t* dec_inputs_tm, void* tape, size_t tape_bytes, int precision, float ce_scale,
                                double* ce_sum, void* stream) {
  ARCVAE_TRY(check_dims_dec(d));
  ARCVAE_REQUIRE(B > 0 && T > 0, "empty batch / sequence");
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(logits_tm != nullptr || ce_sum != nullptr, "logits output");
  cudaStream_t st = (cudaStream_t)stream;
  DecTape tp;
  size_t need = dec_tape_layout(*d, B, T, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "decoder tape too small");

  // host-side schedule: coin[t] (decoder.py:180), levels, per-level timestep lists and feedback lists
  std::vector<uint8_t> coin(T, 0);
  for (int t = 0; t < T; t++) coin[t] = (target != nullptr && tf_mask_host != nullptr && tf_mask_host[t]) ? 1 : 0;
  std::vector<int> level(T, 0);
  int maxlevel = 0;
  for (int t = 1; t < T; t++) {
    level[t] = coin[t - 1] ? 0 : level[t - 1] + 1;
    if (level[t] > maxlevel) maxlevel = level[t];
  }
  std::vector<int> lists;            // concatenated: per level [timesteps..., feedback timesteps...]
  std::vector<int> off_t(maxlevel + 2), n_t(maxlevel + 1), off_f(maxlevel + 1), n_f(maxlevel + 1);
  lists.reserve(2 * T + 2);
  for (int lv = 0; lv <= maxlevel; lv++) {
    off_t[lv] = (int)lists.size();
    for (int t = 0; t < T; t++)
      if (level[t] == lv) lists.push_back(t);
    n_t[lv] = (int)lists.size() - off_t[lv];
    off_f[lv] = (int)lists.size();
    for (int t = 0; t < T; t++)
      if (level[t] == lv && !coin[t] && t + 1 < T) lists.push_back(t);
    n_f[lv] = (int)lists.size() - off_f[lv];
  }
  ARCVAE_REQUIRE((int)lists.size() <= 2 * T + 2, "internal: list overflow");
  // the schedule travels as KERNEL ARGUMENTS (copied at launch): no host buffer has to outlive this call and, unlike a
  // pageable cudaMemcpyAsync + synchronize, the host never waits for the stream, so a whole step can be enqueued ahead
  ARCVAE_TRY(upload_ints(tp.tlists, lists.data(), (int)lists.size(), st));
  {
    std::vector<int> packed((T + 3) / 4, 0);
    for (int t = 0; t < T; t++) packed[t >> 2] |= (int)coin[t] << (8 * (t & 3));
    ARCVAE_TRY(upload_ints(reinterpret_cast<int*>(tp.mask), packed.data(), (int)packed.size(), st));
  }

  ARCVAE_TRY(dec_prepare(*d, p, tp.prep, precision, st));
  {
    long total = (long)T * B;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    k_init_dec_inputs<<<grid, 256, 0, st>>>(target, tp.mask, B, T, tp.in_tok);
    ARCVAE_LAUNCHED();
  }
  CeArgs ce{};
  if (ce_sum != nullptr) {
    ARCVAE_REQUIRE(target != nullptr && dec_fused_ok(*d, precision, B, true) && d->V <= 128,
                   "fused cross-entropy needs targets, the fused bf16 decoder path (H % 64 == 0, B % 128 == 0) and V <= 128");
    std::vector<int> packed((T + 3) / 4, 0);                  // fb[t]: position t feeds argmax(logits_t) to position t+1
    for (int t = 0; t + 1 < T; t++)
      if (!coin[t]) packed[t >> 2] |= 1 << (8 * (t & 3));
    ARCVAE_TRY(upload_ints(reinterpret_cast<int*>(tp.fb), packed.data(), (int)packed.size(), st));
    ARCVAE_TRY(transpose_tokens(target, B, T, tp.xT, st));
    ce = CeArgs{tp.xT, tp.fb, tp.in_tok, tp.dlb, (d->V + 7) / 8 * 8, ce_sum, ce_scale};
  }
  for (int lv = 0; lv <= maxlevel; lv++) {
    RowMap rm{tp.tlists + off_t[lv], B};
    int nrows = n_t[lv] * B;
    ARCVAE_TRY(dec_stack_forward(*d, p, tp.prep, cond, tp.in_tok, B, nrows, rm, (long)T * B, tp.hd, tp.hdb, tp.G,
                                 tp.gates_b, logits_tm, precision, dec_fused_ok(*d, precision, B, true), st,
                                 ce_sum != nullptr ? &ce : nullptr));
    if (ce_sum == nullptr) ARCVAE_TRY(argmax_feedback(logits_tm, tp.tlists + off_f[lv], n_f[lv], B, d->V, tp.in_tok, st));
  }
  if (dec_inputs_tm != nullptr)
    ARCVAE_CUDA(cudaMemcpyAsync(dec_inputs_tm, tp.in_tok, (size_t)T * B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int arcvae_decoder_forward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond,
                                      const int32_t* target, const uint8_t* tf_mask_host, int B, int T,
                                      float* logits_tm, int32_t* dec_inputs_tm, void* tape, size_t tape_bytes,
                                      int precision, void* stream) {
  ARCVAE_REQUIRE(logits_tm != nullptr, "logits output");
  return decoder_forward_impl(d, p, cond, target, tf_mask_host, B, T, logits_tm, dec_inputs_tm, tape, tape_bytes, precision,
                              0.f, nullptr, stream);
}

extern "C" int arcvae_decoder_ce_supported(const arcvae_dims* d, int B, int precision) {
  return (d != nullptr && dec_fused_ok(*d, precision, B, true) && d->V <= 128) ? 1 : 0;
}

extern "C" int arcvae_decoder_forward_ce(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond,
                                         const int32_t* target, const uint8_t* tf_mask_host, int B, int T, float ce_scale,
                                         double* ce_sum, int32_t* dec_inputs_tm, void* tape, size_t tape_bytes,
                                         int precision, void* stream) {
  ARCVAE_REQUIRE(ce_sum != nullptr && ce_scale > 0.f, "ce_sum / ce_scale");
  return decoder_forward_impl(d, p, cond, target, tf_mask_host, B, T, nullptr, dec_inputs_tm, tape, tape_bytes, precision,
                              ce_scale, ce_sum, stream);
}

extern "C" int arcvae_decoder_backward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond, int B,
                                       int T, float* dlogits_tm, void* tape, size_t tape_bytes,
                                       const arcvae_decoder_params* g, void* scratch, size_t scratch_bytes,
                                       int precision, void* stream) {
  ARCVAE_TRY(check_dims_dec(d));
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(g != nullptr, "grad pointers");
  cudaStream_t st = (cudaStream_t)stream;
  DecTape tp;
  size_t need = dec_tape_layout(*d, B, T, tape, tape_bytes, &tp);
  ARCVAE_REQUIRE(tape != nullptr && need <= tape_bytes, "decoder tape too small");
  DecScratch sc;
  size_t need_s = dec_scratch_layout(*d, B, T, scratch, scratch_bytes, &sc);
  ARCVAE_REQUIRE(scratch != nullptr && need_s <= scratch_bytes, "decoder scratch too small");
  const int H = d->H, H3 = 3 * d->H, V = d->V, E = d->E, C = d->C;
  const long R = (long)T * B;
  RowMap id{nullptr, 1};
  const int top = d->NL - 1;

  const bool bf = precision == ARCVAE_PREC_BF16;
  const bool fused = dec_fused_ok(*d, precision, B, true);
  const int Vp = (V + 7) / 8 * 8;            // row pitch of the bf16 copy of dlogits
  // dlogits_tm == NULL: the forward was arcvae_decoder_forward_ce and the bf16 d logits already sit in the tape
  ARCVAE_REQUIRE(dlogits_tm != nullptr || (bf && fused), "dlogits (or a preceding arcvae_decoder_forward_ce)");
  if (dlogits_tm == nullptr) sc.dlb = tp.dlb;
  else if (bf) ARCVAE_TRY(f32_to_bf16_pitched(dlogits_tm, R, V, sc.dlb, Vp, st));
  // fc_out: logits = h_top @ Wout^T + b
  const bool fuse_out = fused && H <= 256 && (H % 64) == 0 && V <= 128 && std::getenv("ARCVAE_NO_FUSED_DWOUT") == nullptr;
  if (fuse_out) {
    // dWout = dlogits^T h_top and the bias gradient (row sums of dlogits^T onehot: every one-hot row sums to 1) in ONE
    // pass over the bf16 d logits
    ARCVAE_TRY(build_onehot(tp.in_tok, R, V, cond, B, C, sc.onehot, st));
    ARCVAE_CUDA(cudaMemsetAsync(sc.segtmp, 0, (size_t)128 * SCATTER_NW * sizeof(float), st));
    TcGemm q{};
    q.M = V; q.N = H; q.K = (int)R;
    q.A = sc.dlb; q.lda = Vp; q.a_mn = true; q.b_mn = true;
    q.accumulate = true; q.rm = id; q.a_rows_total = R;
    q.nseg = 2;
    q.seg[0] = {tp.hdb[top], H, H, 0, g->fc_out_w, H};
    q.seg[1] = {sc.onehot, SCATTER_NW, SCATTER_NW, 0, sc.segtmp, SCATTER_NW};
    long sk = (2 * 148) / 2;
    const long maxs = cdiv(R, 64) / 8;
    if (sk > maxs) sk = maxs;
    q.splitk = sk < 1 ? 1 : (int)sk;
    ARCVAE_TRY(gemm_tc(q, st));
    ARCVAE_TRY(rowsum_add(sc.segtmp, V, SCATTER_NW, V, g->fc_out_b, st));
  } else {
    ARCVAE_TRY(gemm_any(precision, 1, 0, V, H, (int)R, Mat{dlogits_tm, bf ? sc.dlb : nullptr, bf && fused ? Vp : V},
                        Mat{fused ? nullptr : tp.hd[top], bf ? tp.hdb[top] : nullptr, H}, g->fc_out_w, H, nullptr, true, id, R, st));
    if (dlogits_tm != nullptr) ARCVAE_TRY(colsum(dlogits_tm, R, V, V, g->fc_out_b, st));
    else ARCVAE_TRY(colsum_bf16(sc.dlb, R, V, Vp, g->fc_out_b, st));
  }
  float* dh = sc.dh[0];
  float* dh_next = sc.dh[1];

  if (fused) {
    // every "d h = dG_upper @ W" GEMM runs the cell-backward of the layer below in its epilogue; only bf16 dG is stored
    auto fused_gemm = [&](const __nv_bfloat16* A, int K, int lda, const __nv_bfloat16* Bw, int layer_below,
                          __nv_bfloat16* out) -> int {
      TcGemm q{};
      q.M = (int)R; q.N = H; q.K = K;
      q.A = A; q.lda = lda; q.a_mn = false;
      q.B = Bw; q.ldb = H; q.b_mn = true;                       // B[k*H + n]: row-major [K,H]
      q.accumulate = false; q.splitk = 1; q.rm = id; q.a_rows_total = R; q.Hh = H; q.dg_out = out;
      if (layer_below == 0 && dec_cell0_recompute_ok(H, V, C)) {
        // layer 0 has no gate tape: plain GEMM to bf16 d h_0 (in the idle fp32 d h scratch), then the cell reverse with
        // the gates recomputed from the token table
        __nv_bfloat16* dh0b = reinterpret_cast<__nv_bfloat16*>(sc.dh[0]);
        q.Cb = dh0b; q.ldcb = H; q.dg_out = nullptr; q.epi = TC_EPI_PLAIN;
        ARCVAE_TRY(gemm_tc(q, st));
        return dec_cell0_bwd_recompute(tp.prep.table, tp.prep.wc, tp.in_tok, cond, B, C, H, V, R, dh0b, out, st);
      }
      q.epi = TC_EPI_DEC_CELL_BWD; q.gates_b = tp.gates_b[layer_below];
      return gemm_tc(q, st);
    };
    // one-hot of the fed tokens (+ cond hi/lo columns): operand of the layer-0 table scatter and, through its row sums,
    // of the bias gradients of the upper layers
    if (!fuse_out) ARCVAE_TRY(build_onehot(tp.in_tok, R, V, cond, B, C, sc.onehot, st));
    const bool fuse_dw = H <= 256;
    __nv_bfloat16* cur = sc.dGb;
    __nv_bfloat16* nxt = sc.dGb2;
    ARCVAE_TRY(fused_gemm(sc.dlb, V, Vp, tp.prep.Woutb, top, cur));   // d h_top = dlogits @ Wout, then cell backward of `top`
    for (int l = top; l >= 1; l--) {
      // cur = dG_l (tile-permuted compact layout)
      ARCVAE_CUDA(cudaMemsetAsync(sc.dWxc[l], 0, (size_t)H3 * H * sizeof(float), st));
      ARCVAE_CUDA(cudaMemsetAsync(sc.dbc[l], 0, (size_t)H3 * sizeof(float), st));
      if (fuse_dw) {
        // dWx_l = dG^T h_{l-1} and the bias gradient (row sums of dG^T onehot) in ONE pass over dG
        ARCVAE_CUDA(cudaMemsetAsync(sc.segtmp, 0, (size_t)H3 * SCATTER_NW * sizeof(float), st));
        TcGemm q{};
        q.M = H3; q.N = H; q.K = (int)R;
        q.A = cur; q.lda = H3; q.a_mn = true; q.b_mn = true;
        q.accumulate = true; q.rm = id; q.a_rows_total = R;
        q.nseg = 2;
        q.seg[0] = {tp.hdb[l - 1], H, H, 0, sc.dWxc[l], H};
        q.seg[1] = {sc.onehot, SCATTER_NW, SCATTER_NW, 0, sc.segtmp, SCATTER_NW};
        const long tiles = (long)cdiv(H3, 128) * 2;
        long sk = (2 * 148) / tiles;
        const long maxs = cdiv(R, 64) / 8;
        if (sk > maxs) sk = maxs;
        q.splitk = sk < 1 ? 1 : (int)sk;
        ARCVAE_TRY(gemm_tc(q, st));
        ARCVAE_TRY(rowsum_add(sc.segtmp, H3, SCATTER_NW, V, sc.dbc[l], st));
      } else {
        ARCVAE_TRY(gemm_any(precision, 1, 0, H3, H, (int)R, Mat{nullptr, cur, H3}, Mat{nullptr, tp.hdb[l - 1], H}, sc.dWxc[l], H,
                            nullptr, true, id, R, st));
        ARCVAE_TRY(colsum_bf16(cur, R, H3, H3, sc.dbc[l], st));
      }
      ARCVAE_TRY(fused_gemm(cur, H3, H3, tp.prep.Wxpb[l], l - 1, nxt));
      ARCVAE_TRY(expand_perm_gates_add(sc.dWxc[l], H, H, g->Wx[l], st));
      ARCVAE_TRY(expand_perm_gates_add(sc.dbc[l], H, 1, g->bias[l], st));
      __nv_bfloat16* tmp = cur; cur = nxt; nxt = tmp;
    }
    // cur = dG_0 in the tile-permuted compact layout; the table gradient is un-permuted after the scatter
    ARCVAE_CUDA(cudaMemsetAsync(sc.dwc, 0, (size_t)H3 * C * sizeof(float), st));
    ARCVAE_CUDA(cudaMemsetAsync(sc.dWxc[0], 0, (size_t)H3 * (E + C) * sizeof(float), st));
    ARCVAE_CUDA(cudaMemsetAsync(sc.dbc[0], 0, (size_t)H3 * sizeof(float), st));
    ARCVAE_REQUIRE(scatter_onehot_supported(H3, V, C), "fused decoder backward: V + 2C <= 128");
    ARCVAE_TRY(scatter_rows_onehot_tc(cur, nullptr, R, H3, V, sc.onehot, sc.dtable, cond, B, C, sc.dwc, H, sc.dtable_tmp, st));
  } else {
  ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, V, Mat{dlogits_tm, bf ? sc.dlb : nullptr, V},
                      Mat{p->fc_out_w, tp.prep.Woutb, H}, dh, H, nullptr, false, id, R, st));

  for (int l = top; l >= 1; l--) {
    ARCVAE_TRY(dec_cell_bwd(tp.G[l], dh, bf ? sc.dGb : nullptr, H, R, st));  // G[l] <- dG
    ARCVAE_CUDA(cudaMemsetAsync(sc.dWxc[l], 0, (size_t)H3 * H * sizeof(float), st));
    ARCVAE_CUDA(cudaMemsetAsync(sc.dbc[l], 0, (size_t)H3 * sizeof(float), st));
    ARCVAE_TRY(gemm_any(precision, 1, 0, H3, H, (int)R, Mat{tp.G[l], bf ? sc.dGb : nullptr, H3},
                        Mat{tp.hd[l - 1], bf ? tp.hdb[l - 1] : nullptr, H}, sc.dWxc[l], H, nullptr, true, id, R, st));
    ARCVAE_TRY(colsum(tp.G[l], R, H3, H3, sc.dbc[l], st));
    ARCVAE_TRY(gemm_any(precision, 0, 0, (int)R, H, H3, Mat{tp.G[l], bf ? sc.dGb : nullptr, H3},
                        Mat{tp.prep.Wxc[l], tp.prep.Wxcb[l], H}, dh_next, H, nullptr, false, id, R, st));
    ARCVAE_TRY(expand_gates_add(sc.dWxc[l], H, H, g->Wx[l], st));
    ARCVAE_TRY(expand_gates_add(sc.dbc[l], H, 1, g->bias[l], st));
    float* tmp = dh; dh = dh_next; dh_next = tmp;
  }
  // layer 0: gates recomputed from the table; a0 = table[tok] + cond @ wc^T
  ARCVAE_TRY(dec_cell0_bwd(tp.prep.table, tp.prep.wc, tp.in_tok, cond, B, C, H, R, dh, sc.dG0, st));
  ARCVAE_CUDA(cudaMemsetAsync(sc.dtable, 0, (size_t)V * H3 * sizeof(float), st));
  ARCVAE_CUDA(cudaMemsetAsync(sc.dwc, 0, (size_t)H3 * C * sizeof(float), st));
  ARCVAE_CUDA(cudaMemsetAsync(sc.dWxc[0], 0, (size_t)H3 * (E + C) * sizeof(float), st));
  ARCVAE_CUDA(cudaMemsetAsync(sc.dbc[0], 0, (size_t)H3 * sizeof(float), st));
  ARCVAE_TRY(scatter_rows_by_token(sc.dG0, tp.in_tok, R, H3, V, sc.dtable, cond, B, C, sc.dwc, st));
  }
  ARCVAE_TRY(colsum(sc.dtable, V, H3, H3, sc.dbc[0], st));
  // table = Emb @ Wx0c[:, :E]^T + b0c
  ARCVAE_TRY(gemm_f32(0, 0, V, E, H3, sc.dtable, H3, tp.prep.Wxc[0], E + C, g->embedding, E, nullptr, true, id,
                      pick_splitk(V, E, H3), st));
  ARCVAE_TRY(gemm_f32(1, 0, H3, E, V, sc.dtable, H3, p->embedding, E, sc.dWxc[0], E + C, nullptr, true, id, 1, st));
  ARCVAE_TRY(add_strided(sc.dwc, C, sc.dWxc[0] + E, E + C, H3, C, st));
  ARCVAE_TRY(expand_gates_add(sc.dWxc[0], H, E + C, g->Wx[0], st));
  ARCVAE_TRY(expand_gates_add(sc.dbc[0], H, 1, g->bias[0], st));
  return 0;
}

// ---- sampler ---------------------------------------------------------------------------------------
namespace arcvae {
struct SamplerWs {
  DecPrep prep;
  int32_t* cur;      // [B]
  int32_t* ended;    // [B]
  int32_t* ended_count;
  float* hd[ARCVAE_MAX_LAYERS];  // [B,H]
  float* G[ARCVAE_MAX_LAYERS];   // [B,3H]
  float* logits;     // [B,V]
  __nv_bfloat16* hdb[ARCVAE_MAX_LAYERS];  // bf16 [B,H]
  __nv_bfloat16* gates_b[ARCVAE_MAX_LAYERS];  // bf16 [B,3H] (fused path scratch)
  void* sf_prep;                              // operands of the persistent sampler kernel (sampler_fused.cu)
};
static size_t sampler_layout(const arcvae_dims& d, int B, void* base, size_t cap, SamplerWs* w) {
  Arena a(base, cap);
  SamplerWs ww;
  dec_prep_layout(d, a, &ww.prep);
  ww.cur = a.take<int32_t>(B);
  ww.ended = a.take<int32_t>(B);
  ww.ended_count = a.take<int32_t>(4);
  for (int l = 0; l < d.NL; l++) {
    ww.hd[l] = a.take<float>((size_t)B * d.H);
    ww.G[l] = (l >= 1) ? a.take<float>((size_t)B * 3 * d.H) : nullptr;
  }
  ww.logits = a.take<float>((size_t)B * d.V);
  for (int l = 0; l < d.NL; l++) ww.hdb[l] = a.take<__nv_bfloat16>((size_t)B * d.H);
  for (int l = 0; l < d.NL; l++) ww.gates_b[l] = (l >= 1) ? a.take<__nv_bfloat16>((size_t)B * 3 * d.H) : nullptr;
  ww.sf_prep = a.take<char>(sampler_fused_prep_bytes(d));
  if (w) *w = ww;
  return align_up(a.off, 256);
}
}  // namespace arcvae

extern "C" size_t arcvae_sampler_workspace_bytes(const arcvae_dims* d, int B, int max_length) {
  (void)max_length;
  return d ? sampler_layout(*d, B, nullptr, 0, nullptr) : 0;
}

extern "C" int arcvae_sample(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond, int B,
                             int max_length, float temperature, int early_stopping, int multinomial, uint64_t seed,
                             int32_t* tokens, int32_t* t_stop, void* workspace, size_t workspace_bytes, int precision,
                             void* stream) {
  ARCVAE_TRY(check_dims_dec(d));
  ARCVAE_REQUIRE(B > 0 && max_length >= 0, "empty batch");
  ARCVAE_REQUIRE(temperature > 0.f, "temperature must be > 0");
  ARCVAE_REQUIRE(precision == ARCVAE_PREC_FP32 || precision == ARCVAE_PREC_BF16, "precision");
  ARCVAE_REQUIRE(tokens != nullptr && t_stop != nullptr, "outputs");
  cudaStream_t st = (cudaStream_t)stream;
  SamplerWs ws;
  size_t need = sampler_layout(*d, B, workspace, workspace_bytes, &ws);
  ARCVAE_REQUIRE(workspace != nullptr && need <= workspace_bytes, "sampler workspace too small");
  ARCVAE_TRY(dec_prepare(*d, p, ws.prep, precision, st));
  if (sampler_fused_supported(*d, precision)) {
    // one persistent kernel for the whole batch and all steps
    return sampler_fused_run(*d, ws.prep.table, ws.prep.wc, ws.prep.Wxpb, ws.prep.bp, ws.prep.Woutb, p->fc_out_b, cond, B,
                             max_length, temperature, early_stopping, multinomial, seed, tokens, t_stop, ws.ended_count,
                             ws.sf_prep, st);
  }
  ARCVAE_CUDA(cudaMemsetAsync(ws.cur, 0, (size_t)B * sizeof(int32_t), st));      // start token 0 (decoder_sampling.py:78)
  ARCVAE_CUDA(cudaMemsetAsync(ws.ended, 0, (size_t)B * sizeof(int32_t), st));
  ARCVAE_CUDA(cudaMemsetAsync(ws.ended_count, 0, 4 * sizeof(int32_t), st));
  ARCVAE_TRY(set_int(t_stop, max_length, st));
  RowMap id{nullptr, 1};
  for (int t = 0; t < max_length; t++) {
    // the reference checks `early_stopping and all(has_ended)` BEFORE each step (:87-88); the device records the first
    // such step in *t_stop and the host slices; later columns are scratch
    if (early_stopping && t > 0) ARCVAE_TRY(sampler_check_stop(ws.ended_count, B, t, t_stop, st));
    ARCVAE_TRY(dec_stack_forward(*d, p, ws.prep, cond, ws.cur, B, B, id, B, ws.hd, ws.hdb, ws.G, ws.gates_b, ws.logits,
                                 precision, dec_fused_ok(*d, precision, B, false), st));
    ARCVAE_TRY(select_token(ws.logits, B, d->V, temperature, multinomial, seed, t, max_length, d->end_token, tokens,
                            ws.cur, ws.ended, ws.ended_count, st));
  }
  return 0;
}
