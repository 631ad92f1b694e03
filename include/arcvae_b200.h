/*
 * arcvae_b200.h — C ABI of libarcvae_sm100.so: the B200 (sm_100a) implementation of the AR-CVAE
 * training step and sampler of Raiden-Makoto/MLX-VAE.
 *
 * The reference has no FFI/plugin boundary: its hot path is Python over MLX arrays
 * (SURVEY.md section 8b).  Each entry point below therefore replaces one reference *Python call*
 * (cited as file:line relative to the reference tree) and is what a ctypes / cffi / pybind binding
 * on the reference side would bind (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - fp32 everywhere the reference is fp32 (all of it); tokens are int32 (reference: uint32);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless documented;
 *   - return value 0 = ok, non-zero = error; arcvae_last_error() gives the message (thread-local);
 *   - the caller owns every buffer; workspaces are sized by arcvae_*_workspace_bytes();
 *   - gradient buffers ACCUMULATE: zero them before the step (arcvae_zero);
 *   - parameter layouts are the reference's: Linear.weight [out,in], LSTM Wx [4H,D], Wh [4H,H],
 *     bias [4H], gate order i,f,g,o (MLX nn.LSTM), Embedding.weight [V,E].
 */
#ifndef ARCVAE_B200_H
#define ARCVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARCVAE_MAX_LAYERS 8
#define ARCVAE_ABI_VERSION 2

/* model dimensions: train.py:26-31 (argparse defaults V80 E128 H256 L128 C1 NL2) */
typedef struct {
  int32_t V, E, H, L, C, NL;
  int32_t pad_token, end_token; /* decoder.py:25-26 */
} arcvae_dims;

/* models/encoder.py:46-74 — parameter (or gradient) pointers of MLXEncoder */
typedef struct {
  float* embedding;                 /* [V,E] */
  float* Wx[ARCVAE_MAX_LAYERS];     /* [4H, E] (layer 0) / [4H,H] */
  float* Wh[ARCVAE_MAX_LAYERS];     /* [4H,H] */
  float* bias[ARCVAE_MAX_LAYERS];   /* [4H] */
  float* condition_fc_w;            /* [H,C] */
  float* condition_fc_b;            /* [H] */
  float* fc_mu_w;                   /* [L,2H] */
  float* fc_mu_b;                   /* [L] */
  float* fc_logvar_hidden_w;        /* [2H,2H] */
  float* fc_logvar_hidden_b;        /* [2H] */
  float* fc_logvar_w;               /* [L,2H] */
  float* fc_logvar_b;               /* [L] */
} arcvae_encoder_params;

/* models/decoder.py:51-73 — parameter (or gradient) pointers of MLXAutoregressiveDecoder */
typedef struct {
  float* z_to_hidden_w;             /* [H,L]  (dead on the reference path, F1) */
  float* z_to_hidden_b;             /* [H] */
  float* condition_to_hidden_w;     /* [H,C]  (dead) */
  float* condition_to_hidden_b;     /* [H] */
  float* embedding;                 /* [V,E] */
  float* Wx[ARCVAE_MAX_LAYERS];     /* [4H, E+C] (layer 0) / [4H,H] */
  float* Wh[ARCVAE_MAX_LAYERS];     /* [4H,H]  (dead: the decoder LSTM is called without state) */
  float* bias[ARCVAE_MAX_LAYERS];   /* [4H] */
  float* fc_out_w;                  /* [V,H] */
  float* fc_out_b;                  /* [V] */
} arcvae_decoder_params;

/* complete_vae_loss.py:7-20 hyper-parameters (+ the opt-in switches of SURVEY.md section 8b) */
typedef struct {
  float beta, lambda_prop, lambda_collapse, free_bits, lambda_mi, target_mi;
  float collapse_target_mi;  /* posterior_collapse() is always called with its default 4.85 (complete_vae_loss.py:51) */
  int32_t pad_mask;          /* 0 = reference behaviour: unmasked mean over B*T (losses/recon.py:59-60) */
} arcvae_loss_hyper;

/* indices into the `losses` output of arcvae_loss_*: the scalar entries of the dict returned by
 * complete_vae_loss (complete_vae_loss.py:86-99) */
enum {
  ARCVAE_LOSS_TOTAL = 0, ARCVAE_LOSS_RECON, ARCVAE_LOSS_KL, ARCVAE_LOSS_WEIGHTED_KL, ARCVAE_LOSS_COLLAPSE,
  ARCVAE_LOSS_PROP, ARCVAE_LOSS_WEIGHTED_PROP, ARCVAE_LOSS_MI, ARCVAE_LOSS_MI_PENALTY, ARCVAE_LOSS_COUNT
};
/* the batch-statistics buffer: ARCVAE_STATS_REDUCE(L) doubles that ranks all-reduce(sum) under data parallelism
 * [sum_b m (L), sum_b exp(s) (L), sum kl_raw, sum kl_freebits, sum ce, count_tokens, count_batch]
 * followed by one 8-byte slot the kernel uses as a block counter; allocate ARCVAE_STATS_DOUBLES(L), zero it per step */
#define ARCVAE_STATS_REDUCE(L) (2 * (L) + 5)
#define ARCVAE_STATS_DOUBLES(L) (2 * (L) + 6)

/* precision of the tensor-pipe contractions */
enum { ARCVAE_PREC_FP32 = 0,   /* fp32 FFMA tiles: reference precision */
       ARCVAE_PREC_BF16 = 1 }; /* tcgen05 bf16 operands, fp32 accumulate */

const char* arcvae_last_error(void);
int arcvae_abi_version(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
uint64_t arcvae_launch_count(void);
/* tensor-core FLOP (2*M*N*K) this library has actually issued in the calling process for a timing category
 * (4 = tcgen05 GEMMs, 1 = cluster recurrence): bench.py's executed-FLOP roofline fraction */
double arcvae_flop_count(int category);
int arcvae_zero(void* ptr, size_t bytes, void* stream);
/* optional per-category device timing with CUDA events on the launch stream (bench.py's roofline leg).
 * categories: 0 fp32 GEMM, 1 recurrence, 2 fused loss, 3 Adam, 4 tcgen05 GEMM, 5 pointwise, 6 sampler.
 * arcvae_timing_read synchronises the device, fills ms[ncat] / counts[ncat] and resets.  Not thread-safe. */
#define ARCVAE_TIME_NCAT 8
int arcvae_timing_enable(int on);
int arcvae_timing_read(double* ms, int* counts, int ncat);

/* ---- encoder: models/encoder.py:76-132 (MLXEncoder.__call__) ------------------------------- */
size_t arcvae_encoder_tape_bytes(const arcvae_dims* d, int B, int T);
size_t arcvae_encoder_scratch_bytes(const arcvae_dims* d, int B, int T);
/* x [B,T] int32 (row-major, as MoleculeDataset.to_batches yields it), cond [B,C] -> mu, logvar [B,L].
 * `tape` keeps what BPTT needs. */
int arcvae_encoder_forward(const arcvae_dims* d, const arcvae_encoder_params* p, const int32_t* x, const float* cond,
                           int B, int T, float* mu, float* logvar, void* tape, size_t tape_bytes, int precision,
                           void* stream);
/* reverse pass of the above (what mx.value_and_grad records, trainer.py:292): accumulates into `g`. */
int arcvae_encoder_backward(const arcvae_dims* d, const arcvae_encoder_params* p, const float* cond, int B, int T,
                            const float* dmu, const float* dlogvar, void* tape, size_t tape_bytes,
                            const arcvae_encoder_params* g, void* scratch, size_t scratch_bytes, int precision,
                            void* stream);

/* 1 if the bf16 encoder recurrence for hidden size H runs as ONE persistent kernel per layer and direction (W_hh resident
 * on chip for all T steps), 0 if it falls back to one tcgen05 GEMM + cell launch per timestep */
int arcvae_recurrence_is_persistent(int H);

/* Error channel of the persistent kernels.  They wait with BOUNDED spins; a wait that runs out raises a STICKY
 * per-device flag (the library never clears it) and the kernel's results are garbage.  arcvae_adam_step refuses to
 * update while the flag is set, so a time-out cannot corrupt the weights; the host polls the flag when it wants
 * (the trainer mirror: every N steps and at the end of an epoch).
 *   arcvae_device_error_read : *flag_host = flag of the current device; synchronises `stream`
 *   arcvae_device_error_clear: resets it (after the caller has dealt with the failure)
 *   arcvae_debug_raise_device_error: test hook, raises it as a timed-out kernel would */
int arcvae_device_error_read(int* flag_host, void* stream);
int arcvae_device_error_clear(void* stream);
int arcvae_debug_raise_device_error(void* stream);

/* returns non-zero if the flag above is set (kept from ABI v1; the tape arguments are unused since v2).
 * Synchronises the stream. */
int arcvae_encoder_check(const arcvae_dims* d, int B, int T, void* tape, size_t tape_bytes, int precision, void* stream);

/* debug aid: clock64 stamps of CTA 0 of the cluster kernels into buf[64*16] (device, int64); NULL switches it off */
int arcvae_debug_set_rc_stamps(long long* buf);

/* ---- reparameterize: models/encoder.py:134-155 ---------------------------------------------- */
/* z = mu + eps*exp(0.5*logvar); eps==NULL draws N(0,1) from Philox4x32-10(seed, offset) */
int arcvae_reparameterize(const float* mu, const float* logvar, const float* eps, int B, int L, uint64_t seed,
                          uint64_t offset, float* z, void* stream);

/* ---- decoder: models/decoder.py:113-190 (MLXAutoregressiveDecoder.__call__) ------------------ */
size_t arcvae_decoder_tape_bytes(const arcvae_dims* d, int B, int T);
size_t arcvae_decoder_scratch_bytes(const arcvae_dims* d, int B, int T);
/* target [B,T] int32 or NULL; tf_mask_host[t] != 0 replaces the host coin of decoder.py:180 (NULL = all false).
 * Writes logits TIME-MAJOR [T,B,V] (element (b,t,v) at logits[(t*B+b)*V+v]); the Python mirror returns the
 * [B,T,V] view.  dec_inputs [T,B] int32 receives the token fed at every position. */
int arcvae_decoder_forward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond,
                           const int32_t* target, const uint8_t* tf_mask_host, int B, int T, float* logits_tm,
                           int32_t* dec_inputs_tm, void* tape, size_t tape_bytes, int precision, void* stream);
/* fc_out with the cross-entropy in its epilogue (bf16 fused path only: arcvae_decoder_ce_supported): the fp32 logits are
 * never written.  Per row: CE into *ce_sum (device double, accumulated: pass &stats[2L+2] of arcvae_loss_fwd_bwd),
 * d logits = (softmax - onehot(target)) * ce_scale as bf16 into the tape, and the greedy feedback token where the host
 * coin of the position is false (decoder.py:185).  ce_scale = 1 / (GLOBAL number of positions): unmasked mean of
 * losses/recon.py:59-60.  Follow with arcvae_decoder_backward(dlogits_tm = NULL). */
int arcvae_decoder_ce_supported(const arcvae_dims* d, int B, int precision);
int arcvae_decoder_forward_ce(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond,
                              const int32_t* target, const uint8_t* tf_mask_host, int B, int T, float ce_scale,
                              double* ce_sum, int32_t* dec_inputs_tm, void* tape, size_t tape_bytes, int precision,
                              void* stream);
/* dlogits_tm [T,B,V] (destroyed), or NULL after arcvae_decoder_forward_ce; accumulates into `g`. */
int arcvae_decoder_backward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond, int B, int T,
                            float* dlogits_tm, void* tape, size_t tape_bytes, const arcvae_decoder_params* g,
                            void* scratch, size_t scratch_bytes, int precision, void* stream);

/* ---- decoder with carry_state=True — an EXTENSION, not the reference's behaviour: the state built by
 * initialize_hidden_state (models/decoder.py:76-111) is consumed and carried from position to position (the reference
 * computes it at :143 and never passes it to the LSTM layers, :165-168).  Same tensors and conventions as
 * arcvae_decoder_forward / _backward plus z [B,L]; the reverse pass also returns dz [B,L] (d total / d z), which
 * arcvae_reparam_backward folds into d mu / d logvar (z = mu + eps*exp(logvar/2), models/encoder.py:147-153). */
size_t arcvae_decoder_cs_tape_bytes(const arcvae_dims* d, int B, int T);
size_t arcvae_decoder_cs_scratch_bytes(const arcvae_dims* d, int B, int T);
int arcvae_decoder_cs_forward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* z, const float* cond,
                              const int32_t* target, const uint8_t* tf_mask_host, int B, int T, float* logits_tm,
                              int32_t* dec_inputs_tm, void* tape, size_t tape_bytes, int precision, void* stream);
int arcvae_decoder_cs_backward(const arcvae_dims* d, const arcvae_decoder_params* p, const float* z, const float* cond,
                               int B, int T, float* dlogits_tm, void* tape, size_t tape_bytes,
                               const arcvae_decoder_params* g, float* dz, void* scratch, size_t scratch_bytes,
                               int precision, void* stream);
/* dmu += dz ; dlogvar += dz * (z - mu) / 2 */
int arcvae_reparam_backward(const float* dz, const float* z, const float* mu, int B, int L, float* dmu, float* dlogvar,
                            void* stream);

/* ---- losses: losses/recon.py:29-64, losses/kl.py:35-66, losses/info.py:23-78,
 *      complete_vae_loss.py:45-99 — one fused kernel, forward values and gradients --------------- */
/* logits element (b,t,v) at logits[b*ls_b + t*ls_t + v]; targets (b,t) at targets[b*ts_b + t*ts_t].
 * phases: bit0 = batch statistics into `stats` (zeroed by the caller), bit1 = finish (losses[], dlogits,
 * dmu, dlogvar, z).  Single GPU: phases=3 (one cooperative launch).  Data parallel: phases=1,
 * all-reduce(sum) `stats`, phases=2.  Any of logits / mu may be NULL to skip that half.
 * dlogits may alias logits.  eps==NULL -> Philox(seed, offset).
 * phases bit2 (value 4, with logits == NULL and T given): the cross-entropy sum of this shard already sits in
 * stats[2L+2] (arcvae_decoder_forward_ce accumulated it there, scaled d logits are in the decoder tape); the kernel adds
 * the token count B*T and reports recon / total from it. */
int arcvae_loss_fwd_bwd(const float* logits, int64_t ls_b, int64_t ls_t, const int32_t* targets, int64_t ts_b,
                        int64_t ts_t, int B, int T, int V, int pad_token, const float* mu, const float* logvar,
                        const float* eps, int L, const arcvae_loss_hyper* hyper, uint64_t seed, uint64_t offset,
                        int phases, double* stats, float* losses, float* dlogits, float* dmu, float* dlogvar,
                        float* z, void* stream);

/* ---- sampler: models/decoder_sampling.py:48-128 (generate_with_temperature) ------------------ */
size_t arcvae_sampler_workspace_bytes(const arcvae_dims* d, int B, int max_length);
/* tokens [B,max_length] int32 row-major; *t_stop (device int32) = number of valid leading columns
 * (the reference stops before the first step at which every row has emitted end_token, :87-88).
 * multinomial=0: argmax(softmax(logits/temperature)) as the reference (:110-117);
 * multinomial=1: categorical draw with Philox(seed, row, step) (the reference's TODO, :116). */
int arcvae_sample(const arcvae_dims* d, const arcvae_decoder_params* p, const float* cond, int B, int max_length,
                  float temperature, int early_stopping, int multinomial, uint64_t seed, int32_t* tokens,
                  int32_t* t_stop, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ---- optimizer: mlx.optimizers.Adam as used by trainer.py:75-76, :320-324 (no bias correction).
 * A no-op while the device error flag is set (see arcvae_device_error_read). */
int arcvae_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                     float eps, float grad_scale, void* stream);
/* sum of squares of g into *out (device double, accumulated) — for clip_mode='global_norm' (not the reference's
 * no-op clip, trainer.py:489-522) */
int arcvae_sumsq(const float* g, size_t n, double* out, void* stream);

/* ---- building block exposed for tests: C[M,N] (+)= op(A) op(B) (+ bias[N]) in fp32 ----------- */
int arcvae_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                    float* C, int ldc, const float* bias, int accumulate, void* stream);

/* bf16 tensor-core GEMM (tcgen05 + TMA) exposed for tests: A, B point to bf16 data.
 * a_mn=0: A[m*lda+k], 1: A[k*lda+m]; b_mn=0: B[n*ldb+k], 1: B[k*ldb+n]; C fp32 and/or Cb bf16 outputs. */
int arcvae_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb, float* C,
                     int ldc, void* Cb, int ldcb, const float* bias, int accumulate, int splitk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARCVAE_B200_H */
