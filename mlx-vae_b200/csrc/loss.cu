// loss.cu — the fused AR-CVAE loss kernel: forward values AND gradients of
//   reconstruction_loss (losses/recon.py:29-64)      mean_{b,t} [logsumexp(l) - l[x]]   (optionally pad-masked)
//   reparameterize      (models/encoder.py:147-153)  z = mu + eps*exp(logvar/2), eps injected or Philox4x32-10
//   kl_divergence       (losses/kl.py:35-66)         clips, max(.,0), free-bits floor, sum_d, mean_b
//   mutual_information  (losses/info.py:23-50)       mean KL - KL(moment-matched aggregate), clamp >= 0
//   posterior_collapse  (losses/info.py:73-78)       w * max(0, 4.85 - MI)
//   complete_vae_loss   (complete_vae_loss.py:45-99) total and the scalar dict entries
// in ONE cooperative launch (phase 0: batch statistics -> grid.sync -> phase 1: gradients + scalars).
// Under data parallelism the two phases are launched separately around an all-reduce of `stats`.
//
// HBM-bound: algorithmic bytes per molecule = T*V*4 (logits read) + T*V*4 (dlogits write) + T*4 (tokens)
// = 82.4 KB at T=128, V=80.  Rows are handled by sub-warp groups with float4 accesses (V=80: 4 lanes x 5 float4),
// so a warp streams 8 adjacent rows = 2560 contiguous bytes per pass.
//
// Sub-gradient conventions (MLX `maximum(a,b)`: cotangent to a where a > b, else to b) are kept:
//   d max(k,0)/dk = [k>0];  d max(k',fb/L)/dk' = [k'>fb/L];  d max(mi,0) = [mi>0];  d max(0,t-MI)/d(t-MI) = [t-MI>=0].
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace arcvae {

struct LossArgs {
  const float* logits; long ls_b, ls_t;
  const int32_t* targets; long ts_b, ts_t;
  int B, T, V, pad_token, b_fastest;
  const float* mu; const float* logvar; const float* eps; int L;
  arcvae_loss_hyper hp;
  uint64_t seed, offset;
  int phases;
  double* stats;  // [2L+5] + 1 internal counter
  float* losses; float* dlogits; float* dmu; float* dlogvar; float* z;
};

__device__ __forceinline__ float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    r = warp_sum_d(r);
  }
  return r;  // valid in warp 0
}

// ---- cross-entropy rows ---------------------------------------------------------------------------
// float4 path: groups of G lanes, each lane holds up to MAXQ float4 of its row in registers.
template <int G, int MAXQ>
__device__ __forceinline__ double ce_rows_vec(const LossArgs& a, float scale) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;                          // lane within group
  const long groups_per_block = blockDim.x / G;
  const long gid = blockIdx.x * groups_per_block + threadIdx.x / G;
  const long ngroups = (long)gridDim.x * groups_per_block;
  const long R = (long)a.B * a.T;
  const int nq = a.V >> 2;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
  double ce_acc = 0.0;
  for (long rho = gid; rho < R; rho += ngroups) {
    int b, t;
    if (a.b_fastest) { t = (int)(rho / a.B); b = (int)(rho - (long)t * a.B); }
    else             { b = (int)(rho / a.T); t = (int)(rho - (long)b * a.T); }
    const long off = b * a.ls_b + t * a.ls_t;
    const float4* src = reinterpret_cast<const float4*>(a.logits + off);
    const int tgt = a.targets[b * a.ts_b + t * a.ts_t];
    float4 v[MAXQ];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXQ; i++) {
      int q = gl + i * G;
      if (q < nq) {
        v[i] = __ldg(src + q);
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, o));
    float s = 0.f, lt = 0.f;
#pragma unroll
    for (int i = 0; i < MAXQ; i++) {
      int q = gl + i * G;
      if (q < nq) {
        float4 e;
        e.x = expf(v[i].x - mx); e.y = expf(v[i].y - mx); e.z = expf(v[i].z - mx); e.w = expf(v[i].w - mx);
        s += (e.x + e.y) + (e.z + e.w);
        int base = q * 4;
        if (tgt >= base && tgt < base + 4) {
          int k = tgt - base;
          lt = (k == 0) ? v[i].x : (k == 1) ? v[i].y : (k == 2) ? v[i].z : v[i].w;
        }
        v[i] = e;
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      s += __shfl_xor_sync(gmask, s, o);
      lt += __shfl_xor_sync(gmask, lt, o);
    }
    const bool masked = a.hp.pad_mask && (tgt == a.pad_token);
    const float lse = mx + logf(s);
    if (gl == 0 && !masked) ce_acc += (double)(lse - lt);
    if (a.dlogits != nullptr) {
      const float inv = masked ? 0.f : scale / s;
      const float sub = masked ? 0.f : scale;
      float4* dst = reinterpret_cast<float4*>(a.dlogits + off);
#pragma unroll
      for (int i = 0; i < MAXQ; i++) {
        int q = gl + i * G;
        if (q < nq) {
          float4 g;
          g.x = v[i].x * inv; g.y = v[i].y * inv; g.z = v[i].z * inv; g.w = v[i].w * inv;
          int base = q * 4;
          if (tgt >= base && tgt < base + 4) {
            int k = tgt - base;
            if (k == 0) g.x -= sub; else if (k == 1) g.y -= sub; else if (k == 2) g.z -= sub; else g.w -= sub;
          }
          dst[q] = g;
        }
      }
    }
  }
  return ce_acc;
}

// scalar path: one warp per row, any V / alignment.
__device__ __forceinline__ double ce_rows_scalar(const LossArgs& a, float scale) {
  const int lane = threadIdx.x & 31;
  const long wpb = blockDim.x >> 5;
  const long wid = blockIdx.x * wpb + (threadIdx.x >> 5);
  const long nw = (long)gridDim.x * wpb;
  const long R = (long)a.B * a.T;
  double ce_acc = 0.0;
  for (long rho = wid; rho < R; rho += nw) {
    int b, t;
    if (a.b_fastest) { t = (int)(rho / a.B); b = (int)(rho - (long)t * a.B); }
    else             { b = (int)(rho / a.T); t = (int)(rho - (long)b * a.T); }
    const long off = b * a.ls_b + t * a.ls_t;
    const float* row = a.logits + off;
    const int tgt = a.targets[b * a.ts_b + t * a.ts_t];
    float mx = -INFINITY;
    for (int v = lane; v < a.V; v += 32) mx = fmaxf(mx, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int v = lane; v < a.V; v += 32) s += expf(row[v] - mx);
    s = warp_sum(s);
    const bool masked = a.hp.pad_mask && (tgt == a.pad_token);
    const float lse = mx + logf(s);
    if (lane == 0 && !masked) ce_acc += (double)(lse - row[tgt]);
    if (a.dlogits != nullptr) {
      float* drow = a.dlogits + off;
      for (int v = lane; v < a.V; v += 32) {
        float g = masked ? 0.f : (expf(row[v] - mx) / s - (v == tgt ? 1.f : 0.f)) * scale;
        drow[v] = g;
      }
    }
  }
  return ce_acc;
}

__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t ctr) {
  uint32_t r[4];
  philox4x32(seed, ctr, 0ull, r);
  float u1 = u01(r[0]), u2 = u01(r[1]);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

__global__ void __launch_bounds__(256, 4) k_loss_fused(LossArgs a) {
  __shared__ double shd[8];
  const int L = a.L;
  double* st_mu = a.stats;            // [L]
  double* st_var = a.stats + L;       // [L]
  double* st_sc = a.stats + 2 * L;    // kl_raw, kl_fb, ce, count_tokens, count_batch
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(a.stats + 2 * L + 5);
  const float fb_floor = (a.hp.free_bits > 0.f && L > 0) ? a.hp.free_bits / (float)L : 0.f;

  // ------------------------------------------------------------------ phase 0: batch statistics
  if (a.phases & 1) {
    if (a.mu != nullptr) {
      double klr = 0.0, klf = 0.0;
      for (int d = threadIdx.x; d < L; d += blockDim.x) {
        double sm = 0.0, sv = 0.0;
        for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
          float m = clipf(a.mu[(long)b * L + d], -3.f, 3.f);
          float s = clipf(a.logvar[(long)b * L + d], -6.f, 3.f);
          float var = expf(s);
          float k = -0.5f * (1.0f + s - m * m - var);
          float kc = fmaxf(k, 0.f);
          if (a.hp.free_bits > 0.f) kc = fmaxf(kc, fb_floor);
          sm += m; sv += var; klr += k; klf += kc;
        }
        atomicAdd(st_mu + d, sm);
        atomicAdd(st_var + d, sv);
      }
      double r0 = block_sum_d(klr, shd);
      double r1 = block_sum_d(klf, shd);
      if (threadIdx.x == 0) {
        atomicAdd(st_sc + 0, r0);
        atomicAdd(st_sc + 1, r1);
        if (blockIdx.x == 0) atomicAdd(st_sc + 4, (double)a.B);
      }
    }
    if (a.logits == nullptr && (a.phases & 4)) {
      // the cross-entropy sum was accumulated into st_sc[2] by the fc_out epilogue (gemm_tc.cu TC_EPI_CE): only the
      // token count of this shard is missing (unmasked mean, losses/recon.py:59-60)
      if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(st_sc + 3, (double)a.B * (double)a.T);
    }
    if (a.logits != nullptr) {
      double cnt = 0.0;
      if (a.hp.pad_mask) {
        long R = (long)a.B * a.T;
        for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < R; i += (long)gridDim.x * blockDim.x) {
          int b = (int)(i / a.T), t = (int)(i - (long)b * a.T);
          cnt += (a.targets[b * a.ts_b + t * a.ts_t] != a.pad_token) ? 1.0 : 0.0;
        }
        cnt = block_sum_d(cnt, shd);
      } else if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt = (double)a.B * (double)a.T;
      }
      if (threadIdx.x == 0 && cnt != 0.0) atomicAdd(st_sc + 3, cnt);
    }
  }
  if ((a.phases & 3) == 3) {
    __threadfence();
    cg::this_grid().sync();
  }
  if (!(a.phases & 2)) return;

  // ------------------------------------------------------------------ phase 1: gradients
  // volatile reads: the statistics were produced by other blocks / another launch / NCCL
  const volatile double* vst = a.stats;
  const double Bg = vst[2 * L + 4] > 0.0 ? vst[2 * L + 4] : (double)a.B;
  const double ntok = vst[2 * L + 3];

  double mi_raw = 0.0, agg = 0.0, mean_kl = 0.0;
  float w_mi = 0.f;
  if (a.mu != nullptr) {
    double part = 0.0;
    for (int d = threadIdx.x; d < L; d += blockDim.x) {
      double mm = vst[d] / Bg, mv = vst[L + d] / Bg;
      part += 1.0 + log(mv) - mm * mm - mv;
    }
    double r = block_sum_d(part, shd);
    if (threadIdx.x == 0) shd[0] = -0.5 * r;
    __syncthreads();
    agg = shd[0];
    __syncthreads();
    mean_kl = vst[2 * L + 0] / Bg;
    mi_raw = mean_kl - agg;
    float mi = (float)fmax(mi_raw, 0.0);
    float gate_mi = (mi_raw > 0.0) ? 1.f : 0.f;
    float g_col = (a.hp.collapse_target_mi - mi >= 0.f) ? a.hp.lambda_collapse : 0.f;
    float g_pen = (a.hp.target_mi - mi >= 0.f) ? a.hp.lambda_mi : 0.f;
    w_mi = -(g_col + g_pen) * gate_mi;

    const float invB = (float)(1.0 / Bg);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
      for (int d = threadIdx.x; d < L; d += blockDim.x) {
        long idx = (long)b * L + d;
        float mu_raw = a.mu[idx], lv_raw = a.logvar[idx];
        float m = clipf(mu_raw, -3.f, 3.f), s = clipf(lv_raw, -6.f, 3.f);
        float pass_m = (mu_raw >= -3.f && mu_raw <= 3.f) ? 1.f : 0.f;
        float pass_s = (lv_raw >= -6.f && lv_raw <= 3.f) ? 1.f : 0.f;
        float var = expf(s);
        float k = -0.5f * (1.0f + s - m * m - var);
        float gate = (k > 0.f) ? 1.f : 0.f;
        if (a.hp.free_bits > 0.f) gate = (fmaxf(k, 0.f) > fb_floor) ? gate : 0.f;
        float mbar = (float)(vst[d] / Bg), vbar = (float)(vst[L + d] / Bg);
        float gm = a.hp.beta * invB * gate * m + w_mi * invB * (m - mbar);
        float gs = a.hp.beta * invB * gate * 0.5f * (var - 1.f) + w_mi * invB * 0.5f * (var / vbar - 1.f);
        if (a.dmu != nullptr) a.dmu[idx] = gm * pass_m;
        if (a.dlogvar != nullptr) a.dlogvar[idx] = gs * pass_s;
        if (a.z != nullptr) {
          float e = (a.eps != nullptr) ? a.eps[idx] : philox_normal(a.seed, a.offset + (uint64_t)idx);
          a.z[idx] = mu_raw + e * expf(0.5f * lv_raw);
        }
      }
    }
  }

  if (a.logits != nullptr) {
    const float scale = (ntok > 0.0) ? (float)(1.0 / ntok) : 0.f;
    double ce = 0.0;
    const bool vec_ok = ((a.V & 3) == 0) && ((a.ls_b & 3) == 0) && ((a.ls_t & 3) == 0) &&
                        ((((uintptr_t)a.logits) & 15) == 0) &&
                        (a.dlogits == nullptr || (((uintptr_t)a.dlogits) & 15) == 0);
    if (vec_ok && a.V <= 80) ce = ce_rows_vec<4, 5>(a, scale);
    else if (vec_ok && a.V <= 160) ce = ce_rows_vec<8, 5>(a, scale);
    else if (vec_ok && a.V <= 1024) ce = ce_rows_vec<32, 8>(a, scale);
    else ce = ce_rows_scalar(a, scale);
    ce = block_sum_d(ce, shd);
    if (threadIdx.x == 0) atomicAdd(a.stats + 2 * L + 2, ce);
  }

  // ------------------------------------------------------------------ scalars: last block to finish
  __shared__ int is_last;
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned long long prev = atomicAdd(counter, 1ull);
    is_last = (prev == (unsigned long long)gridDim.x - 1ull);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0 && a.losses != nullptr) {
    __threadfence();
    float recon = 0.f, kl = 0.f, collapse = 0.f, mi = 0.f, mi_pen = 0.f;
    if ((a.logits != nullptr || (a.phases & 4)) && ntok > 0.0) recon = (float)(vst[2 * L + 2] / ntok);
    if (a.mu != nullptr) {
      kl = (float)(vst[2 * L + 1] / Bg);
      mi = (float)fmax(mi_raw, 0.0);
      collapse = a.hp.lambda_collapse * fmaxf(0.f, a.hp.collapse_target_mi - mi);
      mi_pen = a.hp.lambda_mi * fmaxf(0.f, a.hp.target_mi - mi);
    }
    const float prop = 0.f;  // property_predictor is None (train.py:186, complete_vae_loss.py:67)
    a.losses[ARCVAE_LOSS_RECON] = recon;
    a.losses[ARCVAE_LOSS_KL] = kl;
    a.losses[ARCVAE_LOSS_WEIGHTED_KL] = a.hp.beta * kl;
    a.losses[ARCVAE_LOSS_COLLAPSE] = collapse;
    a.losses[ARCVAE_LOSS_PROP] = prop;
    a.losses[ARCVAE_LOSS_WEIGHTED_PROP] = a.hp.lambda_prop * prop;
    a.losses[ARCVAE_LOSS_MI] = mi;
    a.losses[ARCVAE_LOSS_MI_PENALTY] = mi_pen;
    a.losses[ARCVAE_LOSS_TOTAL] = recon + a.hp.beta * kl + collapse + a.hp.lambda_prop * prop + mi_pen;
    *counter = 0ull;
  }
}

}  // namespace arcvae

extern "C" int arcvae_loss_fwd_bwd(const float* logits, int64_t ls_b, int64_t ls_t, const int32_t* targets,
                                   int64_t ts_b, int64_t ts_t, int B, int T, int V, int pad_token, const float* mu,
                                   const float* logvar, const float* eps, int L, const arcvae_loss_hyper* hyper,
                                   uint64_t seed, uint64_t offset, int phases, double* stats, float* losses,
                                   float* dlogits, float* dmu, float* dlogvar, float* z, void* stream) {
  using namespace arcvae;
  ARCVAE_REQUIRE(hyper != nullptr && stats != nullptr, "hyper and stats are mandatory");
  ARCVAE_REQUIRE((phases & 3) >= 1 && (phases & ~7) == 0, "phases: bit0 statistics, bit1 finish, bit2 cross-entropy precomputed");
  ARCVAE_REQUIRE(!(phases & 4) || (logits == nullptr && T > 0 && !hyper->pad_mask),
                 "precomputed cross-entropy: no logits, T given, unmasked mean");
  ARCVAE_REQUIRE(logits == nullptr || targets != nullptr, "logits need targets");
  ARCVAE_REQUIRE(mu == nullptr || logvar != nullptr, "mu needs logvar");
  ARCVAE_REQUIRE(B > 0, "empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  LossArgs a;
  a.logits = logits; a.ls_b = ls_b; a.ls_t = ls_t;
  a.targets = targets; a.ts_b = ts_b; a.ts_t = ts_t;
  a.B = B; a.T = T; a.V = V; a.pad_token = pad_token;
  a.b_fastest = (ls_b < ls_t) ? 1 : 0;
  a.mu = mu; a.logvar = logvar; a.eps = eps; a.L = L;
  a.hp = *hyper; a.seed = seed; a.offset = offset; a.phases = phases;
  a.stats = stats; a.losses = losses; a.dlogits = dlogits; a.dmu = dmu; a.dlogvar = dlogvar; a.z = z;

  // co-resident blocks (cooperative launch): the kernel's occupancy is a property of the code (sm_100a only), the SM
  // count is the current device's
  static std::atomic<int> per_sm{0};
  int per = per_sm.load();
  if (per == 0) {
    ARCVAE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_loss_fused, 256, 0));
    if (per > 8) per = 8;
    if (per < 1) per = 1;
    per_sm.store(per);
  }
  const int max_blocks = device_sm_count() * per;
  long rows = (logits != nullptr) ? (long)B * T : 0;
  long want = (rows * 4 + 255) / 256;       // 4 lanes per row on the float4 path
  if (want < B) want = B;
  int grid = (int)(want < max_blocks ? want : max_blocks);
  if (grid < 1) grid = 1;
  TimeScope ts(TIME_LOSS, st);
  if ((phases & 3) == 3) {
    void* args[] = {&a};
    ARCVAE_CUDA(cudaLaunchCooperativeKernel((void*)k_loss_fused, dim3(grid), dim3(256), args, 0, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    k_loss_fused<<<grid, 256, 0, st>>>(a);
    ARCVAE_LAUNCHED();
  }
  return 0;
}

namespace arcvae {
__global__ void k_reparameterize(const float* __restrict__ mu, const float* __restrict__ logvar,
                                 const float* __restrict__ eps, long n, uint64_t seed, uint64_t offset,
                                 float* __restrict__ z) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float e = (eps != nullptr) ? eps[i] : philox_normal(seed, offset + (uint64_t)i);
    z[i] = mu[i] + e * expf(0.5f * logvar[i]);
  }
}
}  // namespace arcvae

extern "C" int arcvae_reparameterize(const float* mu, const float* logvar, const float* eps, int B, int L,
                                     uint64_t seed, uint64_t offset, float* z, void* stream) {
  using namespace arcvae;
  long n = (long)B * L;
  if (n <= 0) return 0;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  k_reparameterize<<<grid, 256, 0, (cudaStream_t)stream>>>(mu, logvar, eps, n, seed, offset, z);
  ARCVAE_LAUNCHED();
  return 0;
}
