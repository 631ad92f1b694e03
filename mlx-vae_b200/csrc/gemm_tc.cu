// gemm_tc.cu — bf16 tensor-core GEMM for the time-parallel contractions of the AR-CVAE step:
//   D[M,N] (+)= A[M,K] * B[K,N] (+ bias), bf16 operands, fp32 accumulation in TMEM, fp32 and/or bf16 output.
//
// sm_100a structure (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) into a multi-stage smem ring, mbarrier complete_tx
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16 per instruction),
//                              tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..9  epilogue       tcgen05.ld (32x32b) of the accumulator quarter they own (two warps per quarter, half of
//                              the columns each), bias / accumulate / split-K atomics / fused decoder cells
// Two accumulators (2 x BN TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Operand layouts (both served by TMA, no transposes in HBM):
//   K-major   A[m*lda + k], B[n*ldb + k]   forward projections, input gradients against pre-transposed weights
//   MN-major  A[k*lda + m], B[k*ldb + n]   weight gradients dW = dA^T @ H with K = B*T (split-K + fp32 atomics)
#include <cstdlib>
#include <mutex>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 8;            // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

struct TcParams {
  int M, N, K;
  int BN;            // N tile (multiple of 16, <= 256)
  int stages;
  int a_mn, b_mn;    // operand major-ness
  int mt, nt, splitk, kb_per;
  long split_stride;
  float* C; int ldc;
  bf16* Cb; int ldcb;
  const float* bias;
  int accumulate;
  RowMap rm;
  // fused decoder epilogues
  int epi;
  bf16* gates_b; bf16* hb_out; bf16* dg_out;
  int Hh;
  const bf16* pre_b; const float* c_prev; float* c_out; float* hf_out;     // TC_EPI_LSTM_FWD
  const int32_t* ce_target; const uint8_t* ce_fb; int32_t* ce_tok; int ce_B; bf16* ce_dl; int ce_ldl; double* ce_sum;
  float ce_scale;                                                          // TC_EPI_CE
  // multi-segment B (weight gradients that share the A operand): column tile ni multiplies A with its own B matrix
  // (rows shifted by seg_shift[ni]; negative TMA coordinates zero-fill) into its own C
  int nseg; int segN[3]; int seg_shift[3]; float* segC[3]; int seg_ldc[3];
  int bm2;           // 256-row tiles (two M=128 MMAs share every B tile): halves the B re-reads of the long-K weight-gradient shapes
  int lp_B;          // TC_EPI_LSTM_DH: rows per timestep of the thread-friendly output
  int fs;            // two-segment weight gradients in ONE CTA: the A tile is loaded once and multiplied with both B segments
                     // (accumulators side by side in TMEM); see the host side for why
  int use_scratch;   // per-warp transposition scratch present after the TcShared block
  int scr_pitch;     // bytes per scratch row
};

__device__ __forceinline__ float tanh_fast_(float x) { return tanh_approx_(x); }
__device__ __forceinline__ float sigmoid_fast_(float x) { return sigmoid_approx_(x); }
// zero-state decoder cell backward: (i, g, o activated, dh) -> pre-activation gradients
__device__ __forceinline__ void dec_cell_grads_fast(float i_, float g_, float o_, float dh, float& dai, float& dag,
                                                    float& dao) {
  const float tcv = tanh_fast_(i_ * g_);
  const float dcc = dh * o_ * (1.f - tcv * tcv);
  dao = dh * tcv * o_ * (1.f - o_);
  dai = dcc * g_ * i_ * (1.f - i_);
  dag = dcc * i_ * (1.f - g_ * g_);
}

// ---- per-warp transposition scratch: the epilogue thread owns a ROW (tcgen05.ld 32x32b), but a warp-wide 16-byte
// access to 32 different rows costs 32 L1 wavefronts instead of 4.  Row segments are therefore staged in shared
// memory (pitch 400 B: conflict-free for the thread=row side) and moved to / from HBM with lanes running along the row.
constexpr int TC_SCR_PITCH_BWD = 400;             // 384-byte rows (a 64-unit block of i|g|o) + 16
constexpr int TC_SCR_PITCH_FWD = 272;             // 256-byte rows + 16
template <int CPR>   // 16-byte chunks per row segment (compile-time: the index split is a shift / constant division)
__device__ __forceinline__ void scr_store_rows_t(const uint8_t* scr, int pitch, uint8_t* gbase, long gstride, int nrows, int lane) {
  const int total = nrows * CPR;
#pragma unroll 4
  for (int idx = lane; idx < total; idx += 32) {
    const int r = idx / CPR, c = idx - r * CPR;
    *reinterpret_cast<uint4*>(gbase + r * gstride + c * 16) = *reinterpret_cast<const uint4*>(scr + r * pitch + c * 16);
  }
}
__device__ __forceinline__ void scr_store_rows(const uint8_t* scr, int pitch, uint8_t* gbase, long gstride, int nb, int nrows, int lane) {
  switch (nb >> 4) {
    case 4: scr_store_rows_t<4>(scr, pitch, gbase, gstride, nrows, lane); break;
    case 8: scr_store_rows_t<8>(scr, pitch, gbase, gstride, nrows, lane); break;
    case 16: scr_store_rows_t<16>(scr, pitch, gbase, gstride, nrows, lane); break;
    case 24: scr_store_rows_t<24>(scr, pitch, gbase, gstride, nrows, lane); break;
    default: {
      const int cpr = nb >> 4, total = nrows * cpr;
      for (int idx = lane; idx < total; idx += 32) {
        const int r = idx / cpr, c = idx - r * cpr;
        *reinterpret_cast<uint4*>(gbase + r * gstride + c * 16) = *reinterpret_cast<const uint4*>(scr + r * pitch + c * 16);
      }
    }
  }
}
// 384-byte row segments (24 chunks): batches of 8 independent 16-byte loads per lane, so one HBM/L2 latency covers 4 KB
// per warp instead of 512 B
__device__ __forceinline__ void scr_load_rows(uint8_t* scr, int pitch, const uint8_t* gbase, long gstride, int nb, int nrows, int lane) {
  const int cpr = nb >> 4, total = nrows * cpr;
  for (int base = 0; base < total; base += 8 * 32) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int idx = base + u * 32 + lane;
      if (idx < total) {
        const int r = (cpr == 24) ? idx / 24 : idx / cpr, c = idx - r * cpr;
        v[u] = __ldg(reinterpret_cast<const uint4*>(gbase + r * gstride + c * 16));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int idx = base + u * 32 + lane;
      if (idx < total) {
        const int r = (cpr == 24) ? idx / 24 : idx / cpr, c = idx - r * cpr;
        *reinterpret_cast<uint4*>(scr + r * pitch + c * 16) = v[u];
      }
    }
  }
}
__device__ __forceinline__ void scr_put16(uint8_t* dst, const float (&v)[16]) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(dst + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
// 16 floats -> 16 bf16 in two 16-byte registers, stored to a (thread-private) global row segment
__device__ __forceinline__ void st_bf16x16(bf16* dst, const float (&v)[16]) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
__device__ __forceinline__ void scr_get16(const uint8_t* src, float (&v)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(src);
  const uint4 b = *reinterpret_cast<const uint4*>(src + 16);
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 t = __bfloat1622float2(pa[j]);
    v[2 * j] = t.x; v[2 * j + 1] = t.y;
    float2 w = __bfloat1622float2(pb[j]);
    v[8 + 2 * j] = w.x; v[8 + 2 * j + 1] = w.y;
  }
}

struct __align__(8) TcShared {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// ---- fc_out epilogue with the cross-entropy (TC_EPI_CE) ---------------------------------------------------------------
// NCH = number of 16-column chunks of the row (0: run-time, up to 8).  Called by ALL lanes of the warp (tcgen05.ld is
// .sync.aligned); rows outside the matrix compute on garbage and store nothing.
template <int NCH>
__device__ __forceinline__ void ce_epilogue(const TcParams& p, uint32_t taddr, int half, int q, int lane, bool row_ok, long grow,
                                            float4 (*ce_x)[32], float& ce_acc) {
  constexpr int MAXC = NCH > 0 ? (NCH + 1) / 2 : 4;            // chunks per warp (first warp takes the larger share)
  const int nch = NCH > 0 ? NCH : ((p.N + 15) >> 4);
  const int c_lo = half == 0 ? 0 : (nch + 1) >> 1;
  const int c_hi = half == 0 ? (nch + 1) >> 1 : nch;
  const int V = p.N;
  float v[MAXC * 16];
#pragma unroll
  for (int i = 0; i < MAXC; i++) {
    const int ch = c_lo + i;
    if (ch < c_hi) {                                           // warp-uniform
      uint32_t r[16];
      tc::tmem_ld16(taddr + ch * 16, r);
      float4 bv[4];
#pragma unroll
      for (int k4 = 0; k4 < 4; k4++)
      {
        const int c4 = ch * 16 + k4 * 4;
        if (c4 + 3 < V) bv[k4] = __ldg(reinterpret_cast<const float4*>(p.bias + ch * 16) + k4);
        else bv[k4] = make_float4(c4 < V ? __ldg(p.bias + c4) : 0.f, c4 + 1 < V ? __ldg(p.bias + c4 + 1) : 0.f,
                                  c4 + 2 < V ? __ldg(p.bias + c4 + 2) : 0.f, 0.f);
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int k4 = 0; k4 < 4; k4++) {
        const float b4[4] = {bv[k4].x, bv[k4].y, bv[k4].z, bv[k4].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int col = ch * 16 + k4 * 4 + j;
          v[i * 16 + k4 * 4 + j] = col < V ? __uint_as_float(r[k4 * 4 + j]) + b4[j] : -INFINITY;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++) v[i * 16 + j] = -INFINITY;
    }
  }
  // local max / argmax (four independent chains, merged lowest-index-first)
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int b4i[4] = {0, 1, 2, 3};
#pragma unroll
  for (int j = 0; j < MAXC * 16; j += 4)
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (v[j + k] > m4[k]) { m4[k] = v[j + k]; b4i[k] = j + k; }
  float mloc = m4[0];
  int bloc = b4i[0];
#pragma unroll
  for (int k = 1; k < 4; k++)
    if (m4[k] > mloc || (m4[k] == mloc && b4i[k] < bloc)) { mloc = m4[k]; bloc = b4i[k]; }
  bloc += c_lo * 16;
  const int tgt = row_ok ? p.ce_target[grow] : 0;
  const int tl = (tgt >= c_lo * 16 && tgt < c_hi * 16) ? tgt - c_lo * 16 : -1;   // target index inside this warp's slice
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
  float lt = 0.f;
#pragma unroll
  for (int j = 0; j < MAXC * 16; j += 4)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (j + k == tl) lt = v[j + k];
      v[j + k] = __expf(v[j + k] - mloc);                      // exp(-inf) = 0 for padding columns
      s4[k] += v[j + k];
    }
  const float sloc = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  // exchange with the other warp of this lane quarter
  ce_x[q * 2 + half][lane] = make_float4(mloc, sloc, lt, __int_as_float(bloc));
  asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
  const float4 o = ce_x[q * 2 + (half ^ 1)][lane];
  asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");    // slot reusable for the next tile
  const float m = fmaxf(mloc, o.x);
  const float fl = __expf(mloc - m), fo = __expf(o.x - m);
  const float ssum = sloc * fl + o.y * fo;
  if (!row_ok) return;
  if (half == 0) {
    ce_acc += (m + __logf(ssum)) - (lt + o.z);
    const int t = (int)(grow / p.ce_B);
    if (p.ce_fb[t]) {                                          // decoder.py:185: next input = argmax(logits_t), lowest index on ties
      const int bo = __float_as_int(o.w);
      p.ce_tok[grow + p.ce_B] = (o.x > mloc) ? bo : bloc;
    }
  }
  const float inv = p.ce_scale * fl / ssum;
  bf16* drow = p.ce_dl + grow * p.ce_ldl + c_lo * 16;
#pragma unroll
  for (int i = 0; i < MAXC; i++) {
    const int ch = c_lo + i;
    if (ch < c_hi) {
#pragma unroll
      for (int h8 = 0; h8 < 2; h8++) {
        if (ch * 16 + h8 * 8 < p.ce_ldl) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int e0 = i * 16 + h8 * 8 + 2 * j;
            const float g0 = v[e0] * inv - (e0 == tl ? p.ce_scale : 0.f);
            const float g1 = v[e0 + 1] * inv - (e0 + 1 == tl ? p.ce_scale : 0.f);
            __nv_bfloat162 tt = __floats2bfloat162_rn(g0, g1);
            pk[j] = *reinterpret_cast<uint32_t*>(&tt);
          }
          *reinterpret_cast<uint4*>(drow + i * 16 + h8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
  }
}

// MODE 1: the instantiation whose epilogue is the fused encoder LSTM step (TC_EPI_LSTM_FWD); kept apart so that its
// register footprint does not touch the code generation of the other epilogues
// MODE 2: fc_out with the cross-entropy / d logits / greedy feedback in the epilogue (TC_EPI_CE)
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BMt = p.bm2 ? 2 * TC_BM : TC_BM;
  const uint32_t a_bytes = (uint32_t)BMt * TC_BK * 2;       // 16 KB (32 KB for 256-row tiles)
  const uint32_t b_bytes = (uint32_t)p.BN * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  TcShared* sh = reinterpret_cast<TcShared*>(smem + (size_t)p.stages * stage_bytes);
  uint8_t* scratch = reinterpret_cast<uint8_t*>(sh) + 256;      // TC_EPI_WARPS x 32 rows x p.scr_pitch when p.use_scratch

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.mt * p.nt * p.splitk;
  const int kblocks = (p.K + TC_BK - 1) / TC_BK;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.BN) tmem_cols <<= 1;
  if (p.bm2 || p.fs) tmem_cols = 512;                      // two 256-column accumulators, one per row half / per-segment accumulators

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; s++) {
      tc::mbar_init(&sh->full[s], 1);
      tc::mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      tc::mbar_init(&sh->tmem_full[a], 1);
      tc::mbar_init(&sh->tmem_empty[a], TC_EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&sh->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // N-fastest tile order: the nt column tiles of one row tile run on neighbouring CTAs at the same time, so the
        // A tile is fetched from HBM once and re-read from L2 (M-fastest streamed A nt times: ncu, r01a)
        const int ni = tile % p.nt;
        const int mi = (tile / p.nt) % p.mt;
        const int ks = tile / (p.mt * p.nt);
        const int m0 = mi * BMt, n0 = ni * p.BN;
        const int gm0 = (int)p.rm(m0);
        const int kb0 = ks * p.kb_per;
        const int kb1 = min(kblocks, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; kb++) {
          tc::mbar_wait(&sh->empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          const int bn_t = p.fs ? p.segN[0] + p.segN[1] : (p.nseg > 1 ? p.segN[ni] : p.BN);
          tc::mbar_expect_tx(&sh->full[stage], a_bytes + (uint32_t)bn_t * TC_BK * 2);
          const int k0 = kb * TC_BK;
          if (!p.a_mn) {
            tc::tma_load_2d(sa, &tmA, &sh->full[stage], k0, gm0);                       // box {64 k, 128 rows}
            if (p.bm2) tc::tma_load_2d(sa + 16384, &tmA, &sh->full[stage], k0, gm0 + TC_BM);   // rows 128..255 of the tile
          } else {
            for (int j = 0; j < BMt / 64; j++)                                          // box {64 m, 64 k} x 2 (x 4)
              tc::tma_load_2d(sa + j * 8192, &tmA, &sh->full[stage], m0 + j * 64, k0);
          }
          if (!p.b_mn) {
            tc::tma_load_2d(sb, &tmB, &sh->full[stage], k0, n0);                        // box {64 k, BN rows}
          } else if (p.fs) {
            for (int j = 0; j < p.segN[0] / 64; j++)
              tc::tma_load_2d(sb + j * 8192, &tmB, &sh->full[stage], j * 64, k0 - p.seg_shift[0]);
            for (int j = 0; j < p.segN[1] / 64; j++)
              tc::tma_load_2d(sb + (p.segN[0] / 64 + j) * 8192, &tmB1, &sh->full[stage], j * 64, k0 - p.seg_shift[1]);
          } else if (p.nseg > 1) {
            const CUtensorMap* tb = ni == 0 ? &tmB : (ni == 1 ? &tmB1 : &tmB2);
            for (int j = 0; j < bn_t / 64; j++)
              tc::tma_load_2d(sb + j * 8192, tb, &sh->full[stage], j * 64, k0 - p.seg_shift[ni]);
          } else {
            for (int j = 0; j < p.BN / 64; j++)
              tc::tma_load_2d(sb + j * 8192, &tmB, &sh->full[stage], n0 + j * 64, k0);  // box {64 n, 64 k}
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const uint32_t idesc = tc::make_idesc_bf16(TC_BM, p.fs ? p.segN[0] : (p.nseg > 1 ? p.segN[tile % p.nt] : p.BN), p.a_mn != 0, p.b_mn != 0);
      const uint32_t idesc1 = tc::make_idesc_bf16(TC_BM, p.fs ? p.segN[1] : 64, p.a_mn != 0, p.b_mn != 0);
      const int ks = tile / (p.mt * p.nt);
      const int kb0 = ks * p.kb_per;
      const int kb1 = min(kblocks, kb0 + p.kb_per);
      const bool one_acc = p.bm2 || p.fs;                    // these tiles own all of TMEM: no accumulator double-buffering
      const int acc = one_acc ? 0 : (it & 1);
      const uint32_t acc_phase = one_acc ? (it & 1) : ((it >> 1) & 1);
      if (lane == 0) tc::mbar_wait(&sh->tmem_empty[acc], acc_phase ^ 1);
      __syncwarp();
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      for (int kb = kb0; kb < kb1; kb++) {
        if (lane == 0) tc::mbar_wait(&sh->full[stage], phase);
        __syncwarp();
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            const uint64_t da = p.a_mn ? tc::make_smem_desc(sa + k * 2048, 8192, 1024)
                                       : tc::make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = p.b_mn ? tc::make_smem_desc(sb + k * 2048, 8192, 1024)
                                       : tc::make_smem_desc(sb + k * 32, 16, 1024);
            tc::mma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (p.fs)                                             // the same A tile against the second segment
              tc::mma_bf16(tmem_base + (uint32_t)p.segN[0], da,
                           tc::make_smem_desc(sb + (p.segN[0] / 64) * 8192 + k * 2048, 8192, 1024), idesc1, (kb > kb0 || k > 0) ? 1u : 0u);
            if (p.bm2)                                            // rows 128..255 of the tile against the same B
              tc::mma_bf16(tmem_base + 256,
                           p.a_mn ? tc::make_smem_desc(sa + 16384 + k * 2048, 8192, 1024)
                                  : tc::make_smem_desc(sa + 16384 + k * 32, 16, 1024),
                           db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(&sh->empty[stage]);                  // smem stage free once these MMAs have read it
          if (kb == kb1 - 1) tc::mma_commit(&sh->tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9 -> TMEM lane quarters 2,3,0,1,2,3,0,1)
    constexpr bool LSTM = (MODE == 1);
    __shared__ float4 ce_x_storage[MODE == 2 ? 8 : 1][32];
    float4 (*ce_x)[32] = ce_x_storage;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;          // which half of the tile's 16-column chunks this warp handles
    int it = 0;
    float ce_acc = 0.f;                        // MODE 2: this thread's cross-entropy sum
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
      const int mi = (tile / p.nt) % p.mt;
      const bool mseg = p.nseg > 1;
      const bool one_acc = p.bm2 || p.fs;
      const int acc = one_acc ? 0 : (it & 1);
      const uint32_t acc_phase = one_acc ? (it & 1) : ((it >> 1) & 1);
      tc::mbar_wait(&sh->tmem_full[acc], acc_phase);
      tc::tc_fence_after();
      for (int fsi = 0; fsi < (p.fs ? 2 : 1); fsi++) {         // fs: the two segments' accumulators sit side by side
      const int ni = p.fs ? fsi : tile % p.nt;
      const uint32_t fs_col = p.fs ? (uint32_t)(fsi * p.segN[0]) : 0u;
      const int nchunks = (mseg ? p.segN[ni] : p.BN) >> 4;
      const int ch_lo = half == 0 ? 0 : (nchunks + 1) >> 1;
      const int ch_hi = half == 0 ? (nchunks + 1) >> 1 : nchunks;
      const int n0 = mseg ? 0 : ni * p.BN;
      float* const Ct = mseg ? p.segC[ni] : (p.split_stride > 0 ? p.C + (long)(tile / (p.mt * p.nt)) * p.split_stride : p.C);
      const int ldct = mseg ? p.seg_ldc[ni] : p.ldc;
      const int Nt = mseg ? p.segN[ni] : p.N;
      for (int sub = 0; sub < (p.bm2 ? 2 : 1); sub++) {      // row halves of a 256-row tile (plain epilogue only)
      const int m0 = mi * BMt + sub * TC_BM;
      const int row_local = m0 + q * 32 + lane;
      const bool row_ok = row_local < p.M;
      const long grow = row_ok ? p.rm(row_local) : 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.bm2 ? sub * 256 : acc * p.BN) + fs_col;
      const bool atomic = p.splitk > 1 && p.split_stride == 0;
      const bool add_bias = p.bias != nullptr && (tile / (p.mt * p.nt)) == 0;
      const int TC_SCR_PITCH = p.scr_pitch;
      uint8_t* scr = scratch + (warp - 2) * 32 * TC_SCR_PITCH;
      uint8_t* srow = scr + lane * TC_SCR_PITCH;
      const int rows_here = min(32, p.M - (m0 + q * 32));          // rows of this warp's quarter inside the matrix
      const long grow0 = rows_here > 0 ? p.rm(m0 + q * 32) : 0;     // first global row (a row tile never straddles timesteps)
      if (MODE == 2) {
        // a row of V <= 128 logits sits in one TMEM lane: the two warps of a lane quarter split its 16-column chunks
        // (warp `half` 0 takes the first ceil(n/2) chunks), reduce max / sum-exp / target logit / argmax locally and
        // combine through 16 bytes of shared memory per row
        const int nchv = (p.N + 15) >> 4;
        switch (nchv) {
          case 5: ce_epilogue<5>(p, taddr, half, q, lane, row_ok, grow, ce_x, ce_acc); break;      // V = 80 (default vocabulary)
          default: ce_epilogue<0>(p, taddr, half, q, lane, row_ok, grow, ce_x, ce_acc); break;
        }
      } else if (LSTM) {
        // one encoder LSTM step (MLX nn.LSTM loop body, models/encoder.py:98-101).  Accumulator columns of this 256-wide
        // tile: [0,64) = i, [64,128) = f, [128,192) = g, [192,256) = o of hidden units 64*ni .. 64*ni+63 (tile-permuted
        // Wh); this warp: units half*32 .. +32 in two chunks of 16.  Thread = batch row: P_t, c_{t-1} in, gates / c_t / h_t
        // out, all in natural column order.
        const int H = p.Hh;
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
          const int cu = half * 32 + cc * 16;                    // unit offset inside the 64-unit block
          const int u0 = ni * 64 + cu;                           // natural hidden-unit index of the chunk
          uint32_t ra[4][16];
#pragma unroll
          for (int gi = 0; gi < 4; gi++) tc::tmem_ld16(taddr + gi * 64 + cu, ra[gi]);
          uint4 pv[4][2];
          float4 cv[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) cv[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok) {
            const bf16* prow = p.pre_b + grow * 4L * H + u0;
#pragma unroll
            for (int gi = 0; gi < 4; gi++) {
              pv[gi][0] = __ldg(reinterpret_cast<const uint4*>(prow + (long)gi * H));
              pv[gi][1] = __ldg(reinterpret_cast<const uint4*>(prow + (long)gi * H) + 1);
            }
            if (p.c_prev != nullptr) {
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) cv[k4] = __ldg(reinterpret_cast<const float4*>(p.c_prev + grow * H + u0) + k4);
            }
          }
          tc::tmem_ld_wait();
          if (row_ok) {
            float gi_[16], gf_[16], gg_[16], go_[16], cn[16], hv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
              // bf16 element k of the 16-wide segment: 32-bit word k/2 of the two 16-byte registers, low / high half
              const uint4 w0 = pv[0][k >> 3], w1 = pv[1][k >> 3], w2 = pv[2][k >> 3], w3 = pv[3][k >> 3];
              const int wi = (k >> 1) & 3;
              const uint32_t x0 = wi == 0 ? w0.x : wi == 1 ? w0.y : wi == 2 ? w0.z : w0.w;
              const uint32_t x1 = wi == 0 ? w1.x : wi == 1 ? w1.y : wi == 2 ? w1.z : w1.w;
              const uint32_t x2 = wi == 0 ? w2.x : wi == 1 ? w2.y : wi == 2 ? w2.z : w2.w;
              const uint32_t x3 = wi == 0 ? w3.x : wi == 1 ? w3.y : wi == 2 ? w3.z : w3.w;
              const float p0 = __uint_as_float((k & 1) ? (x0 & 0xffff0000u) : (x0 << 16));
              const float p1 = __uint_as_float((k & 1) ? (x1 & 0xffff0000u) : (x1 << 16));
              const float p2 = __uint_as_float((k & 1) ? (x2 & 0xffff0000u) : (x2 << 16));
              const float p3 = __uint_as_float((k & 1) ? (x3 & 0xffff0000u) : (x3 << 16));
              const float4 c4 = cv[k >> 2];
              const float cp = (k & 3) == 0 ? c4.x : (k & 3) == 1 ? c4.y : (k & 3) == 2 ? c4.z : c4.w;
              gi_[k] = sigmoid_fast_(__uint_as_float(ra[0][k]) + p0);
              gf_[k] = sigmoid_fast_(__uint_as_float(ra[1][k]) + p1);
              gg_[k] = tanh_fast_(__uint_as_float(ra[2][k]) + p2);
              go_[k] = sigmoid_fast_(__uint_as_float(ra[3][k]) + p3);
              cn[k] = fmaf(gf_[k], cp, gi_[k] * gg_[k]);
              hv[k] = go_[k] * tanh_fast_(cn[k]);
            }
            bf16* grow_g = p.gates_b + grow * 4L * H + u0;
            st_bf16x16(grow_g, gi_);
            st_bf16x16(grow_g + (long)H, gf_);
            st_bf16x16(grow_g + 2L * H, gg_);
            st_bf16x16(grow_g + 3L * H, go_);
            st_bf16x16(p.hb_out + grow * H + u0, hv);
            float4* cdst = reinterpret_cast<float4*>(p.c_out + grow * H + u0);
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) cdst[k4] = make_float4(cn[4 * k4], cn[4 * k4 + 1], cn[4 * k4 + 2], cn[4 * k4 + 3]);
            if (p.hf_out != nullptr) {
              float4* hdst = reinterpret_cast<float4*>(p.hf_out + grow * H + u0);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) hdst[k4] = make_float4(hv[4 * k4], hv[4 * k4 + 1], hv[4 * k4 + 2], hv[4 * k4 + 3]);
            }
          }
        }
      } else if (p.epi == TC_EPI_DEC_CELL_FWD) {
        // accumulator columns of this 192-wide tile: [0,64) = i, [64,128) = g, [128,192) = o of units 64*ni .. 64*ni+63;
        // this warp: units half*32 .. +32.  scratch row = [h 64 B | i 64 B | g 64 B | o 64 B]
        const int H = p.Hh;
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
          const int c0 = half * 32 + cc * 16;
          uint32_t ri[16], rg[16], ro[16];
          tc::tmem_ld16(taddr + c0, ri);
          tc::tmem_ld16(taddr + 64 + c0, rg);
          tc::tmem_ld16(taddr + 128 + c0, ro);
          float4 bv[3][4];                                      // bias of the 16 units x (i,g,o), in flight with the TMEM loads
#pragma unroll
          for (int g3 = 0; g3 < 3; g3++)
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) bv[g3][k4] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g3 * 64 + c0) + k4);
          tc::tmem_ld_wait();
          const float* bfi = reinterpret_cast<const float*>(bv[0]);
          const float* bfg = reinterpret_cast<const float*>(bv[1]);
          const float* bfo = reinterpret_cast<const float*>(bv[2]);
          float gi[16], gg[16], go[16], hv[16];
#pragma unroll
          for (int k = 0; k < 16; k++) {
            gi[k] = sigmoid_fast_(__uint_as_float(ri[k]) + bfi[k]);
            gg[k] = tanh_fast_(__uint_as_float(rg[k]) + bfg[k]);
            go[k] = sigmoid_fast_(__uint_as_float(ro[k]) + bfo[k]);
            hv[k] = go[k] * tanh_fast_(gi[k] * gg[k]);
          }
          scr_put16(srow + cc * 32, hv);
          scr_put16(srow + 64 + cc * 32, gi);
          scr_put16(srow + 128 + cc * 32, gg);
          scr_put16(srow + 192 + cc * 32, go);
        }
        __syncwarp();
        if (rows_here > 0) {
          scr_store_rows(scr, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(p.hb_out + grow0 * H + ni * 64 + half * 32), (long)H * 2, 64, rows_here, lane);
          bf16* gb = p.gates_b + grow0 * 3L * H + n0 + half * 32;
          scr_store_rows(scr + 64, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(gb), 3L * H * 2, 64, rows_here, lane);
          scr_store_rows(scr + 128, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(gb + 64), 3L * H * 2, 64, rows_here, lane);
          scr_store_rows(scr + 192, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(gb + 128), 3L * H * 2, 64, rows_here, lane);
        }
        __syncwarp();
      } else if (p.epi == TC_EPI_DEC_CELL_BWD) {
        // accumulator = d h of units n0 + c; per 64-unit block the saved gates of a row are 384 contiguous bytes
        // [i 64 | g 64 | o 64] and the pre-activation gradients go out in the same layout
        const int H = p.Hh;
        const int nblocks = p.BN >> 6;                          // whole 64-unit blocks are split between the two warps of a quarter
        const int blk_lo = half == 0 ? 0 : (nblocks + 1) >> 1;
        const int blk_hi = half == 0 ? (nblocks + 1) >> 1 : nblocks;
#pragma unroll 1
        for (int blk = blk_lo; blk < blk_hi; blk++) {
          const int nblk = n0 + blk * 64;                       // first unit of the block
          const long goff = grow0 * 3L * H + (long)(nblk / 64) * 192;
          if (rows_here > 0)
            scr_load_rows(scr, TC_SCR_PITCH, reinterpret_cast<const uint8_t*>(p.gates_b + goff), 3L * H * 2, 384, rows_here, lane);
          __syncwarp();
#pragma unroll 1
          for (int cc = 0; cc < 4; cc++) {
            const int ch = blk * 4 + cc;
            uint32_t r[16];
            tc::tmem_ld16(taddr + ch * 16, r);
            tc::tmem_ld_wait();
            float gi[16], gg[16], go[16], dai[16], dag[16], dao[16];
            scr_get16(srow + cc * 32, gi);
            scr_get16(srow + 128 + cc * 32, gg);
            scr_get16(srow + 256 + cc * 32, go);
#pragma unroll
            for (int k = 0; k < 16; k++) dec_cell_grads_fast(gi[k], gg[k], go[k], __uint_as_float(r[k]), dai[k], dag[k], dao[k]);
            scr_put16(srow + cc * 32, dai);
            scr_put16(srow + 128 + cc * 32, dag);
            scr_put16(srow + 256 + cc * 32, dao);
          }
          __syncwarp();
          if (rows_here > 0)
            scr_store_rows(scr, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(p.dg_out + goff), 3L * H * 2, 384, rows_here, lane);
          __syncwarp();
        }
      } else if (p.epi == TC_EPI_LSTM_DH) {
        // d h for the cluster BPTT of the layer below (d X = dA @ Wx): bf16, written in the layout that kernel's epilogue
        // thread (batch row x 32 hidden units) reads — [t][tile][cta = unit/64][lane quarter][unit half][chunk 4][lane] x 16 B —
        // so thread = row stores 16-byte chunks and a warp covers 512 consecutive bytes (no scratch, no fp32 round trip)
        {
          // (tcgen05.ld is warp-collective: every lane runs the loop, rows beyond M only skip the stores)
          const int tt = (int)(grow / p.lp_B), bb = (int)(grow - (long)tt * p.lp_B);
          const long slots_per_t = (long)((p.lp_B + 127) >> 7) * 32;
          const int tile_b = bb >> 7, qq = (bb >> 5) & 3, ln = bb & 31;
          uint4* const base = reinterpret_cast<uint4*>(p.Cb) + (long)tt * slots_per_t * 128 + ln;
#pragma unroll 1
          for (int ch = ch_lo; ch < ch_hi; ch++) {
            uint32_t r[16];
            tc::tmem_ld16(taddr + ch * 16, r);
            tc::tmem_ld_wait();
            const int c0 = n0 + ch * 16;                         // hidden unit of the chunk's first column
            const long warp_slot = (((long)tile_b * 4 + (c0 >> 6)) * 4 + qq) * 2 + ((c0 >> 5) & 1);
            uint4* dst = base + (warp_slot * 4 + ((c0 >> 3) & 3)) * 32;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
              pk[j] = *reinterpret_cast<uint32_t*>(&t2);
            }
            if (row_ok) {
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[32] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      } else if (p.use_scratch) {
        // plain bf16 output, full tiles: this warp's half of the tile's columns, staged and written along the rows
        const int nb = (ch_hi - ch_lo) * 32;                    // bytes per row
#pragma unroll 1
        for (int ch = ch_lo; ch < ch_hi; ch++) {
          uint32_t r[16];
          tc::tmem_ld16(taddr + ch * 16, r);
          float4 bv[4];                                         // bias vector of the chunk, in flight with the TMEM load
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++)
            bv[k4] = add_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + ch * 16) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
          tc::tmem_ld_wait();
          const float* bf = reinterpret_cast<const float*>(bv);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]) + bf[j];
          scr_put16(srow + (ch - ch_lo) * 32, v);
        }
        __syncwarp();
        if (rows_here > 0)
          scr_store_rows(scr, TC_SCR_PITCH, reinterpret_cast<uint8_t*>(p.Cb + grow0 * p.ldcb + n0 + ch_lo * 16), (long)p.ldcb * 2, nb, rows_here, lane);
        __syncwarp();
      } else {
      for (int c0 = ch_lo * 16; c0 < ch_hi * 16; c0 += 16) {
        uint32_t r[16];
        tc::tmem_ld16(taddr + c0, r);
        tc::tmem_ld_wait();
        if (row_ok) {
          const int nbase = n0 + c0;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            v[j] = __uint_as_float(r[j]);
            if (add_bias && nbase + j < Nt) v[j] += p.bias[nbase + j];
          }
          if (Ct != nullptr) {
            float* dst = Ct + grow * ldct + nbase;
            if (atomic) {
#pragma unroll
              for (int j = 0; j < 16; j++)
                if (nbase + j < Nt) atomicAdd(dst + j, v[j]);
            } else if (nbase + 16 <= Nt && ((ldct & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.accumulate) {
                  float4 old = *reinterpret_cast<float4*>(dst + j);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(dst + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j++)
                if (nbase + j < Nt) dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
            }
          }
          if (p.Cb != nullptr) {
            bf16* dstb = p.Cb + grow * p.ldcb + nbase;
            if (nbase + 16 <= Nt && ((p.ldcb & 7) == 0) && ((reinterpret_cast<uintptr_t>(dstb) & 15) == 0)) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; j++) {
                __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&t);
              }
              *reinterpret_cast<uint4*>(dstb) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dstb + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; j++)
                if (nbase + j < Nt) dstb[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      }
      }   // sub
      }   // fsi
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[acc]);
    }
    if (MODE == 2) {
      const float w = warp_sum(ce_acc);
      if (lane == 0 && half == 0 && w != 0.f) atomicAdd(p.ce_sum, (double)w);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode() {
  static std::once_flag once;
  static int rc = 0;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      rc = 3;
    } else {
      g_encode = (PFN_encodeTiled)fn;
    }
  });
  return rc;
}

// row-major bf16 matrix [rows, cols] with pitch ld (elements); box {box_cols (inner), box_rows}
int make_tmap_bf16(CUtensorMap* m, const bf16* ptr, long rows, long cols, long ld, int box_cols, int box_rows) {
  ARCVAE_TRY(get_encode());
  ARCVAE_REQUIRE((ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0),
                 "TMA operands need 16-byte aligned base and pitch (ld multiple of 8 bf16)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 3;
  }
  return 0;
}

// time-major bf16 tensor [T][rows_per_t][cols] with row pitch ld (elements); box {box_cols, box_rows, 1}.  Rows beyond
// rows_per_t are clipped by the TMA unit, so a row tile never spills into the next timestep (ragged last tile).
int make_tmap_bf16_3d(CUtensorMap* m, const bf16* ptr, long T, long rows_per_t, long cols, long ld, int box_cols, int box_rows) {
  ARCVAE_TRY(get_encode());
  ARCVAE_REQUIRE((ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0),
                 "TMA operands need 16-byte aligned base and pitch (ld multiple of 8 bf16)");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_per_t, (cuuint64_t)T};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(bf16), (cuuint64_t)ld * rows_per_t * sizeof(bf16)};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r));
    return 3;
  }
  return 0;
}

static int pick_bn(int N) {
  if (N >= 256) return 256;
  return ((N + 15) / 16) * 16;
}

int pick_splitk_tc(int M, int N, int K);

int gemm_tc(const TcGemm& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  ARCVAE_REQUIRE(g.C != nullptr || g.Cb != nullptr || g.epi != TC_EPI_PLAIN || g.nseg > 1, "gemm_tc needs an output");
  ARCVAE_REQUIRE(g.epi != TC_EPI_LSTM_FWD || g.splitk <= 1, "fused LSTM step: no split-K");
  if (gemm_ws_supported(g) && std::getenv("ARCVAE_NO_WS") == nullptr) return gemm_ws(g, st);
  ARCVAE_REQUIRE(!(g.a_mn && g.rm.tlist != nullptr), "row map needs a K-major A");
  ARCVAE_REQUIRE(g.rm.tlist == nullptr || (g.rm.Bt % TC_BM) == 0, "row-mapped tiles must not straddle timesteps");
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.nseg = g.nseg;
  for (int i = 0; i < 3; i++) { p.segN[i] = 0; p.seg_shift[i] = 0; p.segC[i] = nullptr; p.seg_ldc[i] = 0; }
  if (g.nseg > 1) {
    ARCVAE_REQUIRE(g.nseg <= 3 && g.a_mn && g.b_mn && g.accumulate && g.epi == TC_EPI_PLAIN && g.Cb == nullptr &&
                   g.rm.tlist == nullptr, "multi-segment GEMM: MN-major operands, fp32 accumulate outputs");
    int nmax = 0;
    for (int i = 0; i < g.nseg; i++) {
      ARCVAE_REQUIRE(g.seg[i].N % 64 == 0 && g.seg[i].N <= 256 && g.seg[i].C != nullptr, "segment N: multiple of 64, <= 256");
      p.segN[i] = g.seg[i].N; p.seg_shift[i] = g.seg[i].k_shift; p.segC[i] = g.seg[i].C; p.seg_ldc[i] = g.seg[i].ldc;
      if (g.seg[i].N > nmax) nmax = g.seg[i].N;
    }
    p.N = nmax;
  }
  p.BN = pick_bn(p.N);
  if (g.b_mn) {
    ARCVAE_REQUIRE(p.N % 64 == 0 || p.N < 64, "MN-major B needs N multiple of 64");
    p.BN = p.N >= 256 ? 256 : ((p.N + 63) / 64) * 64;
  }
  p.a_mn = g.a_mn ? 1 : 0; p.b_mn = g.b_mn ? 1 : 0;
  // long-K weight-gradient shapes (both operands MN-major, fp32 atomics): 256-row tiles
  p.bm2 = (g.a_mn && g.b_mn && g.epi == TC_EPI_PLAIN && g.Cb == nullptr && g.accumulate && (g.M % 256) == 0 &&
           g.K >= 64 * 64 && std::getenv("ARCVAE_NO_BM256") == nullptr) ? 1 : 0;
  // Two-segment weight gradients (dA^T [h | onehot]): as separate column tiles the 128-column one-hot segment runs ~35 %
  // faster than its 256-column sibling, drifts out of the sibling's L2 window and the shared A operand comes from HBM
  // 1.5x (ncu r02e: 2.27 GB for 1.47 GB algorithmic).  One CTA per 128-row tile computes BOTH segments from one A tile:
  // every CTA does the same work, A is read exactly once, 384 TMEM columns.
  p.fs = (g.nseg == 2 && g.a_mn && g.b_mn && g.seg[0].N + g.seg[1].N <= 512 && std::getenv("ARCVAE_NO_FUSED_SEGMENTS") == nullptr) ? 1 : 0;
  if (p.fs) { p.bm2 = 0; p.BN = g.seg[0].N + g.seg[1].N; }
  p.mt = cdiv(g.M, p.bm2 ? 2 * TC_BM : TC_BM); p.nt = p.fs ? 1 : (g.nseg > 1 ? g.nseg : cdiv(g.N, p.BN));
  const int kblocks = cdiv(g.K, TC_BK);
  int splitk = g.splitk;
  if (splitk <= 0) {                        // auto: (tiles x splits) fills two waves of the 148 SMs
    long sk = (2L * 148) / ((long)p.mt * p.nt);
    const long maxs = kblocks / 8;
    if (sk > maxs) sk = maxs;
    splitk = sk < 1 ? 1 : (int)sk;
  } else if (p.bm2 && g.a_mn && splitk > 1) {
    splitk *= 2;                            // the caller sized the split for 128-row tiles
  } else if (p.fs && splitk > 1) {
    splitk *= 2;                            // the caller counted one tile per segment
  }
  if (splitk > kblocks) splitk = kblocks;
  p.kb_per = cdiv(kblocks, splitk);
  p.splitk = cdiv(kblocks, p.kb_per);
  ARCVAE_REQUIRE(p.splitk == 1 || (g.accumulate && (g.C != nullptr || g.nseg > 1) && g.Cb == nullptr) ||
                     (g.split_stride > 0 && !g.accumulate && g.C != nullptr && g.Cb == nullptr && g.nseg <= 1 && g.bias == nullptr),
                 "split-K accumulates with fp32 atomics into C, or stores partials (split_stride)");
  p.split_stride = p.splitk > 1 ? g.split_stride : 0;
  p.C = g.C; p.ldc = g.ldc; p.Cb = g.Cb; p.ldcb = g.ldcb; p.bias = g.bias; p.accumulate = g.accumulate ? 1 : 0;
  p.rm = g.rm;
  p.epi = g.epi; p.gates_b = g.gates_b; p.hb_out = g.hb_out; p.dg_out = g.dg_out;
  p.lp_B = g.lp_B;
  ARCVAE_REQUIRE(g.epi != TC_EPI_LSTM_DH || (g.lp_B > 0 && g.Cb != nullptr && g.N == 256 && g.splitk <= 1 && g.rm.tlist == nullptr),
                 "d h epilogue: N = 256, bf16 thread-friendly output");
  p.Hh = g.Hh;
  p.pre_b = g.pre_b; p.c_prev = g.c_prev; p.c_out = g.c_out; p.hf_out = g.hf_out;
  p.ce_target = g.ce_target; p.ce_fb = g.ce_fb; p.ce_tok = g.ce_tok; p.ce_B = g.ce_B; p.ce_dl = g.ce_dl; p.ce_ldl = g.ce_ldl;
  p.ce_sum = g.ce_sum; p.ce_scale = g.ce_scale;
  if (g.epi == TC_EPI_CE) {
    ARCVAE_REQUIRE(g.N <= 128 && !g.a_mn && !g.b_mn && p.splitk == 1 && g.nseg <= 1 && g.bias != nullptr,
                   "fused cross-entropy: one N tile (V <= 128), K-major operands, bias, no split-K");
    ARCVAE_REQUIRE(g.ce_target && g.ce_fb && g.ce_tok && g.ce_dl && g.ce_sum && g.ce_B > 0 && (g.ce_ldl % 8) == 0 &&
                   g.ce_ldl >= g.N && g.ce_ldl <= 128 && (reinterpret_cast<uintptr_t>(g.ce_dl) & 15) == 0,
                   "fused cross-entropy: outputs");
  }
  if (g.epi == TC_EPI_LSTM_FWD) {
    ARCVAE_REQUIRE(g.N == 4 * g.Hh && g.Hh % 64 == 0 && !g.a_mn && !g.b_mn && p.splitk == 1 && g.nseg <= 1 &&
                   g.rm.tlist == nullptr, "fused LSTM step: N = 4H tile-permuted, K-major operands, no split-K");
    ARCVAE_REQUIRE(g.pre_b != nullptr && g.c_out != nullptr && g.gates_b != nullptr && g.hb_out != nullptr, "fused LSTM step: outputs");
    ARCVAE_REQUIRE(((reinterpret_cast<uintptr_t>(g.pre_b) | reinterpret_cast<uintptr_t>(g.c_out) | reinterpret_cast<uintptr_t>(g.gates_b) |
                     reinterpret_cast<uintptr_t>(g.hb_out) | reinterpret_cast<uintptr_t>(g.c_prev) | reinterpret_cast<uintptr_t>(g.hf_out)) & 15) == 0,
                   "fused LSTM step: 16-byte aligned tensors");
    p.BN = 256; p.nt = g.N / 256;
  } else if (g.epi == TC_EPI_DEC_CELL_FWD) {
    ARCVAE_REQUIRE(g.N % 192 == 0 && g.Hh * 3 == g.N && !g.b_mn && g.bias != nullptr && g.gates_b && g.hb_out,
                   "fused decoder cell (forward): N = 3H tile-permuted, K-major B, bias");
    p.BN = 192; p.nt = g.N / 192;
  } else if (g.epi == TC_EPI_DEC_CELL_BWD) {
    ARCVAE_REQUIRE(g.N == g.Hh && g.Hh % 64 == 0 && g.dg_out != nullptr && p.splitk == 1, "fused decoder cell (backward): N = H");
    ARCVAE_REQUIRE(g.gates_b != nullptr, "saved gates");
  }
  const size_t stage_bytes = (size_t)(p.bm2 ? 2 : 1) * TC_BM * TC_BK * 2 + (size_t)p.BN * TC_BK * 2;
  // transposition scratch for the fused cells and for plain bf16-only outputs of full, aligned tiles
  const bool plain_fast = g.nseg <= 1 && g.epi == TC_EPI_PLAIN && g.Cb != nullptr && g.C == nullptr && p.splitk == 1 && (g.N % p.BN) == 0 &&
                          (p.BN % 32) == 0 && p.BN * 2 / 2 <= 384 && (g.ldcb % 8) == 0 &&
                          ((reinterpret_cast<uintptr_t>(g.Cb) & 15) == 0) &&
                          (g.bias == nullptr || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
  p.use_scratch = (g.epi == TC_EPI_DEC_CELL_FWD || g.epi == TC_EPI_DEC_CELL_BWD || plain_fast) ? 1 : 0;
  p.scr_pitch = g.epi == TC_EPI_DEC_CELL_BWD ? TC_SCR_PITCH_BWD : TC_SCR_PITCH_FWD;
  const size_t scratch_bytes = p.use_scratch ? (size_t)TC_EPI_WARPS * 32 * p.scr_pitch : 0;
  int stages = (int)(((g.epi == TC_EPI_CE ? 220 : 225) * 1024 - 2048 - scratch_bytes) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > p.kb_per + 1 && p.kb_per + 1 >= 2) stages = p.kb_per + 1 > 2 ? p.kb_per + 1 : 2;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 256 + scratch_bytes + 1024;
  ARCVAE_REQUIRE(stages >= 2, "gemm_tc: pipeline needs two stages");
  static_assert(sizeof(TcShared) <= 256, "TcShared must fit the 256-byte slot in front of the scratch");

  CUtensorMap tmA, tmB;
  if (!g.a_mn) {
    // rows of A may be reached through the row map -> describe the whole allocation the map can touch
    long rows = g.rm.tlist ? g.a_rows_total : g.M;
    ARCVAE_TRY(make_tmap_bf16(&tmA, g.A, rows, g.K, g.lda, TC_BK, TC_BM));
  } else {
    ARCVAE_TRY(make_tmap_bf16(&tmA, g.A, g.K, g.M, g.lda, 64, TC_BK));
  }
  CUtensorMap tmB1, tmB2;
  if (g.nseg > 1) {
    ARCVAE_TRY(make_tmap_bf16(&tmB, g.seg[0].B, g.K, g.seg[0].N, g.seg[0].ldb, 64, TC_BK));
    ARCVAE_TRY(make_tmap_bf16(&tmB1, g.seg[1].B, g.K, g.seg[1].N, g.seg[1].ldb, 64, TC_BK));
    if (g.nseg > 2) ARCVAE_TRY(make_tmap_bf16(&tmB2, g.seg[2].B, g.K, g.seg[2].N, g.seg[2].ldb, 64, TC_BK));
    else tmB2 = tmB1;
  } else {
    if (!g.b_mn) ARCVAE_TRY(make_tmap_bf16(&tmB, g.B, g.N, g.K, g.ldb, TC_BK, p.BN));
    else ARCVAE_TRY(make_tmap_bf16(&tmB, g.B, g.K, g.N, g.ldb, 64, TC_BK));
    tmB1 = tmB; tmB2 = tmB;
  }

  const int num_sms = device_sm_count();
  if (first_use_on_device(ONCE_GEMM_TC)) {
    ARCVAE_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ARCVAE_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ARCVAE_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));   // + 4 KB static
  }
  ARCVAE_REQUIRE(smem <= (size_t)(g.epi == TC_EPI_CE ? 222 : 227) * 1024, "gemm_tc shared memory budget");
  const int total = p.mt * p.nt * p.splitk;
  const int grid = total < num_sms ? total : num_sms;
  TimeScope ts(TIME_GEMM_TC, st);
  {
    double ncols = 0.0;
    if (g.nseg > 1) for (int i = 0; i < g.nseg; i++) ncols += g.seg[i].N;
    else ncols = g.N;
    count_flops(TIME_GEMM_TC, 2.0 * g.M * ncols * g.K);
  }
  if (g.epi == TC_EPI_LSTM_FWD) gemm_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmB1, tmB2, p);
  else if (g.epi == TC_EPI_CE) gemm_tc_kernel<2><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmB1, tmB2, p);
  else gemm_tc_kernel<0><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmB1, tmB2, p);
  ARCVAE_LAUNCHED();
  return 0;
}

// precision dispatch: tensor cores when asked for and the operands fit the TMA constraints, else fp32 FFMA tiles
static bool tma_ok(const void* p, int ld) { return p != nullptr && (ld % 8) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }

int gemm_any(int precision, int transA, int transB, int M, int N, int K, Mat A, Mat B, float* C, int ldc,
             const float* bias, bool accumulate, RowMap rm, long a_rows_total, cudaStream_t st) {
  ARCVAE_REQUIRE(!(transA && transB), "gemm_any: (T,T) not supported");
  bool tcok = precision == ARCVAE_PREC_BF16 && tma_ok(A.b, A.ld) && tma_ok(B.b, B.ld);
  const bool b_mn = (transB == 0);
  const bool a_mn = (transA != 0);
  if (tcok && b_mn && !(N % 64 == 0 || N < 64)) tcok = false;
  if (tcok && rm.tlist != nullptr && (rm.Bt % TC_BM) != 0) tcok = false;
  if (tcok) {
    TcGemm g{};
    g.M = M; g.N = N; g.K = K;
    g.A = A.b; g.lda = A.ld; g.a_mn = a_mn;
    g.B = B.b; g.ldb = B.ld; g.b_mn = b_mn;
    g.C = C; g.ldc = ldc; g.Cb = nullptr; g.ldcb = 0; g.bias = bias; g.accumulate = accumulate;
    g.splitk = accumulate ? pick_splitk_tc(M, N, K) : 1;
    g.rm = rm; g.a_rows_total = a_rows_total;
    return gemm_tc(g, st);
  }
  ARCVAE_REQUIRE(A.f != nullptr && B.f != nullptr, "gemm_any: fp32 operands missing for the fp32 path");
  int sk = accumulate ? pick_splitk(M, N, K) : 1;
  return gemm_f32(transA, transB, M, N, K, A.f, A.ld, B.f, B.ld, C, ldc, bias, accumulate, rm, sk, st);
}

// split-K factor for weight-gradient shapes: enough (tile, split) work items for ~2 per SM
int pick_splitk_tc(int M, int N, int K) {
  long tiles = (long)cdiv(M, TC_BM) * cdiv(N, pick_bn(N));
  long kblocks = cdiv(K, TC_BK);
  if (tiles >= 148 || kblocks < 16) return 1;
  long want = (2 * 148) / tiles;              // floor: (tiles x splits) <= two full waves, no straggler wave
  if (want < 1) want = 1;
  long maxs = kblocks / 8;
  if (maxs < 1) maxs = 1;
  return (int)(want < maxs ? want : maxs);
}

// ---- fp32 -> bf16 conversion (vectorised, HBM-bound) ------------------------------------------------------------
__global__ void k_f32_to_bf16(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  long n4 = n >> 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
int f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long g = (n / 4 + 255) / 256 + 1;
  if (g > 148 * 8) g = 148 * 8;
  TimeScope ts(TIME_POINTWISE, st);
  k_f32_to_bf16<<<(int)g, 256, 0, st>>>(src, dst, n);
  ARCVAE_LAUNCHED();
  return 0;
}
// rows of V floats -> rows of bf16 with pitch Vp >= V (pad columns zero): operands whose natural pitch is not a
// multiple of 16 bytes cannot be described to the TMA unit
__global__ void k_f32_to_bf16_pitched(const float* __restrict__ src, long R, int V, bf16* __restrict__ dst, int Vp) {
  const long total = R * Vp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / Vp;
    const int c = (int)(i - r * Vp);
    dst[i] = __float2bfloat16(c < V ? src[r * V + c] : 0.f);
  }
}
int f32_to_bf16_pitched(const float* src, long R, int V, bf16* dst, int Vp, cudaStream_t st) {
  if (Vp == V) return f32_to_bf16(src, dst, R * V, st);
  long g = (R * Vp + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  TimeScope ts(TIME_POINTWISE, st);
  k_f32_to_bf16_pitched<<<(int)g, 256, 0, st>>>(src, R, V, dst, Vp);
  ARCVAE_LAUNCHED();
  return 0;
}
// dst[c*R + r] = bf16(src[r*C + c])   (weights only: small)
__global__ void k_transpose_to_bf16(const float* __restrict__ src, int R, int C, bf16* __restrict__ dst) {
  __shared__ float tile[32][33];
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = src[(long)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(long)c * R + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
int transpose_to_bf16(const float* src, int R, int C, bf16* dst, cudaStream_t st) {
  k_transpose_to_bf16<<<dim3(cdiv(C, 32), cdiv(R, 32)), dim3(32, 8), 0, st>>>(src, R, C, dst);
  ARCVAE_LAUNCHED();
  return 0;
}

}  // namespace arcvae

// test entry: D = op(A) op(B) with bf16 operands (device pointers to bf16 data)
extern "C" int arcvae_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                                float* C, int ldc, void* Cb, int ldcb, const float* bias, int accumulate, int splitk,
                                void* stream) {
  using namespace arcvae;
  TcGemm g{};
  g.M = M; g.N = N; g.K = K;
  g.A = (const bf16*)A; g.lda = lda; g.a_mn = a_mn != 0;
  g.B = (const bf16*)B; g.ldb = ldb; g.b_mn = b_mn != 0;
  g.C = C; g.ldc = ldc; g.Cb = (bf16*)Cb; g.ldcb = ldcb; g.bias = bias; g.accumulate = accumulate != 0;
  g.splitk = splitk;
  g.rm = RowMap{nullptr, 1};
  g.a_rows_total = M;
  return gemm_tc(g, (cudaStream_t)stream);
}
