// gemm_ws.cu — WEIGHT-STATIONARY bf16 tensor-core GEMM for the time-parallel projections with a short contraction
// (K <= 256: the encoder's layer >= 1 input projection P = H @ Wx^T + b and the decoder's layer >= 1 projection with the
// zero-state cell in the epilogue; M = T*B = 524,288 rows at the benched configuration).
//
// Why: in gemm_tc_kernel every 128-row tile re-loads its [BN x K] weight panel from L2 (2/3 of the 3.1 GB of L2 -> SM
// traffic of the P projection, ncu r01b) and only 3 stages x 48 KB fit beside the per-warp transposition scratch, so
// the launches ran at 43-47 % of the HBM roofline that bounds them.  Here
//   * a CTA owns ONE column panel of the weights for its whole life: the panel (BN x K bf16 <= 128 KB) is loaded once by
//     TMA and stays in shared memory; only the 16 KB activation tiles stream through a 4-stage ring;
//   * the CTAs that share a row tile (one per panel) walk the row tiles in lock step, so the activations come from HBM
//     once and from L2 for the other panels;
//   * the epilogue writes bf16 results into SWIZZLE_128B staging slabs and ONE thread hands each [128 x 64] slab to the
//     TMA unit (cp.async.bulk.tensor store): no per-thread global stores, no transposition scratch.
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (tcgen05.mma, M=128, N=BN, K=16), warps 2..9 epilogue
// (two warps per TMEM lane quarter).  Two accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

constexpr int WS_BM = 128;
constexpr int WS_BK = 64;
constexpr int WS_STAGES = 4;
constexpr int WS_THREADS = 320;
constexpr int WS_SLAB = WS_BM * 64 * 2;          // one [128 rows x 64 cols] bf16 staging slab = one TMA store box
constexpr int WS_A_BYTES = WS_BM * WS_BK * 2;    // 16 KB

struct WsParams {
  int M, N, K, BN, nt, mt, kblocks;
  const float* bias;
  RowMap rm;
  int epi;        // WS_EPI_*
  int nslab;      // staging slabs
  int lp_B;       // WS_EPI_LSTM_P: rows per timestep
  int lp_vecs;    // WS_EPI_LSTM_P: vectors per thread slot
  uint4* lp_out;  // WS_EPI_LSTM_P: thread-friendly output
};
enum { WS_EPI_PLAIN = 0, WS_EPI_DEC_CELL_FWD = 1, WS_EPI_LSTM_P = 2 };

struct __align__(8) WsShared {
  uint64_t full[WS_STAGES];
  uint64_t empty[WS_STAGES];
  uint64_t w_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

namespace ws {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// 16 fp32 -> 16 bf16 into a SWIZZLE_128B slab: row r, 16-byte chunks c and c+1 (c even)
__device__ __forceinline__ void put16(uint8_t* slab, int r, int c, const float (&v)[16]) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  uint8_t* row = slab + r * 128;
  *reinterpret_cast<uint4*>(row + (((c) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(row + (((c + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
}  // namespace ws

__global__ void __launch_bounds__(WS_THREADS, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmG, const WsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t wk_bytes = (uint32_t)p.BN * WS_BK * 2;                 // one k-block of the weight panel
  uint8_t* wpanel = smem;                                                // kblocks x [BN x 64] bf16, K-major SWIZZLE_128B
  uint8_t* aring = wpanel + (size_t)p.kblocks * wk_bytes;                // WS_STAGES x 16 KB
  uint8_t* stage = aring + (size_t)WS_STAGES * WS_A_BYTES;               // nslab x 16 KB
  WsShared* sh = reinterpret_cast<WsShared*>(stage + (size_t)p.nslab * WS_SLAB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ni = blockIdx.x % p.nt;                  // this CTA's weight panel
  const int mi0 = blockIdx.x / p.nt;
  const int mstride = gridDim.x / p.nt;
  const int n0 = ni * p.BN;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmC);
    for (int s = 0; s < WS_STAGES; s++) {
      tc::mbar_init(&sh->full[s], 1);
      tc::mbar_init(&sh->empty[s], 1);
    }
    tc::mbar_init(&sh->w_full, 1);
    for (int a = 0; a < 2; a++) {
      tc::mbar_init(&sh->tmem_full[a], 1);
      tc::mbar_init(&sh->tmem_empty[a], 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&sh->tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      tc::mbar_expect_tx(&sh->w_full, (uint32_t)p.kblocks * wk_bytes);
      for (int kb = 0; kb < p.kblocks; kb++)
        tc::tma_load_2d(wpanel + (size_t)kb * wk_bytes, &tmB, &sh->w_full, kb * WS_BK, n0);     // box {64 k, BN rows}
      int st = 0;
      uint32_t phase = 0;
      for (int mi = mi0; mi < p.mt; mi += mstride) {
        const int gm0 = (int)p.rm(mi * WS_BM);
        for (int kb = 0; kb < p.kblocks; kb++) {
          tc::mbar_wait(&sh->empty[st], phase ^ 1);
          tc::mbar_expect_tx(&sh->full[st], WS_A_BYTES);
          tc::tma_load_2d(aring + (size_t)st * WS_A_BYTES, &tmA, &sh->full[st], kb * WS_BK, gm0);   // box {64 k, 128 rows}
          if (++st == WS_STAGES) { st = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint32_t idesc = tc::make_idesc_bf16(WS_BM, p.BN, false, false);
    int st = 0, it = 0;
    uint32_t phase = 0;
    if (lane == 0) tc::mbar_wait(&sh->w_full, 0);
    __syncwarp();
    for (int mi = mi0; mi < p.mt; mi += mstride, it++) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (lane == 0) tc::mbar_wait(&sh->tmem_empty[acc], acc_phase ^ 1);
      __syncwarp();
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < p.kblocks; kb++) {
        if (lane == 0) tc::mbar_wait(&sh->full[st], phase);
        __syncwarp();
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(aring + (size_t)st * WS_A_BYTES);
          const uint32_t sb = tc::smem_u32(wpanel + (size_t)kb * wk_bytes);
#pragma unroll
          for (int k = 0; k < WS_BK / 16; k++)
            tc::mma_bf16(d_tmem, tc::make_smem_desc(sa + k * 32, 16, 1024), tc::make_smem_desc(sb + k * 32, 16, 1024), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          tc::mma_commit(&sh->empty[st]);
          if (kb == p.kblocks - 1) tc::mma_commit(&sh->tmem_full[acc]);
        }
        __syncwarp();
        if (++st == WS_STAGES) { st = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue: warps 2..9 -> TMEM lane quarters 2,3,0,1,2,3,0,1
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                        // row inside the tile = TMEM lane
    const bool leader = (threadIdx.x == 64);            // issues the TMA stores and owns their bulk groups
    int it = 0;
    uint32_t slab_ctr = 0;
    for (int mi = mi0; mi < p.mt; mi += mstride, it++) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int gm0 = (int)p.rm(mi * WS_BM);
      tc::mbar_wait(&sh->tmem_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      if (p.epi == WS_EPI_PLAIN) {
        // BN/64 slabs of 64 columns, double-buffered staging: slab s may be written while the store of s-1 still reads
        const int nsl = p.BN >> 6;
#pragma unroll 1
        for (int s = 0; s < nsl; s++, slab_ctr++) {
          uint8_t* slab = stage + (size_t)(slab_ctr & 1) * WS_SLAB;
          if (leader) ws::bulk_wait_read<1>();          // the store that used this buffer two slabs ago has read it
          ws::epi_barrier();
          const int c0 = s * 64 + half * 32;
          uint32_t ra[16], rb[16];
          tc::tmem_ld16(taddr + c0, ra);
          tc::tmem_ld16(taddr + c0 + 16, rb);
          float4 bv[8];
#pragma unroll
          for (int k4 = 0; k4 < 8; k4++)
            bv[k4] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
          tc::tmem_ld_wait();
          float va[16], vb[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            va[4 * k4 + 0] = __uint_as_float(ra[4 * k4 + 0]) + bv[k4].x;
            va[4 * k4 + 1] = __uint_as_float(ra[4 * k4 + 1]) + bv[k4].y;
            va[4 * k4 + 2] = __uint_as_float(ra[4 * k4 + 2]) + bv[k4].z;
            va[4 * k4 + 3] = __uint_as_float(ra[4 * k4 + 3]) + bv[k4].w;
            vb[4 * k4 + 0] = __uint_as_float(rb[4 * k4 + 0]) + bv[4 + k4].x;
            vb[4 * k4 + 1] = __uint_as_float(rb[4 * k4 + 1]) + bv[4 + k4].y;
            vb[4 * k4 + 2] = __uint_as_float(rb[4 * k4 + 2]) + bv[4 + k4].z;
            vb[4 * k4 + 3] = __uint_as_float(rb[4 * k4 + 3]) + bv[4 + k4].w;
          }
          ws::put16(slab, r, half * 4, va);
          ws::put16(slab, r, half * 4 + 2, vb);
          tc::fence_proxy_async();
          ws::epi_barrier();
          if (leader) {
            ws::tma_store_2d(&tmC, slab, n0 + s * 64, gm0);         // rows beyond the tensor are clipped by the TMA unit
            ws::bulk_commit();
          }
        }
      } else if (p.epi == WS_EPI_LSTM_P) {
        // input projection of an encoder LSTM layer for the cluster recurrence: panel ni = gate ni (BN = H = 256), 64-column
        // block s = the units of cluster CTA s, `half` = which 32 of them.  Thread = row writes its 16-byte chunks straight
        // to the layout that kernel reads: the 32 lanes of a warp hit 512 consecutive bytes (no staging, no TMA).
        const long gr = (long)gm0 + r;
        const int tt = (int)(gr / p.lp_B), bb = (int)(gr - (long)tt * p.lp_B);
        const int ntl = (p.lp_B + 127) >> 7;
        const long slots_per_t = (long)ntl * 32;
        const int tile = bb >> 7, qq = (bb >> 5) & 3, ln = bb & 31;
        const bool rvalid = gr < p.M;
#pragma unroll 1
        for (int s = 0; s < 4; s++) {
          const int c0 = s * 64 + half * 32;
          uint32_t ra[16], rb[16];
          tc::tmem_ld16(taddr + c0, ra);
          tc::tmem_ld16(taddr + c0 + 16, rb);
          float4 bv[8];
#pragma unroll
          for (int k4 = 0; k4 < 8; k4++)
            bv[k4] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
          tc::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            pk[2 * k4] = pack_bf16x2(__uint_as_float(ra[4 * k4 + 0]) + bv[k4].x, __uint_as_float(ra[4 * k4 + 1]) + bv[k4].y);
            pk[2 * k4 + 1] = pack_bf16x2(__uint_as_float(ra[4 * k4 + 2]) + bv[k4].z, __uint_as_float(ra[4 * k4 + 3]) + bv[k4].w);
            pk[8 + 2 * k4] = pack_bf16x2(__uint_as_float(rb[4 * k4 + 0]) + bv[4 + k4].x, __uint_as_float(rb[4 * k4 + 1]) + bv[4 + k4].y);
            pk[8 + 2 * k4 + 1] = pack_bf16x2(__uint_as_float(rb[4 * k4 + 2]) + bv[4 + k4].z, __uint_as_float(rb[4 * k4 + 3]) + bv[4 + k4].w);
          }
          if (rvalid) {
            const long warp_slot = (((long)tile * 4 + s) * 4 + qq) * 2 + half;
            uint4* dst = p.lp_out + (((long)tt * slots_per_t + warp_slot) * p.lp_vecs + ni * 4) * 32 + ln;
#pragma unroll
            for (int cu = 0; cu < 4; cu++) dst[cu * 32] = make_uint4(pk[4 * cu], pk[4 * cu + 1], pk[4 * cu + 2], pk[4 * cu + 3]);
          }
        }
      } else {
        // decoder zero-state LSTM cell (models/decoder.py:165-168 with no state): accumulator columns [0,64) = i,
        // [64,128) = g, [128,192) = o of hidden units 64*ni .. 64*ni+63.  Slabs: 0 = h, 1 = i, 2 = g, 3 = o.
        if (leader) ws::bulk_wait_read<0>();
        ws::epi_barrier();
#pragma unroll 1
        for (int cc = 0; cc < 2; cc++) {
          const int c0 = half * 32 + cc * 16;
          uint32_t ri[16], rg[16], ro[16];
          tc::tmem_ld16(taddr + c0, ri);
          tc::tmem_ld16(taddr + 64 + c0, rg);
          tc::tmem_ld16(taddr + 128 + c0, ro);
          float4 bv[3][4];
#pragma unroll
          for (int g3 = 0; g3 < 3; g3++)
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) bv[g3][k4] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g3 * 64 + c0) + k4);
          tc::tmem_ld_wait();
          float gi[16], gg[16], go[16], hv[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const float bi[4] = {bv[0][k4].x, bv[0][k4].y, bv[0][k4].z, bv[0][k4].w};
            const float bg[4] = {bv[1][k4].x, bv[1][k4].y, bv[1][k4].z, bv[1][k4].w};
            const float bo[4] = {bv[2][k4].x, bv[2][k4].y, bv[2][k4].z, bv[2][k4].w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int k = 4 * k4 + j;
              gi[k] = sigmoid_approx_(__uint_as_float(ri[k]) + bi[j]);
              gg[k] = tanh_approx_(__uint_as_float(rg[k]) + bg[j]);
              go[k] = sigmoid_approx_(__uint_as_float(ro[k]) + bo[j]);
              hv[k] = go[k] * tanh_approx_(gi[k] * gg[k]);
            }
          }
          const int ch = half * 4 + cc * 2;
          ws::put16(stage, r, ch, hv);
          ws::put16(stage + WS_SLAB, r, ch, gi);
          ws::put16(stage + 2 * WS_SLAB, r, ch, gg);
          ws::put16(stage + 3 * WS_SLAB, r, ch, go);
        }
        tc::fence_proxy_async();
        ws::epi_barrier();
        if (leader) {
          ws::tma_store_2d(&tmC, stage, ni * 64, gm0);                       // h of this 64-unit block
          ws::tma_store_2d(&tmG, stage + WS_SLAB, n0, gm0);                  // activated gates, tile-permuted compact layout
          ws::tma_store_2d(&tmG, stage + 2 * WS_SLAB, n0 + 64, gm0);
          ws::tma_store_2d(&tmG, stage + 3 * WS_SLAB, n0 + 128, gm0);
          ws::bulk_commit();
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[acc]);
    }
    if (leader) ws::bulk_wait_all();                    // shared memory must outlive the last stores
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

int make_tmap_bf16(CUtensorMap* m, const bf16* ptr, long rows, long cols, long ld, int box_cols, int box_rows);

bool gemm_ws_supported(const TcGemm& g) {
  if (g.a_mn || g.b_mn || g.nseg > 1 || g.splitk > 1 || g.accumulate || g.C != nullptr) return false;
  if (g.K > 256 || (g.K % 64) != 0 || g.M < 8 * WS_BM) return false;
  if (g.rm.tlist != nullptr && (g.rm.Bt % WS_BM) != 0) return false;
  if (g.epi == TC_EPI_PLAIN)
    return g.Cb != nullptr && (g.N % 256) == 0 && (g.ldcb % 8) == 0 && (reinterpret_cast<uintptr_t>(g.Cb) & 15) == 0 &&
           (g.bias == nullptr || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
  if (g.epi == TC_EPI_LSTM_P)
    return g.Cb != nullptr && g.N == 1024 && g.Hh == 256 && g.lp_B > 0 && g.lp_vecs >= 16 && g.rm.tlist == nullptr &&
           (reinterpret_cast<uintptr_t>(g.Cb) & 15) == 0 && (g.bias == nullptr || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
  if (g.epi == TC_EPI_DEC_CELL_FWD)
    return g.N == 3 * g.Hh && (g.Hh % 64) == 0 && g.bias != nullptr && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0 &&
           g.gates_b != nullptr && g.hb_out != nullptr;
  return false;
}

int gemm_ws(const TcGemm& g, cudaStream_t st) {
  ARCVAE_REQUIRE(gemm_ws_supported(g), "gemm_ws: unsupported shape");
  WsParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.epi = g.epi == TC_EPI_DEC_CELL_FWD ? WS_EPI_DEC_CELL_FWD : g.epi == TC_EPI_LSTM_P ? WS_EPI_LSTM_P : WS_EPI_PLAIN;
  p.lp_B = g.lp_B;
  p.lp_vecs = g.lp_vecs;
  p.lp_out = reinterpret_cast<uint4*>(g.Cb);
  p.BN = p.epi == WS_EPI_DEC_CELL_FWD ? 192 : 256;
  p.nt = g.N / p.BN;
  p.mt = cdiv(g.M, WS_BM);
  p.kblocks = g.K / WS_BK;
  p.bias = g.bias;
  p.rm = g.rm;
  p.nslab = p.epi == WS_EPI_DEC_CELL_FWD ? 4 : p.epi == WS_EPI_LSTM_P ? 0 : 2;
  const size_t smem = (size_t)p.kblocks * p.BN * WS_BK * 2 + (size_t)WS_STAGES * WS_A_BYTES + (size_t)p.nslab * WS_SLAB +
                      sizeof(WsShared) + 1024;
  ARCVAE_REQUIRE(smem <= 227 * 1024, "gemm_ws shared-memory budget");
  const long rows = g.rm.tlist ? g.a_rows_total : g.M;
  CUtensorMap tmA, tmB, tmC, tmG;
  ARCVAE_TRY(make_tmap_bf16(&tmA, g.A, rows, g.K, g.lda, WS_BK, WS_BM));
  ARCVAE_TRY(make_tmap_bf16(&tmB, g.B, g.N, g.K, g.ldb, WS_BK, p.BN));
  if (p.epi == WS_EPI_PLAIN) {
    ARCVAE_TRY(make_tmap_bf16(&tmC, g.Cb, rows, g.N, g.ldcb, 64, WS_BM));
    tmG = tmC;
  } else if (p.epi == WS_EPI_LSTM_P) {
    tmC = tmA;            // unused: this epilogue stores per thread
    tmG = tmA;
  } else {
    ARCVAE_TRY(make_tmap_bf16(&tmC, g.hb_out, rows, g.Hh, g.Hh, 64, WS_BM));
    ARCVAE_TRY(make_tmap_bf16(&tmG, g.gates_b, rows, 3L * g.Hh, 3L * g.Hh, 64, WS_BM));
  }
  if (first_use_on_device(ONCE_GEMM_WS))
    ARCVAE_CUDA(cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int sms = device_sm_count();
  int per_panel = sms / p.nt;                          // CTAs per weight panel: they walk the row tiles in lock step
  if (per_panel > p.mt) per_panel = p.mt;
  if (per_panel < 1) per_panel = 1;
  TimeScope ts(TIME_GEMM_TC, st);
  count_flops(TIME_GEMM_TC, 2.0 * g.M * g.N * g.K);
  gemm_ws_kernel<<<per_panel * p.nt, WS_THREADS, smem, st>>>(tmA, tmB, tmC, tmG, p);
  ARCVAE_LAUNCHED();
  return 0;
}

}  // namespace arcvae
