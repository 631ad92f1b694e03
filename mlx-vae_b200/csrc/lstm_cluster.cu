// lstm_cluster.cu — the encoder LSTM recurrence (models/encoder.py:98-101; MLX nn.LSTM loop) as ONE persistent
// thread-block-cluster kernel per layer and direction: all T timesteps inside the kernel, W_hh resident in shared
// memory for the whole sequence, gate math fused into the accumulator read-out, h_t / partial d h_t exchanged between the
// CTAs of a cluster as FLAG-IN-DATA vectors through L2 (every 16-byte vector carries its own step flag; the consumer
// thread polls exactly the vectors it needs: no fence, no barrier, no TMA round trip on the critical path).
//
// Decomposition (H = 256): a cluster of 4 CTAs owns a tile of 128 batch rows; CTA r owns hidden units [64r, 64r+64).
//   forward   acc[128 x 256] = h_{t-1}[128 x 256] . Wh[rows {g*H + 64r + u}, :]^T        (4 gates x 64 units)
//             W slice 256 x 256 bf16 = 128 KB resident (K-major); the A operand is a ring of four [128 x 64] slots, slot s =
//             the units of CTA s: the own slot is written by the CTA's epilogue, the other three by its gather
//   backward  partial d h_{t-1}[128 x 256] = dA_t[128 x (4 gates x own 64 units)] . Wh[those rows, :]   ("K-split")
//             the SAME 128 KB image read as an MN-major operand; the four partials are reduce-scattered: the quarter of
//             the CTA's own units stays in TMEM, the other three travel as flagged bf16 vectors
// Roles per CTA (384 threads = 3 warpgroups, setmaxnreg 232 / 40): warps 0-7 epilogue (thread = batch row x 32 hidden
// units; c_t / dL/dc_t live in registers across the whole sequence), warp 8 MMA thread, warp 9 tape thread (TMA stores of the
// bf16 h / dA tapes) + TMEM allocation, warps 10-11 only donate registers.
// The superseded generations (cluster mbarriers + TMA multicast / DSMEM copies) are archived, with the measurements that
// retired them, in profiles/micro/lstm_rec_gen{1,2}_kernel*.cu.txt.
#include <cooperative_groups.h>
#include <cstdlib>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

constexpr int RC_CL = 4;
constexpr int RC_ROWS = 128;
constexpr int RC_STAGE_BYTES = RC_ROWS * 64 * 2;   // 16 KB
constexpr int RC_W_BYTES = 128 * 1024;
constexpr long RC_SPIN_LIMIT = 1L << 24;           // bounded waits: a protocol bug must not hang the GPU
// The tape between the forward and the backward kernel holds, per hidden unit and step, the SIX coefficients of the cell
// reverse instead of the four gates and the cell state (same 12 bytes): with dh = dL/dh_t and s = dL/dc_t carried,
//   d c = s + dh * kc      kc = o (1 - tanh(c_t)^2)          dA_o = dh  * ko     ko = tanh(c_t) o (1 - o)
//   dA_i = d c * ki        ki = g i (1 - i)                   dA_f = d c * kf     kf = c_{t-1} f (1 - f)
//   dA_g = d c * kg        kg = i (1 - g^2)                   s'   = d c * f
// The forward kernel has every factor in fp32 registers anyway (its loop is MUFU-bound, the extra multiplies ride along);
// the backward loop shrinks from ~22 FP operations + 1 MUFU per unit to 6 and keeps 32 fewer registers alive.
// Vector index inside a thread's slot: k * 4 + chunk, k = 0..5 = (ki, kf, kg, ko, kc, f), chunk = 8 units.  The input
// projection of layers >= 1 (gemm_ws TC_EPI_LSTM_P) is written into vectors gate * 4 + chunk of the SAME slot and is
// overwritten by the coefficients once the step has consumed it.
constexpr int RC_KV = 24;

struct RecParams {
  int B, T, H;
  // forward
  const int32_t* xT;      // [T,B] tokens (table mode)
  const bf16* table0b;    // [V,4H] bf16: P_t = table0[x_t]   (layer 0)   -- or --
  const bf16* Pb;         // [T*B,4H] bf16 pre-activations incl. bias (layers >= 1)
  bf16* hb;               // [T*B,H]  h_t, bf16 (exchange + tape)
  bf16* ktape;            // coefficient tape [T][tile][cta][quarter][half][RC_KV vectors][lane] x 16 B (see RC_KV)
  float* h_last;          // [B,H]    h_{T-1}, fp32 (encoder head)
  // backward
  const float* dh_ext;    // [T*B,H] gradient from the layer above (or null)
  const bf16* dh_tf;      // the same gradient, bf16, in the thread-friendly layout [t][tile][cta][quarter][half][chunk 4][lane] x 16 B
  const float* dh_last;   // [B, dh_last_ld] gradient into h_{T-1} from the head (or null)
  int dh_last_ld;
  bf16* dAb;              // [T*B,4H] pre-activation gradients, bf16 (exchange + tape)
  uint4* xch;             // backward, K-split design: exchange buffers of the partial d h (lstm_cluster_xch_bytes)
  int* err_flag;
  long long* dbg;        // optional: per-step clock64 stamps of CTA 0 (layout: [cta 4][it 64][slot 32])
};

namespace rc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// MUFU.TANH based activations (abs error ~2^-11): used on the bf16 path only, where every gate is rounded to bf16
// (2^-9) before it is stored or fed back through the tensor core anyway
__device__ __forceinline__ float tanh_fast(float x) { return tanh_approx_(x); }
__device__ __forceinline__ float sigmoid_fast(float x) { return sigmoid_approx_(x); }
}  // namespace rc

// per-step globaltimer stamps for profiles/rc_stamps3.py: compiled in only with `make STAMPS=1` (they cost registers in
// loops that are tuned to the last one, and the stamping thread runs ~1 us behind its peers)
#ifdef ARCVAE_RC_STAMPS
#define RC_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x < 4 && it < 64) { unsigned long long _gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_gt)); p.dbg[(blockIdx.x * 64 + it) * 32 + (slot)] = (long long)_gt; } } while (0)
#else
#define RC_STAMP(slot) do { } while (0)
#endif

long long* g_rc_dbg = nullptr;   // set by arcvae_debug_set_rc_stamps (tests only)

// ---- helpers shared by the kernels below ----------------------------------------------------------------------------
namespace rc {
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// bounded waits on a CTA-local failure flag (a protocol bug must not hang the GPU)
__device__ __forceinline__ bool wait_flag(uint64_t* bar, uint32_t parity, volatile int* failed) {
  for (long i = 0; i < RC_SPIN_LIMIT; i++) {
    if (tc::mbar_try_wait(bar, parity)) return true;
    if ((i & 1023) == 1023 && *failed) return false;
  }
  *failed = 1;
  return false;
}
}  // namespace rc

constexpr int RC_THREADS2 = 384;   // 3 full warpgroups: setmaxnreg is a warpgroup-wide operation (warps 10, 11 only donate registers)

// =====================================================================================================================
// FLAG-IN-DATA exchange ("LL": every 16-byte vector that crosses CTAs carries 12 bytes of payload and a 4-byte step
// flag).  The previous generation (profiles/micro/lstm_rec_gen2_kernels.cu.txt) signalled with cluster-scope mbarriers: the producer must wait until its stores
// are acknowledged (TMA store completion 1.3 us forward, fence.acq_rel.cluster 0.8 us backward), then signal, then the
// consumer starts its own L2 round trip (multicast load 0.5 us / 12 loads 0.8 us) — three serialized L2 latencies plus
// the skew of waiting for the slowest of 24 remote warps.  Here the consumer THREAD polls the very vectors it needs
// (ld.relaxed.gpu, L2) until their flags show the current step: one store latency + one load latency, no fence, no
// remote arrive, and thread-granular instead of CTA-granular waiting.  ASSUMPTION: a 16-byte aligned vector store / load
// of one thread is a single L2 transaction on sm_100, so payload and flag become visible together (PTX only promises
// per-element atomicity; ptxas must not split the accesses — see ll_send16 — and
// tests/test_gpu_cluster.py::test_cluster_recurrence_is_deterministic fails on any tearing).  Buffers are double-buffered by step parity; the
// dependency chain of the recurrence itself guarantees that a slot has been consumed before it is rewritten two steps
// later (a producer's step t+2 needs every peer's step t+1, which needed this producer's step t in full).  Every
// producer zeroes its outbound slots before the initial cluster barrier, so stale flags of an earlier launch can never
// match.  The tensor-core side is also re-timed: MMAs are issued per K-chunk as the chunks become available (forward:
// the CTA's own h slice first, then each peer's; backward: after each half of the cell reverse) instead of once
// everything has arrived.
namespace rc {
__device__ __forceinline__ void ll_store(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t flag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
constexpr int LL_NV = 6;                   // 32 bf16 = 16 words = 5 vectors of 3 words + 1 word
constexpr long LL_SPIN_LIMIT = 1L << 22;   // ~0.7 us per poll round: a few seconds, then the kernel gives up
// 16 payload words -> 6 flagged vectors at p[0], p[32], ... (one 512-byte warp access each)
__device__ __forceinline__ void ll_send16(uint4* p, const uint32_t (&w)[16], uint32_t flag) {
#pragma unroll
  for (int v = 0; v < 5; v++) ll_store(p + v * 32, w[3 * v], w[3 * v + 1], w[3 * v + 2], flag);
  ll_store(p + 5 * 32, w[15], flag, flag, flag);   // every word of every vector is consumed by the reader: ptxas
                                                   // splits a 128-bit load whose elements are partly unused, and two
                                                   // scalar loads of one vector can tear (seen: flag new, payload stale)
}
// poll the 6 vectors until every flag equals `flag`; false (and the CTA's failure flag) on time-out
__device__ __forceinline__ bool ll_recv16(const uint4* p, uint32_t (&w)[16], uint32_t flag, volatile int* failed) {
  for (long i = 0; i < LL_SPIN_LIMIT; i++) {
    uint4 v[LL_NV];
#pragma unroll
    for (int k = 0; k < LL_NV; k++) v[k] = ll_load(p + k * 32);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < LL_NV; k++) ok = ok && (v[k].w == flag);
    ok = ok && (v[5].y == flag) && (v[5].z == flag);
    if (ok) {
#pragma unroll
      for (int k = 0; k < 5; k++) { w[3 * k] = v[k].x; w[3 * k + 1] = v[k].y; w[3 * k + 2] = v[k].z; }
      w[15] = v[5].x;
      return true;
    }
    if ((i & 255) == 255 && *failed) return false;
  }
  *failed = 1;
  return false;
}
__device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
  const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
  a = t.x; b = t.y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
}  // namespace rc

struct __align__(8) Bwd3Shared {
  uint64_t w_ready, a_ready[2], a_free, acc_full;
  uint32_t tmem_base;
  int failed;
};

// exchange buffer of the backward: [parity][tile][dst][src][warp 8][vector 6][lane 32] x 16 B
__global__ void __launch_bounds__(RC_THREADS2, 1)
lstm_bwd3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD, const RecParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem;                         // 4 gate panels x [64 k-rows x 256 units] bf16, MN-major, 32 KB each
  uint8_t* At = smem + RC_W_BYTES;             // 4 gate panels x [128 rows x 64 units] bf16, K-major, 16 KB each
  Bwd3Shared* sh = reinterpret_cast<Bwd3Shared*>(At + 4 * RC_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)rc::cluster_ctarank();
  const int tile = blockIdx.x / RC_CL;
  const int ntiles = gridDim.x / RC_CL;
  const int row0 = tile * RC_ROWS;
  const int B = p.B, T = p.T, H = p.H;
  volatile int* failed = &sh->failed;
  auto xch = [&](int par, int dst, int src, int w) -> uint4* {
    return p.xch + (((((long)par * ntiles + tile) * RC_CL + dst) * RC_CL + src) * 8 + w) * (rc::LL_NV * 32) + lane;
  };

  if (threadIdx.x == 0) {
    tc::mbar_init(&sh->w_ready, 1);
    tc::mbar_init(&sh->a_ready[0], 8);
    tc::mbar_init(&sh->a_ready[1], 8);
    tc::mbar_init(&sh->a_free, 2);
    tc::mbar_init(&sh->acc_full, 1);
    sh->failed = 0;
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmW);
    tc::prefetch_tmap(&tmD);
  }
  if (warp < 8) {
    // stale flags of an earlier launch must not match: clear this thread's outbound vectors (both parities)
    for (int par = 0; par < 2; par++)
      for (int c = 0; c < RC_CL; c++)
        if (c != rank)
          for (int v = 0; v < rc::LL_NV; v++) rc::ll_store(xch(par, c, rank, warp) + v * 32, 0u, 0u, 0u, 0u);
  }
  if (warp == 9) tc::tmem_alloc(&sh->tmem_base, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  rc::cluster_sync_all();                      // barriers exist and the cleared flags are ordered before any peer's stores
  const uint32_t tmem_base = sh->tmem_base;

  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    // =========================================================== control: resident weights, then 2 x 8 MMAs per step
    if (lane == 0) {
      tc::mbar_expect_tx(&sh->w_ready, RC_W_BYTES);
      for (int g = 0; g < 4; g++)
        for (int j = 0; j < 4; j++)
          tc::tma_load_2d(Wsm + g * 32768 + j * 8192, &tmW, &sh->w_ready, 64 * j, g * H + 64 * rank);
      bool ok = rc::wait_flag(&sh->w_ready, 0, failed);
      const uint32_t idesc = tc::make_idesc_bf16(RC_ROWS, 256, false, true);
      for (int it = 0; it + 1 < T && ok; it++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
          // half 0: the units every epilogue thread finishes first (16-unit K blocks 0 and 2), half 1: blocks 1 and 3
          ok = ok && rc::wait_flag(&sh->a_ready[half], it & 1, failed);
          if (!ok) break;
          if (half == 0) RC_STAMP(0); else RC_STAMP(1);
          tc::tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const uint32_t a_addr = tc::smem_u32(At + g * RC_STAGE_BYTES);
            const uint32_t b_addr = tc::smem_u32(Wsm + g * 32768);
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {
              const int k = 2 * kk + half;
              tc::mma_bf16(tmem_base, tc::make_smem_desc(a_addr + 32 * k, 16, 1024),
                           tc::make_smem_desc(b_addr + 2048 * k, 8192, 1024), idesc, (half > 0 || g > 0 || kk > 0) ? 1u : 0u);
            }
          }
          if (half == 0) RC_STAMP(3);
        }
        if (!ok) break;
        tc::mma_commit(&sh->a_free);             // operand tile may be rewritten once these MMAs have read it
        tc::mma_commit(&sh->acc_full);
        RC_STAMP(2);
      }
    }
  } else if (warp == 9) {
    // =========================================================== tape: dA_t tile -> HBM by TMA (rows >= B are clipped)
    if (lane == 0) {
      bool ok = true;
      for (int it = 0; it < T && ok; it++) {
        ok = rc::wait_flag(&sh->a_ready[1], it & 1, failed);   // (measured: waiting for the MMAs to finish first is slower)
        if (!ok) break;
        const int t = T - 1 - it;
        RC_STAMP(8);
#pragma unroll
        for (int g = 0; g < 4; g++) rc::tma_store_3d(&tmD, At + g * RC_STAGE_BYTES, g * H + 64 * rank, row0, t);
        rc::bulk_commit();
        rc::bulk_wait_read0();
        RC_STAMP(9);
        tc::mbar_arrive(&sh->a_free);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last tile is in HBM before the kernel ends
    }
  } else if (warp < 8) {
    // =========================================================== epilogue: cell backward, state in registers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int hs = warp >> 2;                // which 32 of the CTA's 64 hidden units
    const int rl = q * 32 + lane;            // row inside the tile
    const int row = row0 + rl;
    const bool valid = row < B;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ub = 64 * rank + hs * 32;      // first global hidden unit of this thread
    const long warp_slot = ((((long)tile * RC_CL + rank) * 4 + q) * 2 + hs);
    const long slots_per_t = (long)ntiles * RC_CL * 4 * 2;
    auto ktape = [&](int tt, int kk, int cu) -> const uint4* {
      return reinterpret_cast<const uint4*>(p.ktape) + (((long)tt * slots_per_t + warp_slot) * RC_KV + kk * 4 + cu) * 32 + lane;
    };

    float state[32];                         // dL/dc_t carried to t-1
#pragma unroll
    for (int i = 0; i < 32; i++) state[i] = 0.f;
    uint4 pf[4][6];                          // [chunk][coefficient] of the step being prefetched
    float dhx[32];                           // d h_t of this thread's units: starts as the gradient from the layer above
                                             // (dh_ext, prefetched), then + own partial (TMEM) + three remote partials
#pragma unroll
    for (int i = 0; i < 32; i++) dhx[i] = 0.f;
    // tape of step tt for one 8-unit chunk
    auto prefetch = [&](int tt, int cu) {
      if (!valid) return;
      if (p.dh_tf != nullptr) {
        float f[8];
        rc::unpack8(__ldcs(reinterpret_cast<const uint4*>(p.dh_tf) + (((long)tt * slots_per_t + warp_slot) * 4 + cu) * 32 + lane), f);
#pragma unroll
        for (int j = 0; j < 8; j++) dhx[8 * cu + j] = f[j];
      } else if (p.dh_ext != nullptr) {
        const float4* e = reinterpret_cast<const float4*>(p.dh_ext + ((long)tt * B + row) * H + ub) + 2 * cu;
        const float4 v0 = __ldg(e), v1 = __ldg(e + 1);
        dhx[8 * cu] = v0.x; dhx[8 * cu + 1] = v0.y; dhx[8 * cu + 2] = v0.z; dhx[8 * cu + 3] = v0.w;
        dhx[8 * cu + 4] = v1.x; dhx[8 * cu + 5] = v1.y; dhx[8 * cu + 6] = v1.z; dhx[8 * cu + 7] = v1.w;
      }
#pragma unroll
      for (int kk = 0; kk < 6; kk++) pf[cu][kk] = __ldcs(ktape(tt, kk, cu));
    };
    if (!valid) {
#pragma unroll
      for (int cu = 0; cu < 4; cu++)
#pragma unroll
        for (int kk = 0; kk < 6; kk++) pf[cu][kk] = make_uint4(0u, 0u, 0u, 0u);   // rows beyond the batch: zero gradients
    }
#pragma unroll
    for (int cu = 0; cu < 4; cu++) prefetch(T - 1, cu);
    bool ok = true;
    for (int it = 0; it < T && ok; it++) {
      const int t = T - 1 - it;
      if (threadIdx.x == 0) RC_STAMP(4);
      // d h_t of this thread's 32 units: the CTA's own partial (still in TMEM; read NOW, the first MMAs of this step
      // overwrite it) plus the three remote partials, polled straight out of the exchange buffer with two sources in
      // flight (one L2 round trip in total when the data is already there)
      if (it > 0) {
        const uint32_t flag = (uint32_t)it;
        const uint4* src[RC_CL - 1];
#pragma unroll
        // poll in the order the peers SEND: CTA s sends to s+1 first, so this CTA's first vectors come from rank-1 = rank+3
        for (int j3 = 1; j3 < RC_CL; j3++) src[j3 - 1] = xch((it - 1) & 1, rank, (rank + RC_CL - j3) & (RC_CL - 1), warp);
        uint4 va[rc::LL_NV], vb[rc::LL_NV];
        auto issue = [&](uint4 (&v)[rc::LL_NV], const uint4* q) {
#pragma unroll
          for (int k = 0; k < rc::LL_NV; k++) v[k] = rc::ll_load(q + k * 32);
        };
        auto ready = [&](const uint4 (&v)[rc::LL_NV]) {
          bool r = true;
#pragma unroll
          for (int k = 0; k < rc::LL_NV; k++) r = r && (v[k].w == flag);
          return r && (v[5].y == flag) && (v[5].z == flag);
        };
        auto add = [&](const uint4 (&v)[rc::LL_NV]) {
#pragma unroll
          for (int k = 0; k < 5; k++) {
            float a, b;
            rc::unpack2(v[k].x, a, b); dhx[6 * k] += a; dhx[6 * k + 1] += b;
            rc::unpack2(v[k].y, a, b); dhx[6 * k + 2] += a; dhx[6 * k + 3] += b;
            rc::unpack2(v[k].z, a, b); dhx[6 * k + 4] += a; dhx[6 * k + 5] += b;
          }
          float a, b;
          rc::unpack2(v[5].x, a, b); dhx[30] += a; dhx[31] += b;
        };
        auto settle = [&](uint4 (&v)[rc::LL_NV], const uint4* q) {
          for (long i = 0; !ready(v); i++) {
            if (i >= rc::LL_SPIN_LIMIT) { *failed = 1; return false; }
            if ((i & 255) == 255 && *failed) return false;
            issue(v, q);
          }
          return true;
        };
        issue(va, src[0]);
        issue(vb, src[1]);
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          uint32_t v[16];
          tc::tmem_ld16(taddr + (uint32_t)(64 * rank + hs * 32 + hf * 16), v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j++) dhx[hf * 16 + j] += __uint_as_float(v[j]);
        }
        tc::tc_fence_before();
        ok = settle(va, src[0]);
        if (ok) { add(va); issue(va, src[2]); ok = settle(vb, src[1]); }
        if (ok) { add(vb); ok = settle(va, src[2]); }
        if (ok) add(va);
        if (ok) ok = rc::wait_flag(&sh->a_free, (it - 1) & 1, failed);         // operand tile reusable
        if (!ok) break;
      }
      if (threadIdx.x == 0) RC_STAMP(6);
      if (lane == 0) RC_STAMP(16 + warp);
#pragma unroll
      for (int cu = 0; cu < 4; cu++) {
        const int ug = ub + cu * 8;                  // global hidden-unit index
        float dh[8];
#pragma unroll
        for (int j = 0; j < 8; j++) dh[j] = dhx[cu * 8 + j];
        if (valid) {
          if (t == T - 1 && p.dh_last != nullptr) {
            const float* e = p.dh_last + (long)row * p.dh_last_ld + ug;
#pragma unroll
            for (int j = 0; j < 8; j++) dh[j] += e[j];
          }
        }
        float ki[8], kf[8], kg[8], ko[8], kc[8], gf[8];
        rc::unpack8(pf[cu][0], ki);
        rc::unpack8(pf[cu][1], kf);
        rc::unpack8(pf[cu][2], kg);
        rc::unpack8(pf[cu][3], ko);
        rc::unpack8(pf[cu][4], kc);
        rc::unpack8(pf[cu][5], gf);
        float ai[8], af[8], ag[8], ao[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float dct = fmaf(dh[j], kc[j], state[cu * 8 + j]);
          ao[j] = dh[j] * ko[j];
          ai[j] = dct * ki[j];
          af[j] = dct * kf[j];
          ag[j] = dct * kg[j];
          state[cu * 8 + j] = dct * gf[j];
        }
        // dA_t of these 8 units -> operand tile (gate panel g, row rl, 16-byte chunk hs*4+cu, SWIZZLE_128B)
        uint8_t* arow = At + rl * 128 + (((hs * 4 + cu) ^ (rl & 7)) << 4);
        *reinterpret_cast<uint4*>(arow) = rc::pack8(ai);
        *reinterpret_cast<uint4*>(arow + RC_STAGE_BYTES) = rc::pack8(af);
        *reinterpret_cast<uint4*>(arow + 2 * RC_STAGE_BYTES) = rc::pack8(ag);
        *reinterpret_cast<uint4*>(arow + 3 * RC_STAGE_BYTES) = rc::pack8(ao);
        if (cu == 1 || cu == 3) {
          tc::fence_proxy_async();           // generic-proxy smem writes -> visible to tcgen05.mma / TMA store
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&sh->a_ready[cu >> 1]);
          if (threadIdx.x == 0) { if (cu == 1) RC_STAMP(5); else RC_STAMP(7); }
          if (lane == 0 && cu == 3) RC_STAMP(24 + warp);
        }
      }
      if (it + 1 < T) {
        // tape of the next step: latency hides behind the MMAs (measured: issued inside the cell reverse they delay it)
#pragma unroll
        for (int i = 0; i < 32; i++) dhx[i] = 0.f;
#pragma unroll
        for (int cu = 0; cu < 4; cu++) prefetch(t - 1, cu);
        if (threadIdx.x == 0) RC_STAMP(10);
        ok = rc::wait_flag(&sh->acc_full, it & 1, failed);
        if (!ok) break;
        if (threadIdx.x == 0) RC_STAMP(11);
        tc::tc_fence_after();
        // partial d h_{t-1}[128 x 256]: the quarter of this CTA's units stays in TMEM (read by the next epilogue), the
        // other three go to their owners as flagged vectors
#pragma unroll
        for (int j3 = 1; j3 < RC_CL; j3++) {
          const int c = (rank + j3) & (RC_CL - 1);
          uint32_t w[16];
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            uint32_t v[16];
            tc::tmem_ld16(taddr + (uint32_t)(64 * c + hs * 32 + hf * 16), v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; j++) w[hf * 8 + j] = rc::pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          }
          rc::ll_send16(xch(it & 1, c, rank, warp), w, (uint32_t)(it + 1));
        }
        if (threadIdx.x == 0) RC_STAMP(12);
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && sh->failed && p.err_flag != nullptr) atomicExch(p.err_flag, 1);
  rc::cluster_sync_all();                    // nobody exits while a peer may still arrive on its barriers
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

// Forward, third design.  h_t leaves the epilogue thread three ways at once: (1) into the CTA's own ring slot (operand of
// its own MMAs and source of the TMA store that writes the bf16 h tape), (2) as six flagged vectors into the exchange
// buffer [parity][tile][src][warp 8][vector 6][lane 32] x 16 B, from which (3) the SAME (warp, lane) thread of each peer
// polls them and writes them into ring slot `src` of its own CTA as a SWIZZLE_128B operand row.  The MMA thread issues
// the four K=16 MMAs of a slot as soon as that slot is complete — own slot first (available ~2 us before the peers').
struct __align__(8) Fwd3Shared {
  uint64_t full[RC_CL];      // slot s holds h_t of source CTA s: 8 arrivals (the epilogue warps that wrote it)
  uint64_t slot_free;        // own slot read by the TMA store of the tape
  uint64_t w_ready, acc_full;
  uint32_t tmem_base;
  int failed;
};

__global__ void __launch_bounds__(RC_THREADS2, 1)
lstm_fwd3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, const RecParams p) {
  constexpr int BN = 256;
  constexpr int WPANEL = BN * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Wsm = smem;
  uint8_t* ring = smem + RC_W_BYTES;
  Fwd3Shared* sh = reinterpret_cast<Fwd3Shared*>(ring + RC_CL * RC_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)rc::cluster_ctarank();
  const int tile = blockIdx.x / RC_CL;
  const int ntiles = gridDim.x / RC_CL;
  const int row0 = tile * RC_ROWS;
  const int B = p.B, T = p.T, H = p.H;
  volatile int* failed = &sh->failed;
  auto xh = [&](int par, int src, int w) -> uint4* {
    return p.xch + ((((long)par * ntiles + tile) * RC_CL + src) * 8 + w) * (rc::LL_NV * 32) + lane;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < RC_CL; s++) tc::mbar_init(&sh->full[s], 8);
    tc::mbar_init(&sh->slot_free, 1);
    tc::mbar_init(&sh->w_ready, 1);
    tc::mbar_init(&sh->acc_full, 1);
    sh->failed = 0;
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmW);
    tc::prefetch_tmap(&tmH);
  }
  if (warp < 8) {
    for (int par = 0; par < 2; par++)
      for (int v = 0; v < rc::LL_NV; v++) rc::ll_store(xh(par, rank, warp) + v * 32, 0u, 0u, 0u, 0u);
  }
  if (warp == 9) tc::tmem_alloc(&sh->tmem_base, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  rc::cluster_sync_all();                      // cleared flags are ordered before any peer's poll
  const uint32_t tmem_base = sh->tmem_base;

  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    // =========================================================== control: weights once, then 4 x 4 MMAs per step
    if (lane == 0) {
      tc::mbar_expect_tx(&sh->w_ready, RC_W_BYTES);
      for (int kp = 0; kp < 4; kp++)
        for (int g = 0; g < 4; g++)
          tc::tma_load_2d(Wsm + kp * WPANEL + g * 8192, &tmW, &sh->w_ready, 64 * kp, g * H + 64 * rank);
      bool ok = rc::wait_flag(&sh->w_ready, 0, failed);
      const uint32_t idesc = tc::make_idesc_bf16(RC_ROWS, BN, false, false);
      // descriptors of slot 0 / panel 0; the start-address field counts 16-byte units, so slots and K steps are plain adds
      // (keeps this thread inside its 48 registers: a spill here is a local-memory load queued behind the tape traffic)
      const uint64_t adesc0 = tc::make_smem_desc(tc::smem_u32(ring), 16, 1024);
      const uint64_t bdesc0 = tc::make_smem_desc(tc::smem_u32(Wsm), 16, 1024);
      for (int it = 1; it < T && ok; it++) {
#pragma unroll 1
        for (int j4 = 0; j4 < RC_CL; j4++) {
          const int kq = (rank + j4) & (RC_CL - 1);          // fixed order (deterministic sums): own slice, then the peers
          ok = rc::wait_flag(&sh->full[kq], (it - 1) & 1, failed);
          if (!ok) break;
          if (j4 == 0) RC_STAMP(0);
          if (j4 == 3) RC_STAMP(1);
          tc::tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(kq * (RC_STAGE_BYTES >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(kq * (WPANEL >> 4));
#pragma unroll
          for (int j = 0; j < 4; j++) tc::mma_bf16(tmem_base, ad + 2 * j, bd + 2 * j, idesc, (j4 > 0 || j > 0) ? 1u : 0u);
        }
        if (!ok) break;
        tc::mma_commit(&sh->acc_full);
        RC_STAMP(2);
      }
    }
  } else if (warp == 9) {
    // =========================================================== tape: own slot -> HBM (bf16 h, rows >= B clipped)
    if (lane == 0) {
      bool ok = true;
      for (int it = 0; it < T && ok; it++) {
        ok = rc::wait_flag(&sh->full[rank], it & 1, failed);
        if (!ok) break;
        RC_STAMP(8);
        rc::tma_store_3d(&tmH, ring + rank * RC_STAGE_BYTES, 64 * rank, row0, it);
        rc::bulk_commit();
        rc::bulk_wait_read0();
        RC_STAMP(9);
        tc::mbar_arrive(&sh->slot_free);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp < 8) {
    // =========================================================== epilogue: gate math, cell state in registers
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;
    const int hs = warp >> 2;
    const int rl = q * 32 + lane;
    const int row = row0 + rl;
    const bool valid = row < B;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ub = 64 * rank + hs * 32;
    const long warp_slot = ((((long)tile * RC_CL + rank) * 4 + q) * 2 + hs);
    const long slots_per_t = (long)ntiles * RC_CL * 4 * 2;
    auto ktape = [&](int tt, int kk, int cu) -> uint4* {
      return reinterpret_cast<uint4*>(p.ktape) + (((long)tt * slots_per_t + warp_slot) * RC_KV + kk * 4 + cu) * 32 + lane;
    };
    const uint32_t row_off = (uint32_t)rl * 128;                     // this thread's row inside a ring slot
    float state[32];                         // c_{t-1}
#pragma unroll
    for (int i = 0; i < 32; i++) state[i] = 0.f;
    uint4 pf[4][4];                          // [chunk][gate] 8 x bf16 of P_t, prefetched one step ahead
    // operands of step tt for one 8-unit chunk
    auto prefetch = [&](int tt, int cu) {
      if (!valid) return;
      if (p.table0b == nullptr && p.Pb == nullptr) {
        // input projection written by the weight-stationary GEMM (TC_EPI_LSTM_P) into THIS layer's tape slots, in the
        // tape's own thread-friendly layout: 512-byte warp accesses; the coefficients overwrite it once consumed
#pragma unroll
        for (int g = 0; g < 4; g++) pf[cu][g] = __ldcs(ktape(tt, g, cu));
        return;
      }
      const long rr = (long)tt * B + row;
      const bf16* prow = (p.table0b != nullptr ? p.table0b + (long)__ldg(p.xT + rr) * 4 * H : p.Pb + rr * 4 * H) + ub;
#pragma unroll
      for (int g = 0; g < 4; g++) pf[cu][g] = __ldg(reinterpret_cast<const uint4*>(prow + g * H + cu * 8));
    };
    if (!valid) {
#pragma unroll
      for (int cu = 0; cu < 4; cu++)
#pragma unroll
        for (int g = 0; g < 4; g++) pf[cu][g] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int cu = 0; cu < 4; cu++) prefetch(0, cu);
    bool ok = true;
    for (int it = 0; it < T && ok; it++) {
      const int t = it;
      if (it > 0) {
        ok = rc::wait_flag(&sh->acc_full, (it - 1) & 1, failed);
        if (!ok) break;
        tc::tc_fence_after();
      }
      if (threadIdx.x == 0) RC_STAMP(4);
      uint4 hq[4];                           // h_t
#pragma unroll
      for (int cu = 0; cu < 4; cu++) {
        const int u0 = hs * 32 + cu * 8;             // unit offset inside the CTA's 64
        float a[4][8];
        if (it > 0) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint32_t rr[8];
            rc::tmem_ld8(taddr + (uint32_t)(g * 64 + u0), rr);
#pragma unroll
            for (int j = 0; j < 8; j++) a[g][j] = __uint_as_float(rr[j]);
          }
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int g = 0; g < 4; g++)
#pragma unroll
            for (int j = 0; j < 8; j++) a[g][j] = 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
          float f[8];
          rc::unpack8(pf[cu][g], f);
#pragma unroll
          for (int j = 0; j < 8; j++) a[g][j] += f[j];
        }
        float hv[8], ki[8], kf[8], kg[8], ko[8], kc[8], gfv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float gi = rc::sigmoid_fast(a[0][j]);
          const float gf = rc::sigmoid_fast(a[1][j]);
          const float gg = rc::tanh_fast(a[2][j]);
          const float go = rc::sigmoid_fast(a[3][j]);
          const float cp = state[cu * 8 + j];                              // t = 0: c_{-1} = 0 -> c = i*g, kf = 0
          const float cn = fmaf(gf, cp, gi * gg);
          state[cu * 8 + j] = cn;
          const float tcv = rc::tanh_fast(cn);
          hv[j] = go * tcv;
          ki[j] = gg * gi * (1.f - gi);
          kf[j] = cp * gf * (1.f - gf);
          kg[j] = gi * (1.f - gg * gg);
          ko[j] = tcv * go * (1.f - go);
          kc[j] = go * (1.f - tcv * tcv);
          gfv[j] = gf;
        }
        hq[cu] = rc::pack8(hv);
        // the tape of this chunk leaves NOW: the LSU is idle during the MUFU-bound gate math, and 24 stores per thread issued
        // in one burst after the exchange were measured to block the thread for > 1 us (1.5 us of the step)
        if (valid) {
          *ktape(t, 0, cu) = rc::pack8(ki);
          *ktape(t, 1, cu) = rc::pack8(kf);
          *ktape(t, 2, cu) = rc::pack8(kg);
          *ktape(t, 3, cu) = rc::pack8(ko);
          *ktape(t, 4, cu) = rc::pack8(kc);
          *ktape(t, 5, cu) = rc::pack8(gfv);
        }
        if (valid && t == T - 1 && p.h_last != nullptr) {
          float* hl = p.h_last + (long)row * H + ub + cu * 8;
          *reinterpret_cast<float4*>(hl) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<float4*>(hl + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        }
      }
      if (threadIdx.x == 0) RC_STAMP(5);
      // ---- exchange: flagged vectors to the peers first (longest latency), then the CTA's own slot
      if (it + 1 < T) {
        uint32_t w[16];
#pragma unroll
        for (int cu = 0; cu < 4; cu++) { w[4 * cu] = hq[cu].x; w[4 * cu + 1] = hq[cu].y; w[4 * cu + 2] = hq[cu].z; w[4 * cu + 3] = hq[cu].w; }
        rc::ll_send16(xh(it & 1, rank, warp), w, (uint32_t)(it + 1));
      }
      if (it > 0) {
        ok = rc::wait_flag(&sh->slot_free, (it - 1) & 1, failed);      // the tape store has read h_{t-1} out of the own slot
        if (!ok) break;
      }
      {
        uint8_t* own_row = ring + rank * RC_STAGE_BYTES + row_off;
#pragma unroll
        for (int cu = 0; cu < 4; cu++) *reinterpret_cast<uint4*>(own_row + (((hs * 4 + cu) ^ (rl & 7)) << 4)) = hq[cu];
      }
      tc::tc_fence_before();                 // this step's TMEM reads are done before the MMAs that overwrite the accumulator
      tc::fence_proxy_async();               // generic-proxy smem writes -> visible to tcgen05.mma / TMA
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sh->full[rank]);
      if (threadIdx.x == 0) RC_STAMP(7);
      if (lane == 0) RC_STAMP(24 + warp);
      // ---- gather: the peers' h_t rows of this thread -> ring slots of this CTA (critical path: before the tape)
      if (it + 1 < T) {
        if (threadIdx.x == 0) RC_STAMP(10);
        const uint32_t flag = (uint32_t)(it + 1);
        uint4 va[rc::LL_NV], vb[rc::LL_NV], vc[rc::LL_NV];
        auto issue = [&](uint4 (&v)[rc::LL_NV], int s) {
          const uint4* q = xh(it & 1, s, warp);
#pragma unroll
          for (int k = 0; k < rc::LL_NV; k++) v[k] = rc::ll_load(q + k * 32);
        };
        long rounds = 0;
        auto settle = [&](uint4 (&v)[rc::LL_NV], int s) {
          for (long i = 0;; i++) {
            bool r = true;
#pragma unroll
            for (int k = 0; k < rc::LL_NV; k++) r = r && (v[k].w == flag);
            rounds++;
            if (r && (v[5].y == flag) && (v[5].z == flag)) return true;
            if (i >= rc::LL_SPIN_LIMIT) { *failed = 1; return false; }
            if ((i & 255) == 255 && *failed) return false;
            issue(v, s);
          }
        };
        // 16 payload words of a peer's row -> ring slot s of this CTA, then tell the MMA thread
        auto deliver = [&](const uint4 (&v)[rc::LL_NV], int s) {
          uint8_t* prow = ring + s * RC_STAGE_BYTES + row_off;
          *reinterpret_cast<uint4*>(prow + (((hs * 4 + 0) ^ (rl & 7)) << 4)) = make_uint4(v[0].x, v[0].y, v[0].z, v[1].x);
          *reinterpret_cast<uint4*>(prow + (((hs * 4 + 1) ^ (rl & 7)) << 4)) = make_uint4(v[1].y, v[1].z, v[2].x, v[2].y);
          *reinterpret_cast<uint4*>(prow + (((hs * 4 + 2) ^ (rl & 7)) << 4)) = make_uint4(v[2].z, v[3].x, v[3].y, v[3].z);
          *reinterpret_cast<uint4*>(prow + (((hs * 4 + 3) ^ (rl & 7)) << 4)) = make_uint4(v[4].x, v[4].y, v[4].z, v[5].x);
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&sh->full[s]);
        };
        const int s1 = (rank + 1) & (RC_CL - 1), s2 = (rank + 2) & (RC_CL - 1), s3 = (rank + 3) & (RC_CL - 1);
        issue(va, s1);                       // all three peers in flight: one L2 round trip when the data is there
        issue(vb, s2);
        issue(vc, s3);
        ok = settle(va, s1);
        if (ok) { deliver(va, s1); ok = settle(vb, s2); }
        if (ok) { deliver(vb, s2); ok = settle(vc, s3); }
        if (ok) deliver(vc, s3);
        if (!ok) break;
        if (threadIdx.x == 0) RC_STAMP(11);
#ifdef ARCVAE_RC_STAMPS
        if (p.dbg != nullptr && threadIdx.x == 0 && blockIdx.x < 4 && it < 64) p.dbg[(blockIdx.x * 64 + it) * 32 + 13] = rounds;
#endif
        if (lane == 0) RC_STAMP(16 + warp);
        // next step's operands, under the MMAs (measured: issued BEFORE the polls they cost 1.4 us per step — the polls
        // queue behind 16 loads from HBM — and inside the gate math they throttle it)
#pragma unroll
        for (int cu = 0; cu < 4; cu++) prefetch(t + 1, cu);
      }
      if (threadIdx.x == 0) RC_STAMP(12);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && sh->failed && p.err_flag != nullptr) atomicExch(p.err_flag, 1);
  rc::cluster_sync_all();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, BN);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
int make_tmap_bf16(CUtensorMap* m, const bf16* ptr, long rows, long cols, long ld, int box_cols, int box_rows);


bool lstm_cluster_supported(int H) { return H == 256; }

int make_tmap_bf16_3d(CUtensorMap* m, const bf16* ptr, long T, long rows_per_t, long cols, long ld, int box_cols, int box_rows);

static int launch_cluster384(const void* fn, size_t smem, int B, const CUtensorMap& a, const CUtensorMap& b, const RecParams& p,
                             cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cdiv(B, RC_ROWS) * RC_CL);
  cfg.blockDim = dim3(RC_THREADS2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = RC_CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  void* args[3] = {(void*)&a, (void*)&b, (void*)&p};
  TimeScope ts(TIME_RECURRENCE, st);
  count_flops(TIME_RECURRENCE, 2.0 * 4 * p.H * p.H * (double)cdiv(B, RC_ROWS) * RC_ROWS * p.T);
  ARCVAE_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

size_t lstm_cluster_xh_bytes(int B) { return (size_t)2 * cdiv(B, RC_ROWS) * RC_CL * 8 * rc::LL_NV * 32 * sizeof(uint4); }
size_t lstm_cluster_xch_bytes(int B) { return (size_t)2 * cdiv(B, RC_ROWS) * RC_CL * RC_CL * 8 * rc::LL_NV * 32 * sizeof(uint4); }
// coefficient tape of one layer (bf16 elements): T x tile-padded batch x 6 coefficients per hidden unit
size_t lstm_cluster_dh_tf_elems(int B, int T, int H) { return (size_t)T * cdiv(B, RC_ROWS) * RC_ROWS * H; }
size_t lstm_cluster_ktape_elems(int B, int T, int H) { return (size_t)T * cdiv(B, RC_ROWS) * RC_ROWS * 6 * H; }

// Pb == nullptr && table0b == nullptr: the input projection sits in the tape slots (gemm_ws TC_EPI_LSTM_P)
int lstm_cluster_forward(int B, int T, int H, const bf16* Whb, const int32_t* xT, const bf16* table0b, const bf16* Pb,
                         bf16* hb, bf16* ktape, float* h_last, void* xh, int* err_flag, cudaStream_t st) {
  ARCVAE_REQUIRE(lstm_cluster_supported(H), "cluster recurrence kernel is built for hidden_dim 256");
  ARCVAE_REQUIRE(T < (1 << 30), "sequence length");
  ARCVAE_REQUIRE(xh != nullptr && ktape != nullptr, "forward exchange buffer / tape");
  CUtensorMap tmW, tmH;
  ARCVAE_TRY(make_tmap_bf16(&tmW, Whb, 4L * H, H, H, 64, 64));
  ARCVAE_TRY(make_tmap_bf16_3d(&tmH, hb, T, B, H, H, 64, RC_ROWS));
  RecParams p{};
  p.B = B; p.T = T; p.H = H;
  p.xT = xT; p.table0b = table0b; p.Pb = Pb; p.hb = hb; p.ktape = ktape; p.h_last = h_last;
  p.xch = reinterpret_cast<uint4*>(xh);
  p.err_flag = err_flag;
  p.dbg = g_rc_dbg;
  const size_t smem = RC_W_BYTES + RC_CL * RC_STAGE_BYTES + sizeof(Fwd3Shared) + 1024;
  if (first_use_on_device(ONCE_FWD3))
    ARCVAE_CUDA(cudaFuncSetAttribute(lstm_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_cluster384((const void*)lstm_fwd3_kernel, smem, B, tmW, tmH, p, st);
}

// K-split backward: Whb is the SAME [4H,H] bf16 matrix the forward kernel uses
int lstm_cluster_backward(int B, int T, int H, const bf16* Whb, const bf16* ktape, const float* dh_ext, const bf16* dh_tf,
                          const float* dh_last, int dh_last_ld, bf16* dAb, void* xch, int* err_flag, cudaStream_t st) {
  ARCVAE_REQUIRE(lstm_cluster_supported(H), "cluster recurrence kernel is built for hidden_dim 256");
  CUtensorMap tmW, tmD;
  ARCVAE_TRY(make_tmap_bf16(&tmW, Whb, 4L * H, H, H, 64, 64));
  ARCVAE_TRY(make_tmap_bf16_3d(&tmD, dAb, T, B, 4L * H, 4L * H, 64, RC_ROWS));
  RecParams p{};
  p.B = B; p.T = T; p.H = H;
  p.ktape = const_cast<bf16*>(ktape);
  p.dh_ext = dh_ext; p.dh_tf = dh_tf; p.dh_last = dh_last; p.dh_last_ld = dh_last_ld; p.dAb = dAb; p.err_flag = err_flag;
  p.xch = reinterpret_cast<uint4*>(xch);
  p.dbg = g_rc_dbg;
  const size_t smem = RC_W_BYTES + 4 * RC_STAGE_BYTES + sizeof(Bwd3Shared) + 1024;
  if (first_use_on_device(ONCE_BWD3))
    ARCVAE_CUDA(cudaFuncSetAttribute(lstm_bwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_cluster384((const void*)lstm_bwd3_kernel, smem, B, tmW, tmD, p, st);
}

}  // namespace arcvae

// debug aid: when non-null, subsequent cluster-kernel launches record globaltimer stamps of CTAs 0-3 into buf[4*64*32]
extern "C" int arcvae_debug_set_rc_stamps(long long* buf) {
  arcvae::g_rc_dbg = buf;
  return 0;
}
