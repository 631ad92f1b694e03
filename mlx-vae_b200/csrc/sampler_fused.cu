// sampler_fused.cu — MLXAutoregressiveDecoderSampling.generate_with_temperature (models/decoder_sampling.py:48-128)
// as ONE persistent kernel per batch: every CTA owns a tile of 128 molecules for all max_length steps; token
// embedding -> stacked zero-state LSTM cells -> fc_out -> /temperature -> argmax or multinomial never leave the SM.
//
// The reference decoder is stateless across positions, so a step is a chain of small dense products on the
// tile's rows.  All of them run on tcgen05 with the activations RESIDENT in shared memory as UMMA operand tiles:
//   layer 0   a0 = onehot(tok)[128 x Vpad] . Tt[Vpad x 3H]           (Tt = (Emb Wx0^T + b0)^T, the layer-0 table; the
//             + cond (x) wc                                            gather becomes a K = V product; cond term in the epilogue)
//   layer l   a_l = h_{l-1}[128 x H] . Wx_l^T                         (compact (i,g,o) rows, tile-permuted so one 192-wide
//                                                                      accumulator block holds i|g|o of the same 64 units)
//   cell      h_l = s(o) * tanh(s(i) * tanh(g))                       (epilogue: TMEM -> registers -> bf16 operand tile in smem)
//   fc_out    logits = h_top . Wout^T + b                             (accumulator -> thread-per-row selection)
// Weights are streamed from L2 through a 4-stage TMA ring in a fixed cyclic order (they are identical every step, so
// the producer warp simply runs ahead); the accumulator is double-buffered so the MMAs of block j+1 overlap the cell
// math of block j.  Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 cell math / selection.
// Bound: MUFU (8 tanh per hidden unit and step over 2 layers); see DESIGN.md.
#include <cstdlib>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace arcvae {

using bf16 = __nv_bfloat16;

#ifdef ARCVAE_RC_STAMPS
extern long long* g_rc_dbg;     // lstm_cluster.cu; profiles/scripts/sampler_stamps.py
#define SF_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x == 0 && t == 10) { unsigned long long _gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_gt)); p.dbg[(slot)] = (long long)_gt; } } while (0)
#else
#define SF_STAMP(slot) do { } while (0)
#endif

constexpr int SF_ROWS = 128;
constexpr int SF_BN = 192;
constexpr int SF_STAGES = 4;
constexpr int SF_THREADS = 320;
constexpr int SF_STAGE_BYTES = SF_BN * 128;   // [192 rows x 64 k] bf16 = 24 KB
constexpr int SF_PANEL = 16384;               // [128 rows x 64 k] bf16, SWIZZLE_128B
constexpr int SF_OUT_COL = 384;               // TMEM column of the logits accumulator (after 2 x 192)

struct SfMaps {
  CUtensorMap w[SF_MAX_LAYERS];   // w[0] = Tt [3H, 128] ; w[l] = Wxpb[l] [3H, H]
  CUtensorMap wout;               // Woutb [V, H]
};

struct SfParams {
  int B, H, V, VN, C, NL, max_length, end_token, multinomial;
  float temperature;
  uint64_t seed;
  const bf16* bpb[SF_MAX_LAYERS];    // [3H] tile-permuted bias (bf16 copy), l >= 1
  const float* bout;                 // [V]
  const float* cond;                 // [B, C]
  int32_t* tokens;                   // [B, max_length]
  int32_t* ended_count;              // [0] rows that emitted end_token, [1] max over rows of (end position + 1)
  long long* dbg;                    // stamps (STAMPS build only)
};

struct __align__(8) SfShared {
  uint64_t full[SF_STAGES];
  uint64_t empty[SF_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t out_full, out_empty;
  uint64_t a_ready;        // the one-hot tile of the step is written (8 arrivals)
  uint64_t p_ready[8];     // K panel kp of the layer output being produced is complete (8 arrivals per production)
  uint32_t tmem_base;
};

__device__ __forceinline__ void sf_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__device__ __forceinline__ void sf_store16(uint8_t* tile_row, int chunk, int row, const float (&v)[16]) {
  // 16 consecutive k-elements (two 16-byte chunks) of one row of a K-major SWIZZLE_128B panel
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(tile_row + (((chunk) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(tile_row + (((chunk + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

__global__ void __launch_bounds__(SF_THREADS, 1)
sampler_fused_kernel(const __grid_constant__ SfMaps maps, const SfParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KP = p.H >> 6;                      // 64-wide K panels of a hidden tile == 192-wide blocks per layer
  const int VKS = (p.V + 2 * p.C + 15) >> 4;    // K16 steps of layer 0 (vocabulary + cond hi/lo columns)
  const int VP = (VKS + 3) >> 2;                // K panels of the one-hot tile
  uint8_t* hA = smem;                                             // outputs of even layers
  uint8_t* hB = hA + (size_t)KP * SF_PANEL;                       // outputs of odd layers; one-hot tile of layer 0
  uint8_t* ring = hB + (size_t)(KP > VP ? KP : VP) * SF_PANEL;
  SfShared* sh = reinterpret_cast<SfShared*>(ring + (size_t)SF_STAGES * SF_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.B + SF_ROWS - 1) / SF_ROWS;
  const int NL = p.NL, T = p.max_length;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SF_STAGES; s++) {
      tc::mbar_init(&sh->full[s], 1);
      tc::mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      tc::mbar_init(&sh->acc_full[a], 1);
      tc::mbar_init(&sh->acc_empty[a], 8);
    }
    tc::mbar_init(&sh->out_full, 1);
    tc::mbar_init(&sh->out_empty, 8);
    tc::mbar_init(&sh->a_ready, 8);
    for (int k = 0; k < 8; k++) tc::mbar_init(&sh->p_ready[k], 8);
    tc::fence_barrier_init();
    for (int l = 0; l < NL; l++) tc::prefetch_tmap(&maps.w[l]);
    tc::prefetch_tmap(&maps.wout);
  }
  if (warp == 1) tc::tmem_alloc(&sh->tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ============================================================ producer: the same weight chunks every step
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int t = 0; t < T; t++) {
          for (int l = 0; l < NL; l++) {
            const int kpn = (l == 0) ? VP : KP;
            for (int j = 0; j < KP; j++)
              for (int kp = 0; kp < kpn; kp++) {
                tc::mbar_wait(&sh->empty[stage], phase ^ 1);
                tc::mbar_expect_tx(&sh->full[stage], SF_STAGE_BYTES);
                tc::tma_load_2d(ring + (size_t)stage * SF_STAGE_BYTES, &maps.w[l], &sh->full[stage], 64 * kp, SF_BN * j);
                if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
              }
          }
          for (int kp = 0; kp < KP; kp++) {
            tc::mbar_wait(&sh->empty[stage], phase ^ 1);
            tc::mbar_expect_tx(&sh->full[stage], (uint32_t)p.VN * 128);
            tc::tma_load_2d(ring + (size_t)stage * SF_STAGE_BYTES, &maps.wout, &sh->full[stage], 64 * kp, 0);
            if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_g = tc::make_idesc_bf16(SF_ROWS, SF_BN, false, false);
      const uint32_t idesc_o = tc::make_idesc_bf16(SF_ROWS, p.VN, false, false);
      int stage = 0;
      uint32_t phase = 0, acc_it = 0, out_it = 0, ar = 0, cons = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int t = 0; t < T; t++) {
          for (int l = 0; l < NL; l++) {
            if (l == 0) {
              tc::mbar_wait(&sh->a_ready, ar & 1);
              ar++;
              tc::tc_fence_after();
            }
            const uint8_t* A = (l == 0) ? hB : ((l & 1) ? hA : hB);   // layer l reads the output tile of layer l-1
            const int kpn = (l == 0) ? VP : KP;
            for (int j = 0; j < KP; j++) {
              const uint32_t acc = acc_it & 1;
              tc::mbar_wait(&sh->acc_empty[acc], ((acc_it >> 1) & 1) ^ 1);
              tc::tc_fence_after();
              SF_STAMP((l * 4 + j) * 4 + 0);
              const uint32_t d_tmem = tmem_base + acc * SF_BN;
              for (int kp = 0; kp < kpn; kp++) {
                if (l > 0 && j == 0) {
                  // K panel kp of this layer's input = block kp of the previous layer's output: start as soon as THAT
                  // block's cell math is done (the first block of a layer overlaps the last blocks of the layer below)
                  tc::mbar_wait(&sh->p_ready[kp], cons & 1);
                  tc::tc_fence_after();
                }
                tc::mbar_wait(&sh->full[stage], phase);
                tc::tc_fence_after();
                const int ks = (l == 0) ? min(4, VKS - 4 * kp) : 4;
                const uint32_t a_addr = tc::smem_u32(A + (size_t)kp * SF_PANEL);
                const uint32_t b_addr = tc::smem_u32(ring + (size_t)stage * SF_STAGE_BYTES);
                for (int k = 0; k < ks; k++)
                  tc::mma_bf16(d_tmem, tc::make_smem_desc(a_addr + 32 * k, 16, 1024),
                               tc::make_smem_desc(b_addr + 32 * k, 16, 1024), idesc_g, (kp > 0 || k > 0) ? 1u : 0u);
                tc::mma_commit(&sh->empty[stage]);
                if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
              }
              tc::mma_commit(&sh->acc_full[acc]);
              SF_STAMP((l * 4 + j) * 4 + 1);
              acc_it++;
              if (l > 0 && j == 0) cons++;
            }
          }
          // fc_out on the top layer's tile, panel by panel as the top layer's blocks complete
          const uint8_t* A = ((NL - 1) & 1) ? hB : hA;
          tc::mbar_wait(&sh->out_empty, (out_it & 1) ^ 1);
          tc::tc_fence_after();
          for (int kp = 0; kp < KP; kp++) {
            tc::mbar_wait(&sh->p_ready[kp], cons & 1);
            tc::tc_fence_after();
            tc::mbar_wait(&sh->full[stage], phase);
            tc::tc_fence_after();
            const uint32_t a_addr = tc::smem_u32(A + (size_t)kp * SF_PANEL);
            const uint32_t b_addr = tc::smem_u32(ring + (size_t)stage * SF_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc::mma_bf16(tmem_base + SF_OUT_COL, tc::make_smem_desc(a_addr + 32 * k, 16, 1024),
                           tc::make_smem_desc(b_addr + 32 * k, 16, 1024), idesc_o, (kp > 0 || k > 0) ? 1u : 0u);
            tc::mma_commit(&sh->empty[stage]);
            if (++stage == SF_STAGES) { stage = 0; phase ^= 1; }
          }
          tc::mma_commit(&sh->out_full);
          out_it++;
          cons++;
        }
      }
    }
  } else {
    // ============================================================ cell math / token selection (8 warps)
    const int q = warp & 3;                    // TMEM lane quarter
    const int hs = (warp - 2) >> 2;            // which 32 of a block's 64 hidden units
    const int row = q * 32 + lane;
    const int ct = threadIdx.x - 64;           // 0..255
    uint32_t acc_it = 0, out_it = 0;
    // byte offset of element (row, k) of the one-hot tile (K-major SWIZZLE_128B panels)
    auto oh = [&](int k) -> uint8_t* {
      return hB + (size_t)(k >> 6) * SF_PANEL + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
    };
    // Cell math of this warp's 32 units of a block in four chunks of 8 units, software-pipelined: the TMEM read of chunk
    // c+1 is in flight while chunk c runs through the MUFU pipe.  (With 16-unit chunks read and processed back to back all
    // eight warps alternated between a TMEM-bound and a MUFU-bound phase: 0.8 + 1.07 us per block instead of ~1.1.)  The
    // tcgen05.wait::ld takes a result of the previous chunk as a dummy operand so that the compiler cannot sink that
    // chunk's math below the wait.
    auto ld8x3 = [&](uint32_t taddr, int c0, uint32_t (&ri)[8], uint32_t (&rg)[8], uint32_t (&ro)[8]) {
      sf_tmem_ld8(taddr + c0, ri);
      sf_tmem_ld8(taddr + 64 + c0, rg);
      sf_tmem_ld8(taddr + 128 + c0, ro);
    };
    auto math8 = [&](const uint32_t (&ri)[8], const uint32_t (&rg)[8], const uint32_t (&ro)[8], const uint4& bi4, const uint4& bg4,
                     const uint4& bo4, uint8_t* drow, int c0) -> uint32_t {
      const __nv_bfloat162* pi = reinterpret_cast<const __nv_bfloat162*>(&bi4);
      const __nv_bfloat162* pg = reinterpret_cast<const __nv_bfloat162*>(&bg4);
      const __nv_bfloat162* po = reinterpret_cast<const __nv_bfloat162*>(&bo4);
      uint32_t pk[4];
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) {
        const float2 bi = __bfloat1622float2(pi[k2]), bg = __bfloat1622float2(pg[k2]), bo = __bfloat1622float2(po[k2]);
        const float ai0 = __uint_as_float(ri[2 * k2]) + bi.x, ai1 = __uint_as_float(ri[2 * k2 + 1]) + bi.y;
        const float ag0 = __uint_as_float(rg[2 * k2]) + bg.x, ag1 = __uint_as_float(rg[2 * k2 + 1]) + bg.y;
        const float ao0 = __uint_as_float(ro[2 * k2]) + bo.x, ao1 = __uint_as_float(ro[2 * k2 + 1]) + bo.y;
        const float h0 = sigmoid_approx_(ao0) * tanh_approx_(sigmoid_approx_(ai0) * tanh_approx_(ag0));
        const float h1 = sigmoid_approx_(ao1) * tanh_approx_(sigmoid_approx_(ai1) * tanh_approx_(ag1));
        __nv_bfloat162 t = __floats2bfloat162_rn(h0, h1);
        pk[k2] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(drow + (((c0 >> 3) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      return pk[0];
    };
    auto wait_ld_after = [&](uint32_t dep) { asm volatile("tcgen05.wait::ld.sync.aligned; // %0" ::"r"(dep) : "memory"); };
    // a_lo / a_hi: bias (bf16) of units [0,16) / [16,32) of this warp: [gate][2 x uint4]
    auto cell32 = [&](uint32_t taddr, int u0, const uint4 (&a_lo)[6], const uint4 (&a_hi)[6], uint8_t* drow) {
      uint32_t ri[8], rg[8], ro[8], si[8], sg[8], so[8];
      ld8x3(taddr, u0, ri, rg, ro);
      tc::tmem_ld_wait();
      ld8x3(taddr, u0 + 8, si, sg, so);
      uint32_t d = math8(ri, rg, ro, a_lo[0], a_lo[2], a_lo[4], drow, u0);
      wait_ld_after(d);
      ld8x3(taddr, u0 + 16, ri, rg, ro);
      d = math8(si, sg, so, a_lo[1], a_lo[3], a_lo[5], drow, u0 + 8);
      wait_ld_after(d);
      ld8x3(taddr, u0 + 24, si, sg, so);
      d = math8(ri, rg, ro, a_hi[0], a_hi[2], a_hi[4], drow, u0 + 16);
      wait_ld_after(d);
      math8(si, sg, so, a_hi[1], a_hi[3], a_hi[5], drow, u0 + 24);
    };
    auto load_bias = [&](const bf16* bl, int n, uint4 (&a)[6]) {     // 16 units x (i,g,o) starting at permuted row n
#pragma unroll
      for (int g = 0; g < 3; g++) {
        a[2 * g] = __ldg(reinterpret_cast<const uint4*>(bl + n + g * 64));
        a[2 * g + 1] = __ldg(reinterpret_cast<const uint4*>(bl + n + g * 64) + 1);
      }
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile * SF_ROWS + row;
      const bool valid = b < p.B;
      // cond enters layer 0 through the GEMM: columns V+2c, V+2c+1 of the one-hot tile hold the bf16 hi / lo split of
      // cond[b, c] and the matching rows of Tt hold wc[:, c]
      uint32_t cpk[4] = {0, 0, 0, 0};
      if (valid)
        for (int c = 0; c < p.C; c++) {
          const float cvf = __ldg(p.cond + (long)b * p.C + c);
          const __nv_bfloat16 hi = __float2bfloat16(cvf);
          const __nv_bfloat16 lo = __float2bfloat16(cvf - __bfloat162float(hi));
          const uint32_t pk = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
          if (c == 0) cpk[0] = pk; else if (c == 1) cpk[1] = pk; else if (c == 2) cpk[2] = pk; else cpk[3] = pk;
        }
      auto write_onehot = [&](int tokv) {
        *reinterpret_cast<uint16_t*>(oh(tokv)) = 0x3F80;                 // bf16 1.0
        for (int c = 0; c < p.C; c++) {
          const uint32_t pk = (c == 0) ? cpk[0] : (c == 1) ? cpk[1] : (c == 2) ? cpk[2] : cpk[3];
          *reinterpret_cast<uint16_t*>(oh(p.V + 2 * c)) = (uint16_t)(pk & 0xFFFF);
          *reinterpret_cast<uint16_t*>(oh(p.V + 2 * c + 1)) = (uint16_t)(pk >> 16);
        }
      };
      bool ended = false;
      // start token 0 (decoder_sampling.py:78)
      for (int i = ct; i < VP * (SF_PANEL / 16); i += 256) reinterpret_cast<uint4*>(hB)[i] = make_uint4(0, 0, 0, 0);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (hs == 0) write_onehot(0);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sh->a_ready);

      for (int t = 0; t < T; t++) {
        for (int l = 0; l < NL; l++) {
          uint8_t* dst = (l & 1) ? hB : hA;
          const bf16* bl = p.bpb[l];
          // additive terms (bias of layers >= 1; layer 0 has none), software-pipelined one half-block ahead so the
          // loads never sit between an accumulator becoming ready and its cell math
          uint4 a0[6], a1[6];
#pragma unroll
          for (int i = 0; i < 6; i++) { a0[i] = make_uint4(0, 0, 0, 0); a1[i] = a0[i]; }
          if (l > 0) load_bias(bl, hs * 32, a0);
          for (int j = 0; j < KP; j++) {
            const uint32_t acc = acc_it & 1;
            tc::mbar_wait(&sh->acc_full[acc], (acc_it >> 1) & 1);
            tc::tc_fence_after();
            if (threadIdx.x == 64) SF_STAMP((l * 4 + j) * 4 + 2);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * SF_BN;
            uint8_t* drow = dst + (size_t)j * SF_PANEL + row * 128;
            const int nb = j * SF_BN + hs * 32;
            if (l > 0) load_bias(bl, nb + 16, a1);
            cell32(taddr, hs * 32, a0, a1, drow);
            if (l > 0 && j + 1 < KP) load_bias(bl, nb + SF_BN, a0);
            tc::tc_fence_before();
            tc::fence_proxy_async();         // this block = K panel j of the next product: generic-proxy writes -> async proxy
            __syncwarp();
            if (lane == 0) {
              tc::mbar_arrive(&sh->acc_empty[acc]);
              tc::mbar_arrive(&sh->p_ready[j]);
            }
            if (threadIdx.x == 64) SF_STAMP((l * 4 + j) * 4 + 3);
            acc_it++;
          }
        }
        // ---- selection (decoder_sampling.py:110-123): thread per row, hs == 0 warps
        tc::mbar_wait(&sh->out_full, out_it & 1);
        tc::tc_fence_after();
        if (threadIdx.x == 64) SF_STAMP(60);
        int choice = 0;
        if (hs == 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + SF_OUT_COL;
          // argmax(softmax(x / T)) == argmax(x) for T > 0: compare the raw logits (lowest index wins ties, mx.argmax)
          float best = -INFINITY;
          int bi = 0;
#pragma unroll 1
          for (int c0 = 0; c0 < p.VN; c0 += 16) {
            uint32_t r[16];
            tc::tmem_ld16(taddr + c0, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; k++) {
              const int v = c0 + k;
              const float x = __uint_as_float(r[k]) + __ldg(p.bout + min(v, p.V - 1));
              if (v < p.V && x > best) { best = x; bi = v; }
            }
          }
          choice = bi;
          if (p.multinomial) {
            // inverse CDF over softmax(x / T) with ONE Philox uniform per (row, step); two passes over the accumulator
            const float inv_t = 1.0f / p.temperature;
            float s = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < p.VN; c0 += 16) {
              uint32_t r[16];
              tc::tmem_ld16(taddr + c0, r);
              tc::tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; k++) {
                const int v = c0 + k;
                const float x = __uint_as_float(r[k]) + __ldg(p.bout + min(v, p.V - 1));
                if (v < p.V) s += __expf((x - best) * inv_t);
              }
            }
            uint32_t r4[4];
            philox4x32(p.seed, (uint64_t)(valid ? b : 0), (uint64_t)t, r4);
            const float target = u01(r4[0]) * s;
            float run = 0.f;
            choice = -1;
#pragma unroll 1
            for (int c0 = 0; c0 < p.VN; c0 += 16) {   // no early exit: tcgen05.ld is warp-collective
              uint32_t r[16];
              tc::tmem_ld16(taddr + c0, r);
              tc::tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; k++) {
                const int v = c0 + k;
                const float x = __uint_as_float(r[k]) + __ldg(p.bout + min(v, p.V - 1));
                if (v < p.V) {
                  run += __expf((x - best) * inv_t);
                  if (choice < 0 && run >= target) choice = v;
                }
              }
            }
            if (choice < 0) choice = bi;    // rounding left the target just above the total mass
          }
          if (valid) {
            p.tokens[(long)b * T + t] = choice;
            if (choice == p.end_token && !ended) {
              ended = true;
              atomicMax(p.ended_count + 1, t + 1);
              __threadfence();
              atomicAdd(p.ended_count, 1);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sh->out_empty);
        out_it++;
        if (t + 1 < T) {
          // next step's one-hot tile (it aliases the tile the odd layers write, so it is rebuilt every step)
          for (int i = ct; i < VP * (SF_PANEL / 16); i += 256) reinterpret_cast<uint4*>(hB)[i] = make_uint4(0, 0, 0, 0);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (hs == 0) write_onehot(choice);
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&sh->a_ready);
          if (threadIdx.x == 64) SF_STAMP(61);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // nobody re-initialises hB for the next tile while a peer still selects
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---- operand preparation -------------------------------------------------------------------------------------------
// tile-permuted row p = j*192 + gi*64 + u  <->  compact natural column gi*H + j*64 + u   (gi: 0 = i, 1 = g, 2 = o)
__device__ __forceinline__ int sf_nat(int prow, int H) {
  const int j = prow / 192, rem = prow % 192;
  return (rem / 64) * H + j * 64 + (rem % 64);
}
// Tt[p, v] = table[v, nat(p)] for v < V, wc[nat(p), c] for v = V+2c, V+2c+1 (bf16, K padded to 128 with zeros);
// biasb[l][p] = bf16(bp_l[p]) for layers >= 1
struct SfBias { const float* b[SF_MAX_LAYERS]; };
__global__ void k_sf_prepare(const float* __restrict__ table, const float* __restrict__ wc, SfBias bp, int H, int V, int C,
                             int NL, bf16* __restrict__ Tt, bf16* __restrict__ biasb) {
  const int H3 = 3 * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H3 * 128; i += gridDim.x * blockDim.x) {
    const int prow = i >> 7, v = i & 127;
    float x = 0.f;
    if (v < V) x = table[(long)v * H3 + sf_nat(prow, H)];
    else if (v < V + 2 * C) x = wc[(long)sf_nat(prow, H) * C + ((v - V) >> 1)];
    Tt[i] = __float2bfloat16(x);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H3 * NL; i += gridDim.x * blockDim.x) {
    const int l = i / H3;
    if (l >= 1) biasb[i] = __float2bfloat16(bp.b[l][i - l * H3]);
  }
}
__global__ void k_sf_finalize(const int32_t* __restrict__ ended_count, int B, int max_length, int early_stopping,
                              int32_t* __restrict__ t_stop) {
  // the reference stops before the first step at which every row has ended (decoder_sampling.py:87-88)
  *t_stop = (early_stopping && ended_count[0] >= B) ? min(max_length, ended_count[1]) : max_length;
}

int make_tmap_bf16(CUtensorMap* m, const bf16* ptr, long rows, long cols, long ld, int box_cols, int box_rows);

bool sampler_fused_supported(const arcvae_dims& d, int precision) {
  if (precision != ARCVAE_PREC_BF16) return false;
  if (d.H % 64 != 0 || d.H > 256 || d.V + 2 * d.C > 128 || d.C > 4 || d.NL > SF_MAX_LAYERS || d.NL < 1) return false;
  return std::getenv("ARCVAE_NO_FUSED_SAMPLER") == nullptr;
}
size_t sampler_fused_prep_bytes(const arcvae_dims& d) {
  return align_up((size_t)3 * d.H * 128 * sizeof(bf16), 256) + align_up((size_t)3 * d.H * SF_MAX_LAYERS * sizeof(bf16), 256);
}

int sampler_fused_run(const arcvae_dims& d, const float* table, const float* wc, __nv_bfloat16* const* Wxpb,
                      float* const* bp, const __nv_bfloat16* Woutb, const float* bout, const float* cond, int B,
                      int max_length, float temperature, int early_stopping, int multinomial, uint64_t seed,
                      int32_t* tokens, int32_t* t_stop, int32_t* ended_count, void* prep, cudaStream_t st) {
  const int H = d.H, H3 = 3 * d.H;
  bf16* Tt = reinterpret_cast<bf16*>(prep);
  bf16* biasb = reinterpret_cast<bf16*>(reinterpret_cast<char*>(prep) + align_up((size_t)H3 * 128 * sizeof(bf16), 256));
  TimeScope ts(TIME_SAMPLER, st);
  SfBias sb{};
  for (int l = 1; l < d.NL && l < SF_MAX_LAYERS; l++) sb.b[l] = bp[l];
  k_sf_prepare<<<cdiv((long)H3 * 128, 256), 256, 0, st>>>(table, wc, sb, H, d.V, d.C, d.NL, Tt, biasb);
  ARCVAE_LAUNCHED();
  ARCVAE_CUDA(cudaMemsetAsync(ended_count, 0, 4 * sizeof(int32_t), st));
  if (max_length > 0) {
    SfMaps maps;
    SfParams p{};
    p.B = B; p.H = H; p.V = d.V; p.VN = ((d.V + 15) / 16) * 16; p.C = d.C; p.NL = d.NL; p.max_length = max_length;
    p.end_token = d.end_token; p.multinomial = multinomial; p.temperature = temperature; p.seed = seed;
    p.bout = bout; p.cond = cond; p.tokens = tokens; p.ended_count = ended_count;
#ifdef ARCVAE_RC_STAMPS
    p.dbg = g_rc_dbg;
#else
    p.dbg = nullptr;
#endif
    ARCVAE_TRY(make_tmap_bf16(&maps.w[0], Tt, H3, 128, 128, 64, SF_BN));
    for (int l = 1; l < SF_MAX_LAYERS; l++) {
      p.bpb[l] = l < d.NL ? biasb + (size_t)l * H3 : nullptr;
      if (l < d.NL) ARCVAE_TRY(make_tmap_bf16(&maps.w[l], Wxpb[l], H3, H, H, 64, SF_BN));
      else maps.w[l] = maps.w[0];
    }
    p.bpb[0] = nullptr;
    ARCVAE_TRY(make_tmap_bf16(&maps.wout, Woutb, d.V, H, H, 64, p.VN));
    const int KP = H / 64, VP = (((d.V + 2 * d.C + 15) / 16) + 3) / 4;
    const size_t smem = (size_t)(KP + (KP > VP ? KP : VP)) * SF_PANEL + (size_t)SF_STAGES * SF_STAGE_BYTES +
                        sizeof(SfShared) + 1024;
    ARCVAE_REQUIRE(smem <= 227 * 1024, "fused sampler shared-memory budget");
    const int num_sms = device_sm_count();
    if (first_use_on_device(ONCE_SAMPLER))
      ARCVAE_CUDA(cudaFuncSetAttribute(sampler_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int ntiles = cdiv(B, SF_ROWS);
    sampler_fused_kernel<<<ntiles < num_sms ? ntiles : num_sms, SF_THREADS, smem, st>>>(maps, p);
    ARCVAE_LAUNCHED();
  }
  k_sf_finalize<<<1, 1, 0, st>>>(ended_count, B, max_length, early_stopping, t_stop);
  ARCVAE_LAUNCHED();
  return 0;
}

}  // namespace arcvae
