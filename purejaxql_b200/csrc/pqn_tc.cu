// tcgen05 / TMEM / TMA path for the dense contractions of the Q-network
// (sm_100a only): one warp-specialised, persistent GEMM kernel used for
//   forward   Z  = H1 [rows,1024] . W1 [1024,128]        (A K-major,  B MN-major)
//   wgrad     dW = H1^T [1024,rows] . dZ [rows,128]      (A MN-major, B MN-major)
//   dgrad     dX = dZ [rows,128] . W1^T [128,1024]       (A K-major,  B K-major)
// with fp32 accuracy from 3xTF32 error compensation: every operand x is fed as
// hi = x (the tensor core reads the top 19 bits) and lo = x - trunc_tf32(x),
// and D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi accumulates in fp32 in TMEM.
//
// Reference arithmetic: nn.Dense(128) of CNN (purejaxql/pqn_minatar.py:48) and
// its autodiff; XLA itself runs these as TF32 tensor-core GEMMs on GPU.
//
// Pipeline (per CTA, 192 threads):
//   warp 0   TMA producer   cp.async.bulk.tensor (SWIZZLE_128B boxes) -> smem ring
//   warp 1   MMA issuer     tcgen05.mma.cta_group::1.kind::tf32 (one thread), TMEM accumulators
//   warps 2-5 epilogue      tcgen05.ld 32x32b -> registers -> fused epilogue -> global
//   warps 6-7 converters    (a_lo_inline) A_lo tile = A tile - trunc_tf32(A tile), smem -> smem, so that the
//                           big activation operand is read from HBM once instead of as a (hi, lo) pair
// Barriers: full[stage]/empty[stage] (TMA <-> MMA), lo_full[stage] (converters -> MMA),
//           main_full/main_empty[2], corr_full/corr_empty[2] (MMA <-> epilogue).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pqn_b200.h"
#include "api_common.h"
#include "tc_common.cuh"

namespace pqn {
namespace tc {

// ---------------------------------------------------------------------------
// epilogues: each epilogue thread owns one row of the 128x128 tile, already
// promoted to fp32 registers (acc[128], fully unrolled static indexing)
// ---------------------------------------------------------------------------
// Coalesced tile-row-block store: the warp's 32 rows x 32 columns chunk goes through a padded shared-memory
// stage so that 8 lanes write one 128-byte row segment (4 rows per store instruction) instead of 32 lanes
// writing 32 different rows.  `vals` = this lane's row, columns [0,32) of the chunk.  MASKED: multiply by
// (mask > 0) read with the same coalesced addressing (dgrad's fused ReLU mask; may alias dst).
constexpr int STG_LD = 36;  // floats per staged row (16-byte aligned, conflict-free for 128-bit accesses)

template <bool MASKED>
__device__ __forceinline__ void store_chunk_coalesced(float* stage, const float (&vals)[32], int lane, float* dst,
                                                      uint32_t mask_bits, int64_t ld, int m_base, int M) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * STG_LD + 4 * j) =
        make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
  __syncwarp();
  const int r_in = lane >> 3, c4 = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + r_in;
    if (m_base + r < M) {
      float4 o = *reinterpret_cast<const float4*>(stage + r * STG_LD + 4 * c4);
      const int64_t off = (int64_t)r * ld + 4 * c4;
      if (MASKED) {
        const uint32_t b = mask_bits >> (it * 4);
        o.x = (b & 1u) ? o.x : 0.f; o.y = (b & 2u) ? o.y : 0.f;
        o.z = (b & 4u) ? o.z : 0.f; o.w = (b & 8u) ? o.w : 0.f;
      }
      *reinterpret_cast<float4*>(dst + off) = o;
    }
  }
  __syncwarp();
}

// ReLU mask of dgrad, fetched BEFORE the tile's MMAs are awaited (it does not depend on them) so the HBM latency
// overlaps the tensor-core work: bit (it*4 + j) of bits[c] = (mask[row it*4 + lane/8][c*32 + 4*(lane%8) + j] > 0),
// i.e. exactly the elements this lane stores in store_chunk_coalesced.
__device__ __forceinline__ void prefetch_mask_bits(const EpiParams& ep, int seed, int m_base, int n0, int lane, int M,
                                                   uint32_t (&bits)[4]) {
  const float* msk = ep.mask + (int64_t)seed * ep.out_seed_stride + (int64_t)m_base * ep.ld_out + n0;
  const int r_in = lane >> 3, c4 = lane & 7;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 h[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + r_in;
      h[it] = (m_base + r < M) ? *reinterpret_cast<const float4*>(msk + (int64_t)r * ep.ld_out + c * 32 + 4 * c4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t b = 0u;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      b |= ((h[it].x > 0.f ? 1u : 0u) | (h[it].y > 0.f ? 2u : 0u) | (h[it].z > 0.f ? 4u : 0u) | (h[it].w > 0.f ? 8u : 0u))
           << (it * 4);
    bits[c] = b;
  }
}

// `m_base` = first row of this warp's 32-row block; the lane's own row is m_base + lane.
// Shared-memory copy of the per-seed epilogue parameters (LN epilogues), staged once per tile by the 128 epilogue
// threads: every thread needs all of them, and 128-bit broadcast reads cost a quarter of the per-element loads.
constexpr int SP_B = 0, SP_SC = 128, SP_BI = 256, SP_HW = 384 /* [PQN_TC_MAX_A][128] */,
              SP_HB = SP_HW + PQN_TC_MAX_A * 128, SP_FLOATS = SP_HB + PQN_TC_MAX_A;
static_assert(SP_FLOATS == 384 + 8 * 128 + 8, "matches the TC_SMEM_BYTES budget in tc_common.cuh");

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int EPI>
__device__ __forceinline__ void stage_epi_params(const EpiParams& ep, float* sp, int seed, int et /*0..127*/) {
  if constexpr (EPI == EPI_LN_TRAIN || EPI == EPI_LN_HEAD) {
    const float* __restrict__ prm = ep.params + (int64_t)seed * ep.P;
    epi_bar_sync();  // everyone is done with the previous tile's parameters
    sp[SP_B + et] = __ldg(prm + ep.off_b + et);
    sp[SP_SC + et] = __ldg(prm + ep.off_scale + et);
    sp[SP_BI + et] = __ldg(prm + ep.off_bias + et);
    if constexpr (EPI == EPI_LN_HEAD) {
      for (int a = 0; a < ep.A; ++a) sp[SP_HW + a * 128 + et] = __ldg(prm + ep.off_hw + (int64_t)et * ep.A + a);
      if (et < ep.A) sp[SP_HB + et] = __ldg(prm + ep.off_hb + et);
    }
    epi_bar_sync();
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_row(const EpiParams& ep, float (&acc)[128], float* stage, const float* sp,
                                             int lane, int seed, int m_base, int n0, int M,
                                             const uint32_t (&mask_bits)[4]) {
  const int m = m_base + lane;
  const bool row_ok = m < M;
  if constexpr (EPI == EPI_STORE || EPI == EPI_RELU_MASK || EPI == EPI_RELU_BITS) {
    // no __restrict__: dgrad runs in place (out == mask)
    float* out = ep.out + (int64_t)seed * ep.out_seed_stride + (int64_t)m_base * ep.ld_out + n0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = acc[c * 32 + j];
        // EPI_RELU_BITS: mask_bits = this lane's own row, one word per 32-column chunk
        if (EPI == EPI_RELU_BITS) v[j] = ((mask_bits[c] >> j) & 1u) ? v[j] : 0.f;
      }
      store_chunk_coalesced<EPI == EPI_RELU_MASK>(stage, v, lane, out + c * 32, mask_bits[c], ep.ld_out, m_base, M);
    }
  } else {
    // bias + LayerNorm(128) + ReLU, then either (h, xhat, rstd) or the fused Q-head
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < 32; ++j4) {
      const float4 b = *reinterpret_cast<const float4*>(sp + SP_B + 4 * j4);
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = 4 * j4 + e;
        acc[j] += bb[e];
        s1 += acc[j];
        s2 = fmaf(acc[j], acc[j], s2);
      }
    }
    const float mean = s1 * (1.0f / 128.f);
    const float var = fmaxf(s2 * (1.0f / 128.f) - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-6f);
    const int64_t grow = (int64_t)seed * ep.rows + m;
    if constexpr (EPI == EPI_LN_TRAIN) {
      float* hbase = ep.H + ((int64_t)seed * ep.rows + m_base) * 128;
      float* xbase = ep.XHAT + ((int64_t)seed * ep.rows + m_base) * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float xh[32], h[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sp + SP_SC + c * 32 + 4 * j4);
          const float4 b4 = *reinterpret_cast<const float4*>(sp + SP_BI + c * 32 + 4 * j4);
          const float ss[4] = {s4.x, s4.y, s4.z, s4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            xh[j] = (acc[c * 32 + j] - mean) * rstd;
            h[j] = fmaxf(xh[j] * ss[e] + bb[e], 0.f);
          }
        }
        store_chunk_coalesced<false>(stage, h, lane, hbase + c * 32, 0u, 128, m_base, M);
        store_chunk_coalesced<false>(stage, xh, lane, xbase + c * 32, 0u, 128, m_base, M);
      }
      if (row_ok) ep.RSTD[grow] = rstd;
    } else {  // EPI_LN_HEAD
      float q[PQN_TC_MAX_A];
#pragma unroll
      for (int a = 0; a < PQN_TC_MAX_A; ++a) q[a] = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 32; ++j4) {
        const float4 s4 = *reinterpret_cast<const float4*>(sp + SP_SC + 4 * j4);
        const float4 b4 = *reinterpret_cast<const float4*>(sp + SP_BI + 4 * j4);
        float h[4];
        h[0] = fmaxf((acc[4 * j4 + 0] - mean) * rstd * s4.x + b4.x, 0.f);
        h[1] = fmaxf((acc[4 * j4 + 1] - mean) * rstd * s4.y + b4.y, 0.f);
        h[2] = fmaxf((acc[4 * j4 + 2] - mean) * rstd * s4.z + b4.z, 0.f);
        h[3] = fmaxf((acc[4 * j4 + 3] - mean) * rstd * s4.w + b4.w, 0.f);
#pragma unroll
        for (int a = 0; a < PQN_TC_MAX_A; ++a)
          if (a < ep.A) {
            const float4 w4 = *reinterpret_cast<const float4*>(sp + SP_HW + a * 128 + 4 * j4);
            q[a] = fmaf(h[0], w4.x, q[a]); q[a] = fmaf(h[1], w4.y, q[a]);
            q[a] = fmaf(h[2], w4.z, q[a]); q[a] = fmaf(h[3], w4.w, q[a]);
          }
      }
      if (row_ok) {
#pragma unroll
        for (int a = 0; a < PQN_TC_MAX_A; ++a)
          if (a < ep.A) ep.Q[grow * ep.A + a] = q[a] + sp[SP_HB + a];
      }
    }
  }
}

// acc[0..128) (+)= the 128 fp32 columns of this thread's TMEM lane at `row_addr`.  The loads are issued back to
// back and awaited once (FIRST: all four 32-column loads straight into acc; otherwise two at a time, 64 staging
// registers) -- the per-load round trip to TMEM was the longest part of the epilogue's dependent chain.
template <bool FIRST, int NC = 128>
__device__ __forceinline__ void tmem_accumulate_row(uint32_t row_addr, float (&acc)[NC], float scale = 1.0f) {
  if constexpr (FIRST && NC == 128) {
    uint32_t v0[32], v1[32], v2[32], v3[32];
    tmem_ld_32x32b_x32(row_addr, v0);
    tmem_ld_32x32b_x32(row_addr + 32, v1);
    tmem_ld_32x32b_x32(row_addr + 64, v2);
    tmem_ld_32x32b_x32(row_addr + 96, v3);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      acc[j] = __uint_as_float(v0[j]); acc[32 + j] = __uint_as_float(v1[j]);
      acc[64 + j] = __uint_as_float(v2[j]); acc[96 + j] = __uint_as_float(v3[j]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < NC / 64; ++c) {
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(row_addr + c * 64, v0);
      tmem_ld_32x32b_x32(row_addr + c * 64 + 32, v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (FIRST) {
          acc[c * 64 + j] = __uint_as_float(v0[j]);
          acc[c * 64 + 32 + j] = __uint_as_float(v1[j]);
        } else {
          acc[c * 64 + j] = fmaf(__uint_as_float(v0[j]), scale, acc[c * 64 + j]);   // scale == 1: exact add
          acc[c * 64 + 32 + j] = fmaf(__uint_as_float(v1[j]), scale, acc[c * 64 + 32 + j]);
        }
      }
    }
  }
}

// SPLIT epilogue (fp16 kernels with a store epilogue: wgrad, dgrad): EIGHT epilogue warps instead of four -- warps 2-5
// take columns 0..63 of the tile, warps 6-9 columns 64..127 (a warp may only touch the TMEM lane quadrant warp % 4, and
// both sets cover the four quadrants).  Half the registers per thread and twice the warps in flight: the dgrad
// epilogue (K = 128, one tile every ~2500 MMA clocks) was latency-bound with one warp per scheduler (ncu r2c: issue
// active 20 %, 4468 warp-instructions per tile).  The 32 x 32 staging tile of a warp is XOR-swizzled instead of padded
// (8 x 4 KB fit next to the three operand stages).
__device__ __forceinline__ void store_chunk_swz(float* stage, const float (&vals)[32], int lane, float* dst, int64_t ld,
                                                int m_base, int M) {
#pragma unroll
  for (int j = 0; j < 8; ++j)   // row = lane, 16-byte chunk j -> physical chunk j ^ (lane & 7)
    *reinterpret_cast<float4*>(stage + lane * 32 + 4 * (j ^ (lane & 7))) =
        make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
  __syncwarp();
  const int r_in = lane >> 3, c4 = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + r_in;
    if (m_base + r < M) {
      const float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + 4 * (c4 ^ (r & 7)));
      *reinterpret_cast<float4*>(dst + (int64_t)r * ld + 4 * c4) = o;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// the kernel
//
// Accuracy: the tensor core truncates (round-toward-zero) when it accumulates into TMEM, so a long in-TMEM
// chain drifts (measured: 3e-5 relative at K=1024).  Two-level accumulation fixes it:
//   * the correction terms A_lo.B_hi + A_hi.B_lo (2^-11 of the result) get their own TMEM accumulator for the
//     whole K — their truncation error is negligible at that magnitude;
//   * the main A_hi.B_hi chain is cut every TC_PROMOTE k-blocks (16 MMAs): the partial is added to fp32
//     registers by the epilogue warps (round-to-nearest FADD) and the MMA warp continues into the other TMEM
//     buffer with accumulate=0.
// TMEM columns: main[2] at 0/128, corr[2] at 256/384.
// ---------------------------------------------------------------------------
template <int EPI, bool F16>
struct TcSplit {
  static constexpr bool value = F16 && (EPI == EPI_STORE || EPI == EPI_RELU_BITS);
  static constexpr int threads = value ? TC_THREADS + 64 : TC_THREADS;
  // operand ring + align slack + barriers + staging (split: 8 x [32][32]; else 4 x [32][36] + epilogue parameters)
  static constexpr int smem = TC_STAGES * TC_STAGE_BYTES + 1024 + 256 +
                              (value ? 8 * 32 * 32 * 4 : 4 * 32 * 36 * 4 + (384 + 8 * 128 + 8) * 4);
};

template <int A_MN, int B_MN, int EPI, bool F16 = false>
__global__ void __launch_bounds__((TcSplit<EPI, F16>::threads), 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const GemmShape gs, const EpiParams ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring (swizzle atoms), then barriers
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;                        // [TC_STAGES]   TMA -> MMA
  uint64_t* empty = bars + TC_STAGES;           // [TC_STAGES]   MMA -> TMA
  uint64_t* main_full = bars + 2 * TC_STAGES;   // [2]           MMA -> epilogue (main partial ready)
  uint64_t* main_empty = main_full + 2;         // [2]           epilogue -> MMA
  uint64_t* corr_full = main_empty + 2;         // [2]
  uint64_t* corr_empty = corr_full + 2;         // [2]
  uint64_t* lo_full = corr_empty + 2;           // [TC_STAGES]   converters -> MMA (A_lo tile written)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_full + TC_STAGES);
  float* stage_all = reinterpret_cast<float*>(smem_al + TC_STAGES * TC_STAGE_BYTES + 256);  // 4 x [32][STG_LD]
  float* sp_all = stage_all + 4 * 32 * STG_LD;                                              // [SP_FLOATS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_b_hi);
    if (gs.split3) { if (!gs.a_lo_inline) prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_b_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&lo_full[i], TC_CONV_THREADS);
    }
    constexpr int EPI_WARPS = TcSplit<EPI, F16>::value ? 8 : 4;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&main_full[i], 1); mbar_init(&main_empty[i], EPI_WARPS);
      mbar_init(&corr_full[i], 1); mbar_init(&corr_empty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // split-K: tile index -> (seed, m/n tile, k range).  k_split == 1 is the plain case.
  const int ksplit = gs.k_split > 1 ? gs.k_split : 1;
  const int kb_per = (gs.k_blocks + ksplit - 1) / ksplit;
  const int tiles_per_seed = gs.m_tiles * gs.n_tiles * ksplit;
  const int num_tiles = tiles_per_seed * gs.S;
  auto decode = [&](int tile, int& seed, int& m0, int& n0, int& kb0, int& kbn) {
    seed = tile / tiles_per_seed;
    const int rem = tile - seed * tiles_per_seed;
    const int ks = rem % ksplit, mn = rem / ksplit;
    m0 = (mn / gs.n_tiles) * 128; n0 = (mn % gs.n_tiles) * 128;
    kb0 = ks * kb_per;
    kbn = min(kb_per, gs.k_blocks - kb0);
    if (kbn < 0) kbn = 0;
    return ks;
  };

  // F16: fp16 operand planes (hi, lo'), 64-element k-blocks, A_lo' always comes from memory (no converter warps)
  const bool a_lo_tma = gs.split3 && (F16 || !gs.a_lo_inline);
  const bool a_lo_conv = !F16 && gs.split3 && gs.a_lo_inline;
  constexpr int BKE = F16 ? TC_BK16 : TC_BK;              // elements per k-block
  constexpr int MNB = F16 ? 2 : 4;                        // MN-major TMA boxes per 128-wide tile
  constexpr int MNB_ELEMS = F16 ? 64 : 32, MNB_BYTES = F16 ? 8192 : 4096;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int seed, m0, n0, kb0, kbn;
        decode(tile, seed, m0, n0, kb0, kbn);
        for (int kb = kb0; kb < kb0 + kbn; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          const uint32_t sb = smem_base + stage * TC_STAGE_BYTES;
          mbar_expect_tx(&full[stage], gs.split3 ? (a_lo_tma ? TC_STAGE_BYTES : 3 * TC_TILE_BYTES) : TC_STAGE_BYTES / 2);
          const int k0 = kb * BKE;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < MNB; ++j) {
              tma_load_3d(sb + TC_A_HI + j * MNB_BYTES, &tm_a_hi, &full[stage], m0 + MNB_ELEMS * j, k0, seed);
              if (a_lo_tma) tma_load_3d(sb + TC_A_LO + j * MNB_BYTES, &tm_a_lo, &full[stage], m0 + MNB_ELEMS * j, k0, seed);
            }
          } else {
            tma_load_3d(sb + TC_A_HI, &tm_a_hi, &full[stage], k0, m0, seed);
            if (a_lo_tma) tma_load_3d(sb + TC_A_LO, &tm_a_lo, &full[stage], k0, m0, seed);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < MNB; ++j) {
              tma_load_3d(sb + TC_B_HI + j * MNB_BYTES, &tm_b_hi, &full[stage], n0 + MNB_ELEMS * j, k0, seed);
              if (gs.split3) tma_load_3d(sb + TC_B_LO + j * MNB_BYTES, &tm_b_lo, &full[stage], n0 + MNB_ELEMS * j, k0, seed);
            }
          } else {
            tma_load_3d(sb + TC_B_HI, &tm_b_hi, &full[stage], k0, n0, seed);
            if (gs.split3) tma_load_3d(sb + TC_B_LO, &tm_b_lo, &full[stage], k0, n0, seed);
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = F16 ? make_idesc_f16(128, 128, A_MN, B_MN) : make_idesc_tf32(128, 128, A_MN, B_MN);
      auto sdesc_a = [](uint32_t tile, int ks) { return F16 ? make_sdesc16<A_MN>(tile, ks) : make_sdesc<A_MN>(tile, ks); };
      auto sdesc_b = [](uint32_t tile, int ks) { return F16 ? make_sdesc16<B_MN>(tile, ks) : make_sdesc<B_MN>(tile, ks); };
      auto umma = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if (F16) umma_f16(d, da, db, id, acc); else umma_tf32(d, da, db, id, acc);
      };
      int stage = 0;
      uint32_t phase = 0;
      int mb = 0, cb = 0;
      uint32_t mb_phase = 0, cb_phase = 0;
      // Short-K mode (k_blocks <= TC_PROMOTE, e.g. dgrad with K = 128): the whole tile is one promotion chunk, so the
      // correction terms share the main accumulator (a chain of <= 48 MMAs keeps the truncation drift ~1e-6) and the
      // four 128-column TMEM regions form one ring of accumulators: the epilogue reads each tile once.
      // (tf32 only: the f16 path keeps the cross products in units of 2^-11, so they need their own accumulator)
      const bool single_acc = !F16 && gs.split3 && gs.k_blocks <= TC_PROMOTE;
      if (single_acc) {
        int ab = 0;
        uint32_t ab_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          uint64_t* e_bar = ab < 2 ? &main_empty[ab] : &corr_empty[ab - 2];
          uint64_t* f_bar = ab < 2 ? &main_full[ab] : &corr_full[ab - 2];
          mbar_wait(e_bar, ab_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_acc = tmem_base + ab * 128;
          bool first = true;
          int seed_, m0_, n0_, kb0_, kbn_;
          decode(tile, seed_, m0_, n0_, kb0_, kbn_);
          for (int kb = 0; kb < kbn_; ++kb) {
            mbar_wait(&full[stage], phase);
            if (a_lo_conv) mbar_wait(&lo_full[stage], phase);
            tcgen05_fence_after();
            const uint32_t sb = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t a_hi = sdesc_a(sb + TC_A_HI, ks), a_lo = sdesc_a(sb + TC_A_LO, ks);
              const uint64_t b_hi = sdesc_b(sb + TC_B_HI, ks), b_lo = sdesc_b(sb + TC_B_LO, ks);
              umma(d_acc, a_lo, b_hi, idesc, first ? 0u : 1u);
              umma(d_acc, a_hi, b_lo, idesc, 1u);
              umma(d_acc, a_hi, b_hi, idesc, 1u);
              first = false;
            }
            umma_commit(&empty[stage]);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(f_bar);
          if (++ab == 4) { ab = 0; ab_phase ^= 1u; }
        }
      } else
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t d_corr = tmem_base + 256 + cb * 128;
        if (gs.split3) {
          mbar_wait(&corr_empty[cb], cb_phase ^ 1u);
          tcgen05_fence_after();
        }
        bool first_corr = true, first_main = true;
        int seed_, m0_, n0_, kb0_, kbn_;
        decode(tile, seed_, m0_, n0_, kb0_, kbn_);
        for (int kb = 0; kb < kbn_; ++kb) {
          if (kb % TC_PROMOTE == 0) {
            mbar_wait(&main_empty[mb], mb_phase ^ 1u);
            tcgen05_fence_after();
            first_main = true;
          }
          const uint32_t d_main = tmem_base + mb * 128;
          mbar_wait(&full[stage], phase);
          if (a_lo_conv) mbar_wait(&lo_full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sb = smem_base + stage * TC_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = sdesc_a(sb + TC_A_HI, ks);
            const uint64_t b_hi = sdesc_b(sb + TC_B_HI, ks);
            if (gs.split3) {
              const uint64_t a_lo = sdesc_a(sb + TC_A_LO, ks);
              const uint64_t b_lo = sdesc_b(sb + TC_B_LO, ks);
              umma(d_corr, a_lo, b_hi, idesc, first_corr ? 0u : 1u);
              umma(d_corr, a_hi, b_lo, idesc, 1u);
              first_corr = false;
            }
            umma(d_main, a_hi, b_hi, idesc, first_main ? 0u : 1u);
            first_main = false;
          }
          umma_commit(&empty[stage]);  // smem slot free once these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
          if ((kb + 1) % TC_PROMOTE == 0 || kb == kbn_ - 1) {
            umma_commit(&main_full[mb]);  // main partial ready for promotion
            if (++mb == 2) { mb = 0; mb_phase ^= 1u; }
          }
        }
        if (gs.split3) {
          umma_commit(&corr_full[cb]);
          if (++cb == 2) { cb = 0; cb_phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 6 && !TcSplit<EPI, F16>::value) {
    // ===================== converter warps (6..7) =====================
    // The split is elementwise, so it is independent of the (swizzled) tile layout: byte i of the A_hi tile maps
    // to byte i of the A_lo tile.  TMA zero-fills out-of-range rows, whose lo is 0 as well.
    if (a_lo_conv) {
      const int ct = threadIdx.x - 6 * 32;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int seed_, m0_, n0_, kb0_, kbn_;
        decode(tile, seed_, m0_, n0_, kb0_, kbn_);
        for (int kb = 0; kb < kbn_; ++kb) {
          mbar_wait(&full[stage], phase);
          const float4* src = reinterpret_cast<const float4*>(smem_al + stage * TC_STAGE_BYTES + TC_A_HI);
          float4* dst = reinterpret_cast<float4*>(smem_al + stage * TC_STAGE_BYTES + TC_A_LO);
          constexpr int PER = TC_TILE_BYTES / 16 / TC_CONV_THREADS;  // 16 float4 per thread
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float4 v[PER / 2];
#pragma unroll
            for (int i = 0; i < PER / 2; ++i) v[i] = src[(half * (PER / 2) + i) * TC_CONV_THREADS + ct];
#pragma unroll
            for (int i = 0; i < PER / 2; ++i)
              dst[(half * (PER / 2) + i) * TC_CONV_THREADS + ct] =
                  make_float4(tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
          }
          fence_proxy_async_smem();
          mbar_arrive(&lo_full[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..5; split mode: 2..9) =====================
    constexpr bool SPLIT = TcSplit<EPI, F16>::value;
    constexpr int NC = SPLIT ? 64 : 128;   // accumulator columns per thread
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (warp id % 4)
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int col_off = SPLIT ? (warp >= 6 ? 64 : 0) : 0;
    int mb = 0, cb = 0, ab = 0;
    uint32_t mb_phase = 0, cb_phase = 0, ab_phase = 0;
    const bool single_acc = !F16 && gs.split3 && gs.k_blocks <= TC_PROMOTE;  // see the MMA issuer
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int seed, m0, n0, kb0, kbn;
      const int ks = decode(tile, seed, m0, n0, kb0, kbn);
      const int partials = (kbn + TC_PROMOTE - 1) / TC_PROMOTE;
      uint32_t mask_bits[4] = {0u, 0u, 0u, 0u};
      if constexpr (EPI == EPI_RELU_MASK) prefetch_mask_bits(ep, seed, m0 + quad * 32, n0, lane, gs.M, mask_bits);
      if constexpr (EPI == EPI_RELU_BITS) {
        // packed ReLU mask written by the conv forward: 16 bytes per (row, 128-column tile)
        const int m = m0 + quad * 32 + lane;
        if (m < gs.M) {
          const uint32_t* bp = ep.relu_bits + ((int64_t)seed * ep.rows + m) * (ep.ld_out >> 5) + (n0 >> 5);
          if constexpr (SPLIT) {
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(bp + (col_off >> 5)));
            mask_bits[0] = b.x; mask_bits[1] = b.y;
          } else {
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(bp));
            mask_bits[0] = b.x; mask_bits[1] = b.y; mask_bits[2] = b.z; mask_bits[3] = b.w;
          }
        }
      }
      if constexpr (!SPLIT) stage_epi_params<EPI>(ep, sp_all, seed, threadIdx.x - 64);
      float acc[NC];
      if (single_acc) {
        uint64_t* e_bar = ab < 2 ? &main_empty[ab] : &corr_empty[ab - 2];
        uint64_t* f_bar = ab < 2 ? &main_full[ab] : &corr_full[ab - 2];
        mbar_wait(f_bar, ab_phase);
        tcgen05_fence_after();
        tmem_accumulate_row<true, NC>(tmem_base + ab * 128 + col_off + lane_off, acc);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(e_bar);
        if (++ab == 4) { ab = 0; ab_phase ^= 1u; }
      } else {
      for (int pi = 0; pi < partials; ++pi) {
        mbar_wait(&main_full[mb], mb_phase);
        tcgen05_fence_after();
        if (pi == 0) tmem_accumulate_row<true, NC>(tmem_base + mb * 128 + col_off + lane_off, acc);
        else tmem_accumulate_row<false, NC>(tmem_base + mb * 128 + col_off + lane_off, acc);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&main_empty[mb]);
        if (++mb == 2) { mb = 0; mb_phase ^= 1u; }
      }
      if (gs.split3) {
        mbar_wait(&corr_full[cb], cb_phase);
        tcgen05_fence_after();
        tmem_accumulate_row<false, NC>(tmem_base + 256 + cb * 128 + col_off + lane_off, acc, F16 ? TC_LO_INV : 1.0f);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&corr_empty[cb]);
        if (++cb == 2) { cb = 0; cb_phase ^= 1u; }
      }
      }
      if (F16 && ep.out_scale != 0.f) {   // undo the power-of-two pre-scaling of an operand (exact)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[j] *= ep.out_scale;
      }
      if constexpr (SPLIT) {
        const int m_base = m0 + quad * 32;
        float* out = ep.out + (int64_t)ks * ep.split_stride + (int64_t)seed * ep.out_seed_stride +
                     (int64_t)m_base * ep.ld_out + n0 + col_off;
        float* stg = stage_all + (warp - 2) * 32 * 32;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = acc[c * 32 + j];
            if (EPI == EPI_RELU_BITS) v[j] = ((mask_bits[c] >> j) & 1u) ? v[j] : 0.f;
          }
          store_chunk_swz(stg, v, lane, out + c * 32, ep.ld_out, m_base, gs.M);
        }
      } else {
        epilogue_row<EPI>(ep, acc, stage_all + (warp - 2) * 32 * STG_LD, sp_all, lane, seed, m0 + quad * 32, n0, gs.M,
                          mask_bits);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D fp32 tensor [seeds][mid][inner] with a {32, box_mid, 1} box; SWIZZLE_128B for K-major operand tiles,
// SWIZZLE_128B_ATOM_32B for MN-major ones (see make_sdesc).
int make_tmap(CUtensorMap* tm, const float* base, uint64_t inner, uint64_t mid, uint64_t seeds, uint64_t mid_stride_elems,
              uint64_t seed_stride_elems, uint32_t box_mid, int mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(PQN_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {inner, mid, seeds};
  cuuint64_t strides[2] = {mid_stride_elems * 4, seed_stride_elems * 4};
  cuuint32_t box[3] = {32, box_mid, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PQN_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return PQN_OK;
}

static int num_sms() { return device_sm_count(); }

template <int A_MN, int B_MN, int EPI, bool F16 = false>
static int launch_t(const CUtensorMap* t, const GemmShape& gs, const EpiParams& ep, cudaStream_t st, int kid) {
  auto kfn = tc_gemm_kernel<A_MN, B_MN, EPI, F16>;
  constexpr int SMEM = TcSplit<EPI, F16>::smem, THREADS = TcSplit<EPI, F16>::threads;
  static_assert(SMEM <= 227 * 1024, "dynamic shared memory of the GEMM kernel exceeds the per-CTA limit");
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
    return check_launch("tc_gemm(cudaFuncSetAttribute)");
  const int tiles = gs.m_tiles * gs.n_tiles * gs.S * (gs.k_split > 1 ? gs.k_split : 1);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  {
    LaunchScope _ls(kid < 0 ? (int)K_TC_GEMM : kid, st);
    kfn<<<grid, THREADS, SMEM, st>>>(t[0], t[1], t[2], t[3], gs, ep);
  }
  return check_launch("tc_gemm");
}

int launch_gemm(int a_mn, int b_mn, int epi, const CUtensorMap* t, const GemmShape& gs, const EpiParams& ep,
                cudaStream_t st, int kernel_id) {
#define PQN_TC_CASE(A, B, E) \
  if (a_mn == A && b_mn == B && epi == E) return launch_t<A, B, E>(t, gs, ep, st, kernel_id);
  PQN_TC_CASE(0, 1, EPI_STORE)
  PQN_TC_CASE(0, 1, EPI_LN_TRAIN)
  PQN_TC_CASE(0, 1, EPI_LN_HEAD)
  PQN_TC_CASE(1, 1, EPI_STORE)
  PQN_TC_CASE(0, 0, EPI_STORE)
  PQN_TC_CASE(0, 0, EPI_RELU_MASK)
  PQN_TC_CASE(0, 0, EPI_RELU_BITS)
#undef PQN_TC_CASE
  return set_error(PQN_E_UNSUPPORTED, "tc_gemm: combination a_mn=%d b_mn=%d epi=%d not instantiated", a_mn, b_mn, epi);
}

int launch_gemm16(int a_mn, int b_mn, int epi, const CUtensorMap* t, const GemmShape& gs, const EpiParams& ep,
                  cudaStream_t st, int kernel_id) {
#define PQN_TC_CASE(A, B, E) \
  if (a_mn == A && b_mn == B && epi == E) return launch_t<A, B, E, true>(t, gs, ep, st, kernel_id);
  PQN_TC_CASE(0, 1, EPI_STORE)
  PQN_TC_CASE(0, 1, EPI_LN_TRAIN)
  PQN_TC_CASE(0, 1, EPI_LN_HEAD)
  PQN_TC_CASE(1, 1, EPI_STORE)
  PQN_TC_CASE(0, 0, EPI_STORE)
  PQN_TC_CASE(0, 0, EPI_RELU_MASK)
  PQN_TC_CASE(0, 0, EPI_RELU_BITS)
#undef PQN_TC_CASE
  return set_error(PQN_E_UNSUPPORTED, "tc_gemm16: combination a_mn=%d b_mn=%d epi=%d not instantiated", a_mn, b_mn, epi);
}

int make_tmap16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t mid, uint64_t seeds, uint64_t mid_stride_elems,
                uint64_t seed_stride_elems, uint32_t box_mid) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(PQN_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {inner, mid, seeds};
  cuuint64_t strides[2] = {mid_stride_elems * 2, seed_stride_elems * 2};
  cuuint32_t box[3] = {64, box_mid, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PQN_E_CUDA, "cuTensorMapEncodeTiled(fp16) failed (%d)", (int)r);
  return PQN_OK;
}

__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 o;
  o.x = tf32_lo(v.x); o.y = tf32_lo(v.y); o.z = tf32_lo(v.z); o.w = tf32_lo(v.w);
  reinterpret_cast<float4*>(lo)[i] = o;
}

// fp16 split planes of an fp32 tensor (optionally pre-scaled by a power of two): hi = fp16(x*scale),
// lo = fp16((x*scale - hi) * 2^11)
__global__ void split16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n4,
                               float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  __half2 h0, h1, l0, l1;
  split16x2(v.x * scale, v.y * scale, h0, l0);
  split16x2(v.z * scale, v.w * scale, h1, l1);
  reinterpret_cast<uint2*>(hi)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  reinterpret_cast<uint2*>(lo)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
}

// Debug kernel: one CTA, one 128x128x32 tile, no pipelining.  Dumps the smem tiles TMA produced and the TMEM
// accumulator after `nk` k-steps so that descriptor/layout problems can be diagnosed from the host.
template <int A_MN, int B_MN>
__global__ void __launch_bounds__(128, 1)
    tc_debug_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    float* __restrict__ dump_a, float* __restrict__ dump_b, float* __restrict__ out_d, int nk,
                    uint32_t* __restrict__ info) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + 2 * TC_TILE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    info[0] = tmem_base;
    info[1] = smem_base;
    mbar_expect_tx(&bars[0], 2 * TC_TILE_BYTES);
    if (A_MN) {
      for (int j = 0; j < 4; ++j) tma_load_3d(smem_base + j * 4096, &tm_a, &bars[0], 32 * j, 0, 0);
    } else {
      tma_load_3d(smem_base, &tm_a, &bars[0], 0, 0, 0);
    }
    if (B_MN) {
      for (int j = 0; j < 4; ++j) tma_load_3d(smem_base + TC_TILE_BYTES + j * 4096, &tm_b, &bars[0], 32 * j, 0, 0);
    } else {
      tma_load_3d(smem_base + TC_TILE_BYTES, &tm_b, &bars[0], 0, 0, 0);
    }
  }
  mbar_wait(&bars[0], 0);
  const float* sa = reinterpret_cast<const float*>(smem_al);
  const float* sb = reinterpret_cast<const float*>(smem_al + TC_TILE_BYTES);
  for (int i = threadIdx.x; i < 4096; i += 128) { dump_a[i] = sa[i]; dump_b[i] = sb[i]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, 128, A_MN, B_MN);
    info[2] = idesc;
    for (int ks = 0; ks < nk; ++ks) {
      const uint64_t da = make_sdesc<A_MN>(smem_base, ks);
      const uint64_t db = make_sdesc<B_MN>(smem_base + TC_TILE_BYTES, ks);
      if (ks == 0) { info[4] = (uint32_t)da; info[5] = (uint32_t)(da >> 32); info[6] = (uint32_t)db; info[7] = (uint32_t)(db >> 32); }
      umma_tf32(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tcgen05_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out_d[row * 128 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace tc
}  // namespace pqn

using namespace pqn;
using namespace pqn::tc;

extern "C" {

int pqn_tc_split_lo(const float* x, float* lo, int64_t n, void* stream) {
  if (!x || !lo || n < 0 || (n & 3)) return set_error(PQN_E_INVALID, "pqn_tc_split_lo: bad argument (n %% 4 == 0)");
  if (n == 0) return PQN_OK;
  {
    LaunchScope _ls(K_TC_SPLIT, (cudaStream_t)stream);
    split_lo_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, lo, n / 4);
  }
  return check_launch("pqn_tc_split_lo");
}

int pqn_tc_split16(const float* x, void* hi, void* lo, int64_t n, float scale, void* stream) {
  if (!x || !hi || !lo || n < 0 || (n & 3)) return set_error(PQN_E_INVALID, "pqn_tc_split16: bad argument (n %% 4 == 0)");
  if (n == 0) return PQN_OK;
  {
    LaunchScope _ls(K_TC_SPLIT, (cudaStream_t)stream);
    split16_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)hi, (__half*)lo, n / 4,
                                                                                       scale);
  }
  return check_launch("pqn_tc_split16");
}

// Test hook: D[s] = A[s] . B[s] * out_scale on the fp16-split tcgen05 path; operands are the (hi, lo') planes written
// by pqn_tc_split16.  Layout flags as in pqn_tc_gemm_test; N % 128 == 0.
int pqn_tc_gemm16_test(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, float* d, int32_t S,
                       int32_t M, int32_t N, int32_t K, int a_mn, int b_mn, float out_scale, void* stream) {
  if (!a_hi || !a_lo || !b_hi || !b_lo || !d || S <= 0 || M <= 0 || N <= 0 || K <= 0 || (N % 128) || (K % 8) || (M % 8))
    return set_error(PQN_E_INVALID, "pqn_tc_gemm16_test: bad argument");
  CUtensorMap t[4];
  int rc;
  const void* ap[2] = {a_hi, a_lo};
  const void* bp[2] = {b_hi, b_lo};
  for (int i = 0; i < 2; ++i) {
    if (a_mn) { if ((rc = make_tmap16(&t[i], ap[i], M, K, S, M, (uint64_t)M * K, 64))) return rc; }
    else { if ((rc = make_tmap16(&t[i], ap[i], K, M, S, K, (uint64_t)M * K, 128))) return rc; }
    if (b_mn) { if ((rc = make_tmap16(&t[2 + i], bp[i], N, K, S, N, (uint64_t)N * K, 64))) return rc; }
    else { if ((rc = make_tmap16(&t[2 + i], bp[i], K, N, S, K, (uint64_t)N * K, 128))) return rc; }
  }
  GemmShape gs = {};
  gs.S = S; gs.M = M; gs.m_tiles = (M + 127) / 128; gs.n_tiles = N / 128; gs.k_blocks = (K + TC_BK16 - 1) / TC_BK16;
  gs.split3 = 1;
  EpiParams ep = {};
  ep.out = d; ep.ld_out = N; ep.out_seed_stride = (int64_t)M * N; ep.out_scale = out_scale;
  return launch_gemm16(a_mn, b_mn, EPI_STORE, t, gs, ep, (cudaStream_t)stream);
}

// Debug hook (tests only): single 128x128x32 tile; a: [128][32] (a_mn=0) or [32][128] (a_mn=1); same for b.
int pqn_tc_debug(const float* a, const float* b, float* dump_a, float* dump_b, float* out_d, uint32_t* info, int a_mn,
                 int b_mn, int nk, void* stream) {
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) { if ((rc = make_tmap(&ta, a, 128, 32, 1, 128, 4096, 32, 1))) return rc; }
  else { if ((rc = make_tmap(&ta, a, 32, 128, 1, 32, 4096, 128, 0))) return rc; }
  if (b_mn) { if ((rc = make_tmap(&tb, b, 128, 32, 1, 128, 4096, 32, 1))) return rc; }
  else { if ((rc = make_tmap(&tb, b, 32, 128, 1, 32, 4096, 128, 0))) return rc; }
  const int smem = 2 * TC_TILE_BYTES + 1024 + 64;
  cudaStream_t st = (cudaStream_t)stream;
#define PQN_DBG(A, B)                                                                                        \
  if (a_mn == A && b_mn == B) {                                                                              \
    cudaFuncSetAttribute(tc_debug_kernel<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    tc_debug_kernel<A, B><<<1, 128, smem, st>>>(ta, tb, dump_a, dump_b, out_d, nk, info);                    \
  }
  PQN_DBG(0, 0) PQN_DBG(0, 1) PQN_DBG(1, 0) PQN_DBG(1, 1)
#undef PQN_DBG
  return check_launch("pqn_tc_debug");
}

// Test hook: D[s] = A[s] . B[s] on the tcgen05 path (fp32 in, fp32 out).
//   a_mn = 0: A is [S][M][K] (K contiguous)   a_mn = 1: A is [S][K][M] (M contiguous)
//   b_mn = 0: B is [S][N][K] (K contiguous)   b_mn = 1: B is [S][K][N] (N contiguous)
//   split3 = 1: 3xTF32 (a_lo / b_lo must hold x - trunc_tf32(x)); 2: same, but A_lo is derived in the kernel
//   (a_lo unused); 0: single-pass TF32.
// M, N multiples of 128 are not required for M (rows are guarded); N % 128 == 0, K % 32 == 0 or zero-filled.
int pqn_tc_gemm_test(const float* a, const float* a_lo, const float* b, const float* b_lo, float* d, int32_t S,
                     int32_t M, int32_t N, int32_t K, int a_mn, int b_mn, int split3, void* stream) {
  if (!a || !b || !d || S <= 0 || M <= 0 || N <= 0 || K <= 0 || (N % 128) || (split3 && !b_lo) || (split3 == 1 && !a_lo))
    return set_error(PQN_E_INVALID, "pqn_tc_gemm_test: bad argument");
  CUtensorMap t[4];
  int rc;
  const float* al = split3 == 1 ? a_lo : a;
  const float* bl = split3 ? b_lo : b;
  if (a_mn) {
    if ((rc = make_tmap(&t[0], a, M, K, S, M, (uint64_t)M * K, 32, 1))) return rc;
    if ((rc = make_tmap(&t[1], al, M, K, S, M, (uint64_t)M * K, 32, 1))) return rc;
  } else {
    if ((rc = make_tmap(&t[0], a, K, M, S, K, (uint64_t)M * K, 128, 0))) return rc;
    if ((rc = make_tmap(&t[1], al, K, M, S, K, (uint64_t)M * K, 128, 0))) return rc;
  }
  if (b_mn) {
    if ((rc = make_tmap(&t[2], b, N, K, S, N, (uint64_t)N * K, 32, 1))) return rc;
    if ((rc = make_tmap(&t[3], bl, N, K, S, N, (uint64_t)N * K, 32, 1))) return rc;
  } else {
    if ((rc = make_tmap(&t[2], b, K, N, S, K, (uint64_t)N * K, 128, 0))) return rc;
    if ((rc = make_tmap(&t[3], bl, K, N, S, K, (uint64_t)N * K, 128, 0))) return rc;
  }
  GemmShape gs = {};
  gs.S = S; gs.M = M; gs.m_tiles = (M + 127) / 128; gs.n_tiles = N / 128; gs.k_blocks = (K + TC_BK - 1) / TC_BK;
  gs.split3 = split3 ? 1 : 0;
  gs.a_lo_inline = split3 == 2;
  EpiParams ep = {};
  ep.out = d; ep.ld_out = N; ep.out_seed_stride = (int64_t)M * N;
  return launch_gemm(a_mn, b_mn, EPI_STORE, t, gs, ep, (cudaStream_t)stream);
}

}  // extern "C"
