// jax.random.permutation(key, n) index shuffle, batched over S independent keys -- the minibatch permutation of
// purejaxql/pqn_minatar.py:299-315 (`jax.random.permutation(rng, x)` on every leaf with the same key).
//
// jax's `_shuffle` (third party, restated): ceil(3 ln n / ln(2^32 - 1)) rounds of
//     key, sub = split(key);  sort_keys = random_bits(sub, 32, (n,));  x = stable_sort_by_key(sort_keys, x)
// starting from x = arange(n).  A stable sort by a 32-bit key is the sort by the 64-bit composite (key << 32 | current
// position), whose values are unique -- so any exact sort of the composites reproduces it.  The sort keys are uniform
// random bits, which makes a one-level bucket sort exact AND balanced:
//   A  histogram of the top B key bits per seed (integer atomics), B chosen so that a bucket holds ~64 elements
//   B  exclusive scan of the histogram per seed (bucket offsets; also primes the scatter cursors)
//   C  scatter the composites into their bucket's range (order inside the range is arbitrary)
//   D  one warp per bucket ranks every composite among the bucket's composites (count of smaller ones; staged in
//      shared memory, a fixed-size fallback reads global memory for an improbable oversized bucket) and writes the
//      payload -- the element's index in round 0, the previous round's value afterwards -- to offset + rank.
// The result does not depend on the scatter order, so it is deterministic and bit-identical to the stable sort.
// Replaces the torch.sort (cub segmented radix sort) + gather plumbing of rounds 1-2.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pqn_b200.h"
#include "api_common.h"
#include "threefry.cuh"

namespace pqn {

static int g_perm_bucket_log2 = 6;       // target elements per bucket = 64 (pqn_set_permutation_bucket_log2: test hook)
constexpr int PERM_STAGE = 256;          // composites a warp can rank from shared memory
constexpr int PERM_WARPS = 8;            // buckets per CTA of the rank kernel

struct PermPlan {
  int rounds, bucket_bits;
  int64_t nb;
};
static PermPlan perm_plan(int64_t n) {
  PermPlan p;
  p.rounds = (int)ceil(3.0 * log((double)(n > 1 ? n : 1)) / log(4294967295.0));
  int b = 0;
  while (b < 24 && (n >> b) > ((int64_t)1 << g_perm_bucket_log2)) ++b;
  p.bucket_bits = b;
  p.nb = (int64_t)1 << b;
  return p;
}

struct PermWs {
  uint64_t* comp;   // [S][n]
  int32_t* ping;    // [S][n]
  int32_t *hist, *offs, *cursor;  // [S][nb]
};
static int64_t perm_carve(int64_t n, int32_t S, char* base, PermWs* w) {
  const PermPlan p = perm_plan(n);
  int64_t off = 0;
  auto take = [&](int64_t bytes) -> char* {
    char* q = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return q;
  };
  PermWs tmp;
  PermWs* ww = w ? w : &tmp;
  ww->comp = reinterpret_cast<uint64_t*>(take((int64_t)S * n * 8));
  ww->ping = reinterpret_cast<int32_t*>(take((int64_t)S * n * 4));
  ww->hist = reinterpret_cast<int32_t*>(take((int64_t)S * p.nb * 4));
  ww->offs = reinterpret_cast<int32_t*>(take((int64_t)S * p.nb * 4));
  ww->cursor = reinterpret_cast<int32_t*>(take((int64_t)S * p.nb * 4));
  return off;
}

// sort-key generator of round r for seed `seed`: key_{j+1}, sub_j = split(key_j); returns sub_r
__device__ __forceinline__ Key perm_round_key(const uint32_t* __restrict__ keys, int seed, int round, int part) {
  Key k{keys[2 * seed], keys[2 * seed + 1]}, sub = k;
  for (int j = 0; j <= round; ++j) {
    Key c0, c1;
    split2(k, part, c0, c1);
    k = c0;
    sub = c1;
  }
  return sub;
}

constexpr int PERM_ITEMS = 4;  // elements per thread of the histogram / scatter kernels

// A: hist[seed][bits >> (32 - B)] += 1          C: comp[seed][cursor[seed][bucket]++] = bits << 32 | i
template <bool SCATTER>
__global__ void __launch_bounds__(256) perm_bucket_kernel(const uint32_t* __restrict__ keys, int64_t n, int round, int part,
                                                          int bucket_bits, int64_t nb, int32_t* __restrict__ hist_or_cursor,
                                                          uint64_t* __restrict__ comp) {
  __shared__ Key sub_s;
  const int seed = blockIdx.y;
  if (threadIdx.x == 0) sub_s = perm_round_key(keys, seed, round, part);
  __syncthreads();
  const Key sub = sub_s;
  const int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * PERM_ITEMS;
#pragma unroll
  for (int it = 0; it < PERM_ITEMS; ++it) {
    const int64_t i = base + it;
    if (i >= n) break;
    const uint32_t bits = bits_at(sub, (uint32_t)n, (uint32_t)i, part);
    const uint32_t b = bucket_bits ? bits >> (32 - bucket_bits) : 0u;
    int32_t* slot = hist_or_cursor + (int64_t)seed * nb + b;
    if (SCATTER) {
      const int32_t pos = atomicAdd(slot, 1);
      comp[(int64_t)seed * n + pos] = ((uint64_t)bits << 32) | (uint64_t)(uint32_t)i;
    } else {
      atomicAdd(slot, 1);
    }
  }
}

// B: offs[seed][b] = cursor[seed][b] = sum_{b' < b} hist[seed][b'];  hist is cleared for the next round.  grid = S
__global__ void __launch_bounds__(1024) perm_scan_kernel(int32_t* __restrict__ hist, int32_t* __restrict__ offs,
                                                         int32_t* __restrict__ cursor, int64_t nb) {
  __shared__ int32_t part_sum[1024];
  const int seed = blockIdx.x, tid = threadIdx.x;
  int32_t* __restrict__ h = hist + (int64_t)seed * nb;
  const int64_t per = (nb + 1023) / 1024;
  const int64_t lo = (int64_t)tid * per, hi = lo + per < nb ? lo + per : nb;
  int32_t s = 0;
  for (int64_t b = lo; b < hi; ++b) s += h[b];
  part_sum[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan of the 1024 chunk sums
    const int32_t v = tid >= o ? part_sum[tid - o] : 0;
    __syncthreads();
    part_sum[tid] += v;
    __syncthreads();
  }
  int32_t run = part_sum[tid] - s;       // exclusive prefix of this thread's chunk
  for (int64_t b = lo; b < hi; ++b) {
    const int32_t c = h[b];
    offs[(int64_t)seed * nb + b] = run;
    cursor[(int64_t)seed * nb + b] = run;
    h[b] = 0;
    run += c;
  }
}

// D: one warp per bucket.  out position p of seed s lives at out[s * n + p] (chunk == 0) or, for the last round of a
// minibatch layout, at out[(p / chunk) * S * chunk + s * chunk + p % chunk]  ([n / chunk][S][chunk]).
__global__ void __launch_bounds__(PERM_WARPS * 32) perm_rank_kernel(const uint64_t* __restrict__ comp, const int32_t* __restrict__ offs,
                                                                    int64_t n, int64_t nb, const int32_t* __restrict__ prev,
                                                                    int32_t* __restrict__ out, int64_t chunk) {
  __shared__ uint64_t stage[PERM_WARPS][PERM_STAGE];
  const int seed = blockIdx.y, S = gridDim.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * PERM_WARPS + warp;
  if (b >= nb) return;
  const int32_t* __restrict__ o = offs + (int64_t)seed * nb;
  const int64_t lo = o[b], hi = b + 1 < nb ? (int64_t)o[b + 1] : n;
  const int m = (int)(hi - lo);
  if (m <= 0) return;
  const uint64_t* __restrict__ src = comp + (int64_t)seed * n + lo;
  const bool staged = m <= PERM_STAGE;
  if (staged) {
    for (int j = lane; j < m; j += 32) stage[warp][j] = src[j];
    __syncwarp();
  }
  for (int j = lane; j < m; j += 32) {
    const uint64_t c = staged ? stage[warp][j] : src[j];
    int rank = 0;
    if (staged) {
      for (int k = 0; k < m; ++k) rank += stage[warp][k] < c ? 1 : 0;
    } else {
      for (int k = 0; k < m; ++k) rank += src[k] < c ? 1 : 0;
    }
    const uint32_t idx = (uint32_t)c;                      // position of the element before this round
    const int32_t val = prev ? prev[(int64_t)seed * n + idx] : (int32_t)idx;
    const int64_t p = lo + rank;
    const int64_t dst = chunk > 0 ? (p / chunk) * ((int64_t)S * chunk) + (int64_t)seed * chunk + p % chunk
                                  : (int64_t)seed * n + p;
    out[dst] = val;
  }
}

__global__ void perm_identity_kernel(int32_t* __restrict__ out, int64_t n, int64_t chunk) {
  const int seed = blockIdx.y, S = gridDim.y;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t dst = chunk > 0 ? (p / chunk) * ((int64_t)S * chunk) + (int64_t)seed * chunk + p % chunk : (int64_t)seed * n + p;
  out[dst] = (int32_t)p;
}

}  // namespace pqn

using namespace pqn;

extern "C" {

int pqn_set_permutation_bucket_log2(int log2_elems) {
  const int prev = g_perm_bucket_log2;
  if (log2_elems >= 0 && log2_elems <= 20) g_perm_bucket_log2 = log2_elems;
  return prev;
}

int64_t pqn_permutation_workspace_bytes(int64_t n, int32_t S) {
  if (n <= 0 || S <= 0) return 0;
  return perm_carve(n, S, nullptr, nullptr);
}

int pqn_permutation(const uint32_t* keys, int64_t n, int32_t S, int rng_mode, int32_t* out, int64_t out_chunk,
                    void* workspace, void* stream) {
  if (!keys || !out || n <= 0 || S <= 0 || n > 0x7fffffffLL || S > 65535)
    return set_error(PQN_E_INVALID, "pqn_permutation: bad argument");
  if (out_chunk < 0 || (out_chunk > 0 && n % out_chunk != 0))
    return set_error(PQN_E_INVALID, "pqn_permutation: out_chunk must divide n");
  const PermPlan pl = perm_plan(n);
  cudaStream_t st = (cudaStream_t)stream;
  if (pl.rounds == 0) {
    LaunchScope _ls(K_PERM, st);
    perm_identity_kernel<<<dim3((unsigned)((n + 255) / 256), S), 256, 0, st>>>(out, n, out_chunk);
    return check_launch("pqn_permutation");
  }
  if (!workspace) return set_error(PQN_E_INVALID, "pqn_permutation: workspace required");
  PermWs w;
  perm_carve(n, S, (char*)workspace, &w);
  if (cudaMemsetAsync(w.hist, 0, (size_t)S * pl.nb * 4, st) != cudaSuccess) return check_launch("pqn_permutation(memset)");
  const dim3 eg((unsigned)((n + 256 * PERM_ITEMS - 1) / (256 * PERM_ITEMS)), S);
  const dim3 rg((unsigned)((pl.nb + PERM_WARPS - 1) / PERM_WARPS), S);
  // the last round writes `out`; earlier rounds alternate between the two buffers so that no round reads its own output
  int32_t* bufs[2] = {out, w.ping};
  const int32_t* prev = nullptr;
  for (int r = 0; r < pl.rounds; ++r) {
    const bool last = r == pl.rounds - 1;
    int32_t* dst = bufs[(pl.rounds - 1 - r) & 1];
    { LaunchScope _ls(K_PERM, st);
      perm_bucket_kernel<false><<<eg, 256, 0, st>>>(keys, n, r, rng_mode, pl.bucket_bits, pl.nb, w.hist, nullptr); }
    { LaunchScope _ls(K_PERM, st); perm_scan_kernel<<<S, 1024, 0, st>>>(w.hist, w.offs, w.cursor, pl.nb); }
    { LaunchScope _ls(K_PERM, st);
      perm_bucket_kernel<true><<<eg, 256, 0, st>>>(keys, n, r, rng_mode, pl.bucket_bits, pl.nb, w.cursor, w.comp); }
    { LaunchScope _ls(K_PERM, st);
      perm_rank_kernel<<<rg, PERM_WARPS * 32, 0, st>>>(w.comp, w.offs, n, pl.nb, prev, dst, last ? out_chunk : 0); }
    prev = dst;
  }
  return check_launch("pqn_permutation");
}

}  // extern "C"
