// jax.random (threefry2x32) semantics, as consumed by the PQN hot path.
//
// Replaces, on device, the PRNG arithmetic the reference reaches through
// jax.random.split / uniform / randint / choice at
//   purejaxql/pqn_minatar.py:107-112 (per-env key split), :116-125 (eps-greedy),
//   :183 (3-way split of the scan carry), and inside gymnax Environment.step.
// Both counter layouts are supported (`part` = jax_threefry_partitionable):
//   part=0  "original" layout, default for the reference's pinned jax<=0.4.38
//   part=1  "partitionable" layout, default from jax 0.5
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PQN_HD __host__ __device__ __forceinline__
#else
#define PQN_HD inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif
#endif

namespace pqn {

struct Key {
  uint32_t k0, k1;
};

PQN_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// Threefry-2x32, 20 rounds (Random123).
PQN_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define PQN_TF_R(r) \
  x0 += x1;         \
  x1 = rotl32(x1, r) ^ x0;
  PQN_TF_R(13) PQN_TF_R(15) PQN_TF_R(26) PQN_TF_R(6)
  x0 += k1; x1 += k2 + 1u;
  PQN_TF_R(17) PQN_TF_R(29) PQN_TF_R(16) PQN_TF_R(24)
  x0 += k2; x1 += k0 + 2u;
  PQN_TF_R(13) PQN_TF_R(15) PQN_TF_R(26) PQN_TF_R(6)
  x0 += k0; x1 += k1 + 3u;
  PQN_TF_R(17) PQN_TF_R(29) PQN_TF_R(16) PQN_TF_R(24)
  x0 += k1; x1 += k2 + 4u;
  PQN_TF_R(13) PQN_TF_R(15) PQN_TF_R(26) PQN_TF_R(6)
  x0 += k2; x1 += k0 + 5u;
#undef PQN_TF_R
}

// Word j (0 <= j < 2*num) of the flat output of the original-layout
// threefry_2x32(key, iota(2*num)): blocks are (c, c+num), outputs concat(y0, y1).
PQN_HD uint32_t split_word_original(Key k, uint32_t num, uint32_t j) {
  uint32_t c = (j < num) ? j : j - num;
  uint32_t x0 = c, x1 = c + num;
  threefry2x32(k.k0, k.k1, x0, x1);
  return (j < num) ? x0 : x1;
}

// jax.random.split(key, num)[i]
PQN_HD Key split_at(Key k, uint32_t num, uint32_t i, int part) {
  Key out;
  if (part) {
    uint32_t x0 = 0u, x1 = i;
    threefry2x32(k.k0, k.k1, x0, x1);
    out.k0 = x0;
    out.k1 = x1;
  } else {
    out.k0 = split_word_original(k, num, 2u * i);
    out.k1 = split_word_original(k, num, 2u * i + 1u);
  }
  return out;
}

// jax.random.split(key) -> both children (num = 2); 2 blocks in the original
// layout ((0,2) -> a0,b0 ; (1,3) -> a1,b1 ; child0 = (a0,a1), child1 = (b0,b1)).
PQN_HD void split2(Key k, int part, Key& c0, Key& c1) {
  if (part) {
    uint32_t a0 = 0u, a1 = 0u, b0 = 0u, b1 = 1u;
    threefry2x32(k.k0, k.k1, a0, a1);
    threefry2x32(k.k0, k.k1, b0, b1);
    c0.k0 = a0; c0.k1 = a1; c1.k0 = b0; c1.k1 = b1;
  } else {
    uint32_t a0 = 0u, b0 = 2u, a1 = 1u, b1 = 3u;
    threefry2x32(k.k0, k.k1, a0, b0);
    threefry2x32(k.k0, k.k1, a1, b1);
    c0.k0 = a0; c0.k1 = a1; c1.k0 = b0; c1.k1 = b1;
  }
}

// jax.random.split(key, 3): original layout blocks (0,3),(1,4),(2,5) ->
// out = [a0,a1,a2,b0,b1,b2] -> children (a0,a1),(a2,b0),(b1,b2).
PQN_HD void split3(Key k, int part, Key& c0, Key& c1, Key& c2) {
  if (part) {
    c0 = split_at(k, 3, 0, 1); c1 = split_at(k, 3, 1, 1); c2 = split_at(k, 3, 2, 1);
  } else {
    uint32_t a0 = 0u, b0 = 3u, a1 = 1u, b1 = 4u, a2 = 2u, b2 = 5u;
    threefry2x32(k.k0, k.k1, a0, b0);
    threefry2x32(k.k0, k.k1, a1, b1);
    threefry2x32(k.k0, k.k1, a2, b2);
    c0.k0 = a0; c0.k1 = a1; c1.k0 = a2; c1.k1 = b0; c2.k0 = b1; c2.k1 = b2;
  }
}

// random_bits(key, 32, shape)[i] for a shape with n elements.
PQN_HD uint32_t bits_at(Key k, uint32_t n, uint32_t i, int part) {
  if (part) {
    uint32_t x0 = 0u, x1 = i;
    threefry2x32(k.k0, k.k1, x0, x1);
    return x0 ^ x1;
  }
  const uint32_t half = (n + 1u) >> 1;
  const bool lo = i < half;
  const uint32_t c = lo ? i : i - half;
  uint32_t x0 = c, x1 = c + half;
  if ((n & 1u) && c == half - 1u) x1 = 0u;  // the odd-size pad counter is a literal 0
  threefry2x32(k.k0, k.k1, x0, x1);
  return lo ? x0 : x1;
}

// random_bits(key, 32, ()) : scalar draw.
PQN_HD uint32_t bits_scalar(Key k, int part) { return bits_at(k, 1u, 0u, part); }

PQN_HD float bits_to_unit_float(uint32_t bits) {
  const uint32_t fb = (bits >> 9) | 0x3F800000u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(fb) - 1.0f;
#else
  union { uint32_t u; float f; } cv;
  cv.u = fb;
  return cv.f - 1.0f;
#endif
}

// jax.random.uniform(key, (), f32, minval, maxval)
PQN_HD float uniform_from_bits(uint32_t bits, float minval, float maxval) {
  float f = bits_to_unit_float(bits) * (maxval - minval) + minval;
  return f < minval ? minval : f;
}
PQN_HD float uniform_scalar(Key k, int part) { return uniform_from_bits(bits_scalar(k, part), 0.0f, 1.0f); }

// jax.random.randint(key, (), 0, span) with 1 <= span < 2^16 (small action sets):
//   k1,k2 = split(key); off = ((hi % span) * mult + lo % span) % span,
//   mult = ((2^16 % span)^2) % span, all in uint32.
PQN_HD int32_t randint_scalar(Key k, uint32_t span, int part) {
  Key k1, k2;
  split2(k, part, k1, k2);
  const uint32_t hi = bits_scalar(k1, part);
  const uint32_t lo = bits_scalar(k2, part);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  const uint32_t off = ((hi % span) * mult + (lo % span)) % span;
  return (int32_t)off;
}

// randint(key, (n,), lo, hi)[i] — vector draw element (used by Freeway etc.)
PQN_HD int32_t randint_at(Key k, uint32_t n, uint32_t i, int32_t minval, int32_t maxval, int part) {
  Key k1, k2;
  split2(k, part, k1, k2);
  const uint32_t hi = bits_at(k1, n, i, part);
  const uint32_t lo = bits_at(k2, n, i, part);
  uint32_t span = (uint32_t)(maxval - minval);
  if (maxval <= minval) span = 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  const uint32_t off = ((hi % span) * mult + (lo % span)) % span;
  return minval + (int32_t)off;
}

}  // namespace pqn
