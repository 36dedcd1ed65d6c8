// Q-network forward / loss-gradient kernels for sm_100a, batched over S
// independent seeds (blockIdx.y or .z = seed; every seed has its own weights).
//
// Reference: QNetwork/CNN  purejaxql/pqn_minatar.py:24-69,
//            MLP QNetwork  purejaxql/pqn_gymnax.py:29-58,
//            _loss_fn      purejaxql/pqn_minatar.py:271-291.
// fp32 throughout (the north star asks for 1e-5 agreement with an fp32 oracle):
// the dense contractions are register-tiled FFMA GEMMs (128x128x16 CTA tiles,
// 8x8 per thread, double-buffered shared memory, 128-bit LDS) with the
// bias + LayerNorm + ReLU (+ Q-head) epilogue fused in registers; the 3x3 conv
// consumes the bit-packed observations directly (its input is {0,1}/255, so a
// tap contributes either w/255 or nothing) and skips taps no lane of the warp
// has set.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pqn_b200.h"
#include "api_common.h"
#include "tc_common.cuh"

namespace pqn {

constexpr int GT = 256;   // threads per GEMM CTA (16 x 16)
constexpr int BK = 16;    // reduce-dim tile
constexpr float LN_EPS = 1e-6f;
constexpr int CONV_O = 16;   // conv output channels
constexpr int CONV_PIX = 64; // 8x8 output pixels
constexpr int HID_CNN = 128;
constexpr int FLAT_CNN = CONV_PIX * CONV_O;  // 1024

// tcgen05 path for the CNN dense layer (pqn_set_tensor_core_path); default on
static int g_use_tc = 2;  // 0 FFMA, 1 tcgen05 3xTF32 (A_lo derived in the kernel), 2 tcgen05 fp16-split planes (default)
// warp-level tensor-core (mma.sync tf32) conv kernels (pqn_set_conv_mma_path); default on
static int g_conv_mma = 1;   // 0: fp32 CUDA cores, 1: fp16 mma.sync forward (default), 2: tcgen05 forward, 3: tf32 mma.sync forward
                             // (1-3: tf32 mma.sync backward)

static inline int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

static int make_layout(const pqn_net_desc_t* d, pqn_net_layout_t* L) {
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += align4(n); return o; };
  L->d1_w = L->d1_b = L->ln1_scale = L->ln1_bias = L->conv_w = L->conv_b = -1;
  L->gru_ir_w = L->gru_ir_b = L->gru_iz_w = L->gru_iz_b = L->gru_in_w = L->gru_in_b = -1;
  L->gru_hr_w = L->gru_hz_w = L->gru_hn_w = L->gru_hn_b = -1;
  const int A = d->num_actions;
  const bool has_norm = d->norm_type != PQN_NORM_NONE;   // layer_norm and batch_norm have (scale, bias) of the same shape
  if (d->kind == PQN_NET_MINATAR_CNN) {
    const int C = d->in_c;
    L->bn_scale = take(C); L->bn_bias = take(C);
    L->conv_w = take(9 * C * CONV_O); L->conv_b = take(CONV_O);
    L->ln0_scale = L->ln0_bias = -1;   // norm_type none: the network has no normalisation parameters
    if (has_norm) { L->ln0_scale = take(CONV_O); L->ln0_bias = take(CONV_O); }
    L->d0_w = take((int64_t)FLAT_CNN * HID_CNN); L->d0_b = take(HID_CNN);
    if (has_norm) { L->ln1_scale = take(HID_CNN); L->ln1_bias = take(HID_CNN); }
    L->head_w = take((int64_t)HID_CNN * A); L->head_b = take(A);
  } else if (d->kind == PQN_NET_MLP || d->kind == PQN_NET_RNN) {
    const int D = d->in_c, H = d->hidden;
    L->bn_scale = take(D); L->bn_bias = take(D);
    L->d0_w = take((int64_t)D * H); L->d0_b = take(H);
    L->ln0_scale = L->ln0_bias = -1;
    if (has_norm) { L->ln0_scale = take(H); L->ln0_bias = take(H); }
    if (d->layers == 2) {
      L->d1_w = take((int64_t)H * H); L->d1_b = take(H);
      if (has_norm) { L->ln1_scale = take(H); L->ln1_bias = take(H); }
    }
    if (d->kind == PQN_NET_RNN) {   // flax GRUCell: input denses with bias, recurrent ones without (except hn)
      L->gru_ir_w = take((int64_t)(H + A) * H); L->gru_ir_b = take(H);
      L->gru_iz_w = take((int64_t)(H + A) * H); L->gru_iz_b = take(H);
      L->gru_in_w = take((int64_t)(H + A) * H); L->gru_in_b = take(H);
      L->gru_hr_w = take((int64_t)H * H); L->gru_hz_w = take((int64_t)H * H);
      L->gru_hn_w = take((int64_t)H * H); L->gru_hn_b = take(H);
    }
    L->head_w = take((int64_t)H * A); L->head_b = take(A);
  } else {
    return -1;
  }
  L->total = off;
  return 0;
}

static int check_desc(const pqn_net_desc_t* d, const char* who) {
  if (!d) return set_error(PQN_E_INVALID, "%s: desc is NULL", who);
  if (d->num_actions < 1 || d->num_actions > 32) return set_error(PQN_E_INVALID, "%s: num_actions=%d out of [1,32]", who, d->num_actions);
  if (d->norm_type < 0 || d->norm_type > 2 || d->norm_input < 0 || d->norm_input > 1)
    return set_error(PQN_E_INVALID, "%s: norm_type=%d norm_input=%d", who, d->norm_type, d->norm_input);
  {
    // row_bwd_kernel keeps [A][N] head weights + eight warp-private [A][N] gradient slices in shared memory
    const int Nh = d->kind == PQN_NET_MINATAR_CNN ? 128 : d->hidden;
    const size_t need = (size_t)(24 * Nh + 9 * d->num_actions * Nh + 8 * d->num_actions + 16) * sizeof(float);
    if (need > 227u * 1024u)
      return set_error(PQN_E_UNSUPPORTED, "%s: hidden=%d with num_actions=%d needs %zu B of shared memory for the head "
                       "backward (limit 227 KB)", who, Nh, d->num_actions, need);
  }
  if (d->kind == PQN_NET_MINATAR_CNN) {
    if (d->in_c != 4 && d->in_c != 6 && d->in_c != 7 && d->in_c != 10)
      return set_error(PQN_E_UNSUPPORTED, "%s: CNN in_c=%d (MinAtar uses 4/6/7/10)", who, d->in_c);
    return PQN_OK;
  }
  if (d->kind == PQN_NET_MLP || d->kind == PQN_NET_RNN) {
    if (d->hidden != 128 && d->hidden != 256) return set_error(PQN_E_UNSUPPORTED, "%s: MLP hidden=%d (128 or 256 built)", who, d->hidden);
    if (d->layers != 1 && d->layers != 2) return set_error(PQN_E_UNSUPPORTED, "%s: MLP layers=%d (1 or 2 built)", who, d->layers);
    if (d->in_c < 1 || d->in_c > 4096) return set_error(PQN_E_INVALID, "%s: MLP in dim %d", who, d->in_c);
    return PQN_OK;
  }
  return set_error(PQN_E_INVALID, "%s: unknown net kind %d", who, d->kind);
}

// ---------------------------------------------------------------------------
// GEMM tile core: acc[TM][TN] += As[k][rows of thread] * Bs[k][cols of thread]
// thread (tx, ty) owns rows  {c*64 + ty*4 + i}  and cols {c*64 + tx*4 + j}.
// ---------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, const float* __restrict__ Bs, int tx, int ty,
                                         float (&acc)[TM][TN]) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    float a[TM], b[TN];
#pragma unroll
    for (int c = 0; c < TM / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(As + kk * BM + c * 64 + ty * 4);
      a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int c = 0; c < TN / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(Bs + kk * BN + c * 64 + tx * 4);
      b[4 * c] = v.x; b[4 * c + 1] = v.y; b[4 * c + 2] = v.z; b[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// Load 4 consecutive reduce-dim elements src[0..3] (bounds: r0+j < R), vector if allowed.
__device__ __forceinline__ float4 load4_guard(const float* __restrict__ src, int r0, int R, bool vec_ok, bool row_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  if (vec_ok && r0 + 3 < R) return __ldg(reinterpret_cast<const float4*>(src));
  if (r0 < R) v.x = __ldg(src);
  if (r0 + 1 < R) v.y = __ldg(src + 1);
  if (r0 + 2 < R) v.z = __ldg(src + 2);
  if (r0 + 3 < R) v.w = __ldg(src + 3);
  return v;
}

// ---------------------------------------------------------------------------
// dense_fwd: Y = LayerNorm(X @ W + b) -> ReLU  [-> head]
//   MODE 0: write H          (inference hidden layer)
//   MODE 1: write H, XHAT, RSTD  (training forward: what the backward needs)
//   MODE 2: write Q = H @ Wh + bh only (inference last layer, Q-head fused)
//   MODE 3: write the raw pre-activation X @ W + b to H (no LayerNorm; modular NORM_TYPE path)
// grid = (ceil(rows/BM), S)
// ---------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(GT) dense_fwd_kernel(
    const float* __restrict__ X, int64_t x_seed_stride, int ldx, const float* __restrict__ params, int64_t P,
    int64_t off_w, int64_t off_b, int64_t off_scale, int64_t off_bias, int64_t off_hw, int64_t off_hb, int A,
    float* __restrict__ H, float* __restrict__ XHAT, float* __restrict__ RSTD, float* __restrict__ Q, int rows,
    int K) {
  constexpr int BM = (BN == 128) ? 128 : 64;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_LD = BM / 64, B_LD = BN / 64;
  __shared__ __align__(16) float As[2][BK * BM];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int seed = blockIdx.y;
  const int m0 = blockIdx.x * BM;
  const float* __restrict__ Xs = X + (int64_t)seed * x_seed_stride;
  const float* __restrict__ prm = params + (int64_t)seed * P;
  const float* __restrict__ W = prm + off_w;
  const bool vecA = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(Xs) & 15) == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  float4 ra[A_LD], rb[B_LD];
  auto gload = [&](int kt) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int f = tid + i * GT;
      const int m = f >> 2, r4 = f & 3;
      const int row = m0 + m;
      const int k = kt * BK + r4 * 4;
      ra[i] = load4_guard(Xs + (int64_t)row * ldx + k, k, K, vecA, row < rows);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int f = tid + i * GT;
      const int r = f / (BN / 4), n4 = f % (BN / 4);
      const int k = kt * BK + r;
      rb[i] = (k < K) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * BN + n4 * 4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int f = tid + i * GT;
      const int m = f >> 2, r4 = f & 3;
      float* dst = &As[buf][(r4 * 4) * BM + m];
      dst[0] = ra[i].x; dst[BM] = ra[i].y; dst[2 * BM] = ra[i].z; dst[3 * BM] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int f = tid + i * GT;
      const int r = f / (BN / 4), n4 = f % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][r * BN + n4 * 4]) = rb[i];
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) gload(kt + 1);
    tile_mma<BM, BN, TM, TN>(As[kt & 1], Bs[kt & 1], tx, ty, acc);
    if (kt + 1 < nk) sstore((kt + 1) & 1);
    __syncthreads();
  }

  // ---- epilogue: bias, LayerNorm over the BN columns of each row, ReLU
  const float* __restrict__ bvec = prm + (off_b >= 0 ? off_b : 0);
  const float* __restrict__ sc = prm + off_scale;
  const float* __restrict__ bi = prm + off_bias;
  float colb[TN], cols_[TN], colbi[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int col = (j >> 2) * 64 + tx * 4 + (j & 3);
    colb[j] = off_b >= 0 ? __ldg(bvec + col) : 0.f; cols_[j] = __ldg(sc + col); colbi[j] = __ldg(bi + col);
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (MODE == 3) {  // raw pre-activation (bias only): the modular NORM_TYPE path normalises in its own kernels
      if (row < rows) {
        const int64_t grow3 = (int64_t)seed * rows + row;
#pragma unroll
        for (int c = 0; c < TN / 4; ++c)
          *reinterpret_cast<float4*>(H + grow3 * BN + c * 64 + tx * 4) =
              make_float4(acc[i][4 * c] + colb[4 * c], acc[i][4 * c + 1] + colb[4 * c + 1], acc[i][4 * c + 2] + colb[4 * c + 2],
                          acc[i][4 * c + 3] + colb[4 * c + 3]);
      }
      continue;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[i][j] += colb[j];
      s1 += acc[i][j];
      s2 += acc[i][j] * acc[i][j];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * (1.0f / BN);
    const float var = fmaxf(s2 * (1.0f / BN) - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    float xh[TN], h[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      xh[j] = (acc[i][j] - mean) * rstd;
      h[j] = fmaxf(xh[j] * cols_[j] + colbi[j], 0.f);
    }
    const bool row_ok = row < rows;
    const int64_t grow = (int64_t)seed * rows + row;
    if (MODE == 0 || MODE == 1) {
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < TN / 4; ++c) {
          const int col = c * 64 + tx * 4;
          *reinterpret_cast<float4*>(H + grow * BN + col) = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
          if (MODE == 1)
            *reinterpret_cast<float4*>(XHAT + grow * BN + col) =
                make_float4(xh[4 * c], xh[4 * c + 1], xh[4 * c + 2], xh[4 * c + 3]);
        }
        if (MODE == 1 && tx == 0) RSTD[grow] = rstd;
      }
    } else {
      const float* __restrict__ HW = prm + off_hw;
      const float* __restrict__ HB = prm + off_hb;
      for (int a = 0; a < A; ++a) {
        float pq = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int col = (j >> 2) * 64 + tx * 4 + (j & 3);
          pq = fmaf(h[j], __ldg(HW + (int64_t)col * A + a), pq);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) pq += __shfl_xor_sync(0xffffffffu, pq, o);
        if (tx == 0 && row_ok) Q[grow * A + a] = pq + __ldg(HB + a);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// wgrad: dW[kin][n] (+)= sum_rows X[row][kin] * dZ[row][n]
// grid = (ceil(Kin/128), N/128, S*splits); splits > 1 writes per-split partials (launch through run_wgrad_ffma)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GT) wgrad_kernel(const float* __restrict__ X, int64_t x_seed_stride, int ldx,
                                                   const float* __restrict__ DZ, int64_t dz_seed_stride, int N,
                                                   float* __restrict__ grads, int64_t P, int64_t off_w, int rows,
                                                   int Kin, int splits, float* __restrict__ part) {
  constexpr int BM = 128, BN = 128, TM = 8, TN = 8;
  __shared__ __align__(16) float As[2][BK * BM];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int seed = blockIdx.z / splits, split = blockIdx.z % splits;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* __restrict__ Xs = X + (int64_t)seed * x_seed_stride;
  const float* __restrict__ Zs = DZ + (int64_t)seed * dz_seed_stride;
  int chunk = (rows + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  const int r_begin = split * chunk;
  const int r_end = min(rows, r_begin + chunk);
  const bool vecA = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(Xs) & 15) == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (r_end > r_begin) ? (r_end - r_begin + BK - 1) / BK : 0;
  float4 ra[2], rb[2];
  auto gload = [&](int kt) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * GT;
      const int r = f >> 5, m4 = f & 31;
      const int row = r_begin + kt * BK + r;
      const int kin = m0 + m4 * 4;
      ra[i] = load4_guard(Xs + (int64_t)row * ldx + kin, kin, Kin, vecA, row < r_end);
      rb[i] = (row < r_end) ? __ldg(reinterpret_cast<const float4*>(Zs + (int64_t)row * N + n0 + m4 * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * GT;
      const int r = f >> 5, m4 = f & 31;
      *reinterpret_cast<float4*>(&As[buf][r * BM + m4 * 4]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][r * BN + m4 * 4]) = rb[i];
    }
  };
  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) gload(kt + 1);
    tile_mma<BM, BN, TM, TN>(As[kt & 1], Bs[kt & 1], tx, ty, acc);
    if (kt + 1 < nk) sstore((kt + 1) & 1);
    __syncthreads();
  }
  // splits == 1: the gradient itself; otherwise this row range's partial [split][seed][Kin][N], added in split order by
  // wgrad_split_reduce_kernel (no float atomics)
  float* __restrict__ dW = splits == 1 ? grads + (int64_t)seed * P + off_w
                                       : part + ((int64_t)split * (gridDim.z / splits) + seed) * Kin * N;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int kin = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (kin >= Kin) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      dW[(int64_t)kin * N + n] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------
// dgrad: OUT[row][c] = (HPREV[row][c] > 0) * sum_n dZ[row][n] * W[c][n]
// (ReLU mask of the previous layer fused; OUT may alias HPREV)
// grid = (ceil(rows/128), Kprev/128, S)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GT) dgrad_kernel(const float* __restrict__ DZ, int64_t dz_seed_stride, int N,
                                                   const float* __restrict__ params, int64_t P, int64_t off_w,
                                                   const float* HPREV, float* OUT, int64_t h_seed_stride, int rows,
                                                   int Kprev, int accumulate = 0) {
  constexpr int BM = 128, BN = 128, TM = 8, TN = 8;
  __shared__ __align__(16) float As[2][BK * BM];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int seed = blockIdx.z;
  const int m0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
  const float* __restrict__ Zs = DZ + (int64_t)seed * dz_seed_stride;
  const float* __restrict__ W = params + (int64_t)seed * P + off_w;  // [Kprev][N]

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int nk = N / BK;
  float4 ra[2], rb[2];
  auto gload = [&](int kt) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * GT;
      const int m = f >> 2, r4 = f & 3;
      const int n = kt * BK + r4 * 4;
      const int row = m0 + m;
      ra[i] = (row < rows) ? __ldg(reinterpret_cast<const float4*>(Zs + (int64_t)row * N + n))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(c0 + m) * N + n));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * GT;
      const int m = f >> 2, r4 = f & 3;
      float* da = &As[buf][(r4 * 4) * BM + m];
      da[0] = ra[i].x; da[BM] = ra[i].y; da[2 * BM] = ra[i].z; da[3 * BM] = ra[i].w;
      float* db = &Bs[buf][(r4 * 4) * BN + m];
      db[0] = rb[i].x; db[BN] = rb[i].y; db[2 * BN] = rb[i].z; db[3 * BN] = rb[i].w;
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) gload(kt + 1);
    tile_mma<BM, BN, TM, TN>(As[kt & 1], Bs[kt & 1], tx, ty, acc);
    if (kt + 1 < nk) sstore((kt + 1) & 1);
    __syncthreads();
  }
  const float* Hs = HPREV + (int64_t)seed * h_seed_stride;   // ReLU mask source (the layer output: h > 0)
  float* Os = OUT + (int64_t)seed * h_seed_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (row >= rows) continue;
#pragma unroll
    for (int c = 0; c < TN / 4; ++c) {
      const int col = c0 + c * 64 + tx * 4;
      const float4 hv = *reinterpret_cast<const float4*>(Hs + (int64_t)row * Kprev + col);
      float4 o;
      o.x = hv.x > 0.f ? acc[i][4 * c] : 0.f;
      o.y = hv.y > 0.f ? acc[i][4 * c + 1] : 0.f;
      o.z = hv.z > 0.f ? acc[i][4 * c + 2] : 0.f;
      o.w = hv.w > 0.f ? acc[i][4 * c + 3] : 0.f;
      if (accumulate) {   // OUT += ... (sum of several products, e.g. the three GRU gates; OUT must not alias HPREV)
        const float4 prev = *reinterpret_cast<const float4*>(Os + (int64_t)row * Kprev + col);
        o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
      }
      *reinterpret_cast<float4*>(Os + (int64_t)row * Kprev + col) = o;
    }
  }
}

// ---------------------------------------------------------------------------
// row backward: one warp per sample.
//   HEAD=1: q = h @ Wh + bh, loss/qsa accumulation, dq, dWh, dbh, dh = dq*Wh[:,a],
//           ReLU mask from h.      HEAD=0: dy read from DH (already ReLU-masked).
//   then LayerNorm backward -> DZ, accumulating dscale, dbias and db (= sum dz).
// grid = (ceil(rows/ROWS_PER_CTA), S)
// ---------------------------------------------------------------------------

template <int N, bool HEAD>
__global__ void __launch_bounds__(256) row_bwd_kernel(
    const float* __restrict__ Hh, const float* __restrict__ XHAT, const float* __restrict__ RSTD,
    const float* DH, float* DZ, float* DZLO, __half* DZ16H, __half* DZ16L, float gscale,
    const float* __restrict__ params, float* __restrict__ grads, int64_t P,
    int64_t off_scale, int64_t off_dscale, int64_t off_dbias, int64_t off_db, int64_t off_hw, int64_t off_hb, int A,
    const int32_t* __restrict__ gather, const int32_t* __restrict__ action, const float* __restrict__ target,
    int64_t tr_rows_per_seed, float* __restrict__ part, int rows) {
  // Reductions over rows are deterministic: lane-private registers -> warp-private shared-memory slices -> a fixed-order
  // sum over the 8 warps -> this CTA's partial vector part[seed][cta][3N (+ A*N + A + 2)], which row_bwd_final_kernel
  // adds over the CTAs in index order (no float atomics anywhere).
  constexpr int F = N / 32;  // features per lane, in float4 chunks at lane*4 + c*128
  extern __shared__ float smem[];
  float* s_red3 = smem;           // [8 warps][3N]  d scale, d bias, d dense-bias slices
  float* s_hw = smem + 24 * N;    // [A][N]  head weights, transposed copy (HEAD)
  float* s_dhw = s_hw + (HEAD ? A * N : 0);   // [8 warps][A][N]  warp-private head-weight gradient slices (HEAD);
                                              // 16-byte aligned for the float4 accesses, scalars go last
  float* s_dhb = s_dhw + (HEAD ? 8 * A * N : 0);  // [8 warps][A]
  float* s_ls = s_dhb + (HEAD ? 8 * A : 0);       // [8 warps][2] loss, qsa
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seed = blockIdx.y;
  const float* __restrict__ prm = params + (int64_t)seed * P;
  const int nsm = 24 * N + (HEAD ? A * N + 8 * A + 16 + 8 * A * N : 0);
  for (int i = tid; i < nsm; i += 256) smem[i] = 0.f;
  __syncthreads();
  if (HEAD)
    for (int i = tid; i < A * N; i += 256) s_hw[(i % A) * N + i / A] = __ldg(prm + off_hw + i);
  __syncthreads();
  float scale[F];
#pragma unroll
  for (int j = 0; j < F; ++j) scale[j] = __ldg(prm + off_scale + (j >> 2) * 128 + lane * 4 + (j & 3));
  float a_dsc[F], a_dbi[F], a_db[F];
#pragma unroll
  for (int j = 0; j < F; ++j) a_dsc[j] = a_dbi[j] = a_db[j] = 0.f;
  float a_loss = 0.f, a_qsa = 0.f;
  const float invB = 1.0f / (float)rows;
  float* my_dhw = s_dhw + warp * A * N;

  // grid-stride over the seed's rows, one row per warp (grid.x is sized in whole waves, see conv_mma_ctas)
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const int64_t grow = (int64_t)seed * rows + row;
    float h[F], xh[F], dy[F];
#pragma unroll
    for (int c = 0; c < F / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(XHAT + grow * N + c * 128 + lane * 4);
      xh[4 * c] = v.x; xh[4 * c + 1] = v.y; xh[4 * c + 2] = v.z; xh[4 * c + 3] = v.w;
    }
    const float rstd = RSTD[grow];
    if (HEAD) {
#pragma unroll
      for (int c = 0; c < F / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(Hh + grow * N + c * 128 + lane * 4);
        h[4 * c] = v.x; h[4 * c + 1] = v.y; h[4 * c + 2] = v.z; h[4 * c + 3] = v.w;
      }
      const int src = gather ? gather[(int64_t)seed * rows + row] : row;
      const int act = action[(int64_t)seed * tr_rows_per_seed + src];
      const float tgt = target[(int64_t)seed * tr_rows_per_seed + src];
      // q_sa only needs column `act` of the head
      float pq = 0.f;
      float wcol[F];
#pragma unroll
      for (int c = 0; c < F / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(s_hw + act * N + c * 128 + lane * 4);
        wcol[4 * c] = v.x; wcol[4 * c + 1] = v.y; wcol[4 * c + 2] = v.z; wcol[4 * c + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < F; ++j) pq = fmaf(h[j], wcol[j], pq);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pq += __shfl_xor_sync(0xffffffffu, pq, o);
      const float q_sa = pq + __ldg(prm + off_hb + act);
      const float diff = q_sa - tgt;
      const float dq = diff * invB;
      if (lane == 0) {
        a_loss += 0.5f * diff * diff * invB;
        a_qsa += q_sa * invB;
        s_dhb[warp * A + act] += dq;   // warp-private slot, one writer
      }
#pragma unroll
      for (int c = 0; c < F / 4; ++c) {  // warp-private slice: plain read-modify-write, no atomics
        float4* dst = reinterpret_cast<float4*>(my_dhw + act * N + c * 128 + lane * 4);
        float4 v = *dst;
        v.x = fmaf(h[4 * c], dq, v.x); v.y = fmaf(h[4 * c + 1], dq, v.y);
        v.z = fmaf(h[4 * c + 2], dq, v.z); v.w = fmaf(h[4 * c + 3], dq, v.w);
        *dst = v;
      }
#pragma unroll
      for (int j = 0; j < F; ++j) dy[j] = h[j] > 0.f ? dq * wcol[j] : 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < F / 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(DH + grow * N + c * 128 + lane * 4);
        dy[4 * c] = v.x; dy[4 * c + 1] = v.y; dy[4 * c + 2] = v.z; dy[4 * c + 3] = v.w;
      }
    }
    float m1 = 0.f, m2 = 0.f;
    float dxh[F];
#pragma unroll
    for (int j = 0; j < F; ++j) {
      a_dsc[j] = fmaf(dy[j], xh[j], a_dsc[j]);
      a_dbi[j] += dy[j];
      dxh[j] = dy[j] * scale[j];
      m1 += dxh[j];
      m2 = fmaf(dxh[j], xh[j], m2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.0f / N);
    m2 *= (1.0f / N);
    float dz[F];
#pragma unroll
    for (int j = 0; j < F; ++j) {
      dz[j] = rstd * (dxh[j] - m1 - xh[j] * m2);
      a_db[j] += dz[j];
    }
#pragma unroll
    for (int c = 0; c < F / 4; ++c)
    {
      if (DZ16H != nullptr) {  // fp16-split planes of dz * gscale for the tensor-core wgrad / dgrad (no fp32 copy)
        __half2 h0, h1, l0, l1;
        tc::split16x2(dz[4 * c] * gscale, dz[4 * c + 1] * gscale, h0, l0);
        tc::split16x2(dz[4 * c + 2] * gscale, dz[4 * c + 3] * gscale, h1, l1);
        *reinterpret_cast<uint2*>(DZ16H + grow * N + c * 128 + lane * 4) =
            make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(DZ16L + grow * N + c * 128 + lane * 4) =
            make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
        continue;
      }
      *reinterpret_cast<float4*>(DZ + grow * N + c * 128 + lane * 4) =
          make_float4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
      if (DZLO != nullptr)
        *reinterpret_cast<float4*>(DZLO + grow * N + c * 128 + lane * 4) =
            make_float4(tc::tf32_lo(dz[4 * c]), tc::tf32_lo(dz[4 * c + 1]), tc::tf32_lo(dz[4 * c + 2]),
                        tc::tf32_lo(dz[4 * c + 3]));
    }
  }
  // warp-private slices (each lane owns its features: plain stores)
#pragma unroll
  for (int j = 0; j < F; ++j) {
    const int f = (j >> 2) * 128 + lane * 4 + (j & 3);
    s_red3[warp * 3 * N + f] = a_dsc[j];
    s_red3[warp * 3 * N + N + f] = a_dbi[j];
    s_red3[warp * 3 * N + 2 * N + f] = a_db[j];
  }
  if (HEAD && lane == 0) { s_ls[warp * 2] = a_loss; s_ls[warp * 2 + 1] = a_qsa; }
  __syncthreads();
  const int stride = 3 * N + (HEAD ? A * N + A + 2 : 0);
  float* __restrict__ o = part + ((int64_t)seed * gridDim.x + blockIdx.x) * stride;
  for (int i = tid; i < 3 * N; i += 256) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += s_red3[wv * 3 * N + i];
    o[i] = v;
  }
  if (HEAD) {
    for (int i = tid; i < A * N; i += 256) {   // [A][N]
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) v += s_dhw[wv * A * N + i];
      o[3 * N + i] = v;
    }
    if (tid < A + 2) {
      float v = 0.f;
      for (int wv = 0; wv < 8; ++wv) v += tid < A ? s_dhb[wv * A + tid] : s_ls[wv * 2 + (tid - A)];
      o[3 * N + A * N + tid] = v;
    }
  }
}

// Fixed-order column sum of `count` partial vectors (p[c * stride], c = 0..count-1) by a 256-thread block laid out as
// (256 / SL) columns x SL slices: slice s adds the partials c = s, s+SL, ... in order, the SL slice sums are then added
// in slice order.  Every thread of the block must call it; the result is valid in the threads of slice 0
// (threadIdx.x < 256 / SL).  SL = 8 for a few partials per seed (many seeds), 32 for hundreds (one seed): the chain of
// dependent-latency loads per thread stays short either way.
template <int SL>
__device__ __forceinline__ float ordered_partial_sum(const float* __restrict__ p, int count, int64_t stride, bool valid) {
  constexpr int COLS = 256 / SL;
  __shared__ float slice_sum[SL][COLS];
  const int col = threadIdx.x % COLS, sl = threadIdx.x / COLS;
  float v = 0.f;
  if (valid) {
#pragma unroll 4
    for (int c = sl; c < count; c += SL) v += p[(int64_t)c * stride];
  }
  slice_sum[sl][col] = v;
  __syncthreads();
  float r = 0.f;
  if (sl == 0) {
#pragma unroll
    for (int k = 0; k < SL; ++k) r += slice_sum[k][col];
  }
  return r;
}
// slices for `count` partials per output element
static inline int final_slices(int count) { return count > 48 ? 32 : 8; }

// Sums the per-CTA partial vectors of row_bwd_kernel in a fixed order and writes the gradients (and adds the
// minibatch's loss / mean q_sa to the running sums).  grid = (ceil(stride / (256 / SL)), S), block = 256
template <int SL>
__global__ void row_bwd_final_kernel(const float* __restrict__ part, int nctas, int N, int A, int head,
                                     float* __restrict__ grads, int64_t P, int64_t off_dscale, int64_t off_dbias,
                                     int64_t off_db, int64_t off_hw, int64_t off_hb, float* __restrict__ loss_sum,
                                     float* __restrict__ qsa_sum) {
  const int seed = blockIdx.y;
  const int stride = 3 * N + (head ? A * N + A + 2 : 0);
  const int i = blockIdx.x * (256 / SL) + threadIdx.x % (256 / SL);
  const float v = ordered_partial_sum<SL>(part + (int64_t)seed * nctas * stride + i, nctas, stride, i < stride);
  if (i >= stride || threadIdx.x >= 256 / SL) return;
  float* __restrict__ g = grads + (int64_t)seed * P;
  if (i < N) g[off_dscale + i] = v;
  else if (i < 2 * N) g[off_dbias + (i - N)] = v;
  else if (i < 3 * N) g[off_db + (i - 2 * N)] = v;
  else if (i < 3 * N + A * N) {
    const int j = i - 3 * N, a = j / N, f = j - a * N;   // partial layout [A][N]; the parameter is [N][A]
    g[off_hw + (int64_t)f * A + a] = v;
  } else if (i < 3 * N + A * N + A) g[off_hb + (i - 3 * N - A * N)] = v;
  else if (i == 3 * N + A * N + A) loss_sum[seed] += v;
  else qsa_sum[seed] += v;
}

// ---------------------------------------------------------------------------
// MinAtar conv 3x3 (C -> 16, VALID) + LayerNorm(16) + ReLU from bit-packed obs.
// thread = (sample, output pixel); block = 4 samples; grid = (ceil(rows/4), S)
// ---------------------------------------------------------------------------
template <int C>
struct ConvCfg {
  static constexpr int TAPS = 9 * C;
  static constexpr int OBS_BITS = 100 * C;
  static constexpr int OBS_WORDS = (OBS_BITS + 31) / 32;
  static constexpr int PW = (OBS_WORDS + 3) / 4 * 4;       // packed row words (matches env OBS_WORDS_PAD)
  static constexpr int SW = PW + 1;                        // smem row (+1 so the funnel shift may read past the end)
};

// the C channel bits of input pixel p of a packed row held in shared memory
template <int C>
__device__ __forceinline__ uint32_t pixel_bits(const uint32_t* __restrict__ so, int p) {
  const int f0 = p * C;
  const uint32_t lo = so[f0 >> 5], hi = so[(f0 >> 5) + 1];
  return __funnelshift_r(lo, hi, f0 & 31) & ((1u << C) - 1u);
}

// conv pre-activation for one output pixel; ws = conv kernel * (1/255), [tap][16]
template <int C>
__device__ __forceinline__ void conv_pixel(const uint32_t* __restrict__ so, const float* __restrict__ ws,
                                           const float* __restrict__ cb, int y, int x, float (&acc)[CONV_O]) {
#pragma unroll
  for (int o = 0; o < CONV_O; ++o) acc[o] = cb[o];
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) {
      const uint32_t nib = pixel_bits<C>(so, (y + di) * 10 + (x + dj));
      if (__ballot_sync(0xffffffffu, nib != 0u) == 0u) continue;  // nobody in the warp has this patch cell set
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const bool bit = (nib >> c) & 1u;
        if (__ballot_sync(0xffffffffu, bit) == 0u) continue;
        const float* __restrict__ w = ws + ((di * 3 + dj) * C + c) * CONV_O;
#pragma unroll
        for (int o4 = 0; o4 < CONV_O / 4; ++o4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + 4 * o4);
          if (bit) {
            acc[4 * o4] += wv.x; acc[4 * o4 + 1] += wv.y; acc[4 * o4 + 2] += wv.z; acc[4 * o4 + 3] += wv.w;
          }
        }
      }
    }
}

__device__ __forceinline__ void ln16(const float (&z)[CONV_O], float& mean, float& rstd) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int o = 0; o < CONV_O; ++o) { s1 += z[o]; s2 = fmaf(z[o], z[o], s2); }
  mean = s1 * (1.0f / CONV_O);
  const float var = fmaxf(s2 * (1.0f / CONV_O) - mean * mean, 0.f);
  rstd = 1.0f / sqrtf(var + LN_EPS);
}

template <int C>
__device__ __forceinline__ void conv_load_consts(const float* __restrict__ prm, const pqn_net_layout_t& L, float* ws,
                                                 float* cb, float* sc, float* bi) {
  const float inv255 = 1.0f / 255.0f;  // x/255 for x in {0,1}  (pqn_minatar.py:66)
  for (int i = threadIdx.x; i < ConvCfg<C>::TAPS * CONV_O; i += blockDim.x) ws[i] = __ldg(prm + L.conv_w + i) * inv255;
  if (threadIdx.x < CONV_O) {
    cb[threadIdx.x] = __ldg(prm + L.conv_b + threadIdx.x);
    sc[threadIdx.x] = __ldg(prm + L.ln0_scale + threadIdx.x);
    bi[threadIdx.x] = __ldg(prm + L.ln0_bias + threadIdx.x);
  }
}

template <int C, bool TRAIN>
__global__ void __launch_bounds__(256) conv_fwd_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed,
                                                       const int32_t* __restrict__ gather,
                                                       const float* __restrict__ params, int64_t P,
                                                       pqn_net_layout_t L, float* __restrict__ H1,
                                                       float* __restrict__ H1LO, float* __restrict__ bn_sums,
                                                       int rows) {
  using Cfg = ConvCfg<C>;
  __shared__ __align__(16) float ws[Cfg::TAPS * CONV_O];
  __shared__ float cb[CONV_O], sc[CONV_O], bi[CONV_O];
  __shared__ uint32_t so[4][Cfg::SW];
  __shared__ float s_cnt[C];
  const int tid = threadIdx.x, sl = tid >> 6, pix = tid & 63;
  const int seed = blockIdx.y;
  const int row = blockIdx.x * 4 + sl;
  const bool valid = row < rows;
  conv_load_consts<C>(params + (int64_t)seed * P, L, ws, cb, sc, bi);
  if (TRAIN && tid < C) s_cnt[tid] = 0.f;
  if (pix < Cfg::SW) {
    uint32_t w = 0u;
    if (valid && pix < Cfg::PW) {
      const int64_t src = gather ? gather[(int64_t)seed * rows + row] : row;
      w = __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + pix);
    }
    so[sl][pix] = w;
  }
  __syncthreads();
  const int y = pix >> 3, x = pix & 7;
  float acc[CONV_O];
  conv_pixel<C>(so[sl], ws, cb, y, x, acc);
  float mean, rstd;
  ln16(acc, mean, rstd);
  if (valid) {
    float4* __restrict__ out = reinterpret_cast<float4*>(H1 + ((int64_t)seed * rows + row) * FLAT_CNN + pix * CONV_O);
#pragma unroll
    for (int o4 = 0; o4 < CONV_O / 4; ++o4) {
      float4 v;
      v.x = fmaxf((acc[4 * o4] - mean) * rstd * sc[4 * o4] + bi[4 * o4], 0.f);
      v.y = fmaxf((acc[4 * o4 + 1] - mean) * rstd * sc[4 * o4 + 1] + bi[4 * o4 + 1], 0.f);
      v.z = fmaxf((acc[4 * o4 + 2] - mean) * rstd * sc[4 * o4 + 2] + bi[4 * o4 + 2], 0.f);
      v.w = fmaxf((acc[4 * o4 + 3] - mean) * rstd * sc[4 * o4 + 3] + bi[4 * o4 + 3], 0.f);
      out[o4] = v;
      if (H1LO != nullptr) {  // 3xTF32 error-compensation operand for the tcgen05 GEMMs
        float4* __restrict__ olo =
            reinterpret_cast<float4*>(H1LO + ((int64_t)seed * rows + row) * FLAT_CNN + pix * CONV_O);
        olo[o4] = make_float4(tc::tf32_lo(v.x), tc::tf32_lo(v.y), tc::tf32_lo(v.z), tc::tf32_lo(v.w));
      }
    }
  }
  if (TRAIN && bn_sums != nullptr) {
    // dummy input BatchNorm statistics: per-channel count of set bits (x in {0,1} => sum x == sum x^2)
    uint32_t b0 = pixel_bits<C>(so[sl], pix);
    uint32_t b1 = (pix + 64 < 100) ? pixel_bits<C>(so[sl], pix + 64) : 0u;
    if (!valid) { b0 = 0u; b1 = 0u; }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int cnt = (int)((b0 >> c) & 1u) + (int)((b1 >> c) & 1u);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if ((tid & 31) == 0 && cnt) atomicAdd(&s_cnt[c], (float)cnt);
    }
    __syncthreads();
    if (tid < C && s_cnt[tid] != 0.f) {
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + tid, s_cnt[tid]);
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + C + tid, s_cnt[tid]);
    }
  }
}

// conv backward, one warp per sample (no block-level sync in the loop):
//   phase A  lane = output pixel (2 per lane): recompute conv + LN statistics, LN backward from DY1 (already
//            ReLU-masked by dgrad) -> dz staged in the warp's shared memory; d(ln0 scale/bias) in registers.
//   phase B  dW[tap][o] += x[pixel+tap] * dz[pixel][o] driven by the SET input bits only (MinAtar observations
//            are sparse): for each set bit (input pixel q, channel c) the 9 taps that see it add dz[q - tap][:]
//            into lane-private accumulators; lane = (tap parity, o), accumulators indexed [c][tap/2] statically.
//   d(conv bias)[o] = sum over pixels of dz, also taken from the staged dz.
// grid = (CONV_BWD_CTAS_X, S); each CTA strides over the seed's samples 8 at a time.
constexpr int CONV_BWD_WARPS = 8;
constexpr int SDZ_LD = 20;  // padded dz row (conflict-free 128-bit stores from 32 pixel-lanes)

template <int C>
__global__ void __launch_bounds__(CONV_BWD_WARPS * 32, 2)
    conv_bwd_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                    const float* __restrict__ params, int64_t P, pqn_net_layout_t L, const float* __restrict__ DY1,
                    float* __restrict__ grads, int rows) {
  using Cfg = ConvCfg<C>;
  constexpr int TAPS = Cfg::TAPS;
  __shared__ __align__(16) float ws[TAPS * CONV_O];
  __shared__ float cb[CONV_O], sc[CONV_O], bi[CONV_O];
  __shared__ uint32_t so[CONV_BWD_WARPS][Cfg::SW];
  __shared__ __align__(16) float sdz[CONV_BWD_WARPS][CONV_PIX * SDZ_LD];
  __shared__ float s_red[3 * CONV_O];
  float* s_w = &sdz[0][0];  // [TAPS*16], aliases the dz staging area once the sample loop is done
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seed = blockIdx.y;
  const float* __restrict__ prm = params + (int64_t)seed * P;
  conv_load_consts<C>(prm, L, ws, cb, sc, bi);
  if (tid < 3 * CONV_O) s_red[tid] = 0.f;
  __syncthreads();

  float a_dsc[CONV_O], a_dbi[CONV_O];
#pragma unroll
  for (int o = 0; o < CONV_O; ++o) a_dsc[o] = a_dbi[o] = 0.f;
  // phase-B role: lane = (tap parity, output channel); 5 iterations cover the 9 taps
  const int o_b = lane & 15, tpar = lane >> 4;
  int t_di[5], t_dj[5];
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const int tap = it * 2 + tpar;
    t_di[it] = tap < 9 ? tap / 3 : 100;  // 100 => never valid
    t_dj[it] = tap < 9 ? tap % 3 : 0;
  }
  float wacc[C][5];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int it = 0; it < 5; ++it) wacc[c][it] = 0.f;
  float a_dcb = 0.f;  // lane (tpar, o): sum of dz[pixel][o] over pixels of parity tpar

  uint32_t* __restrict__ my_so = so[warp];
  float* __restrict__ my_dz = sdz[warp];
  for (int row = blockIdx.x * CONV_BWD_WARPS + warp; row < rows; row += gridDim.x * CONV_BWD_WARPS) {
    __syncwarp();
    {
      const int64_t src = gather ? gather[(int64_t)seed * rows + row] : row;
      const uint32_t* __restrict__ orow = obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW;
      for (int wi = lane; wi < Cfg::SW; wi += 32) my_so[wi] = wi < Cfg::PW ? __ldg(orow + wi) : 0u;
    }
    __syncwarp();
    // ---- phase A: two output pixels per lane
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      const int pix = lane + 32 * hh;
      float acc[CONV_O];
      conv_pixel<C>(my_so, ws, cb, pix >> 3, pix & 7, acc);
      float mean, rstd;
      ln16(acc, mean, rstd);
      const float4* __restrict__ dyp =
          reinterpret_cast<const float4*>(DY1 + ((int64_t)seed * rows + row) * FLAT_CNN + pix * CONV_O);
      float dy[CONV_O];
#pragma unroll
      for (int o4 = 0; o4 < CONV_O / 4; ++o4) {
        const float4 v = __ldg(dyp + o4);
        dy[4 * o4] = v.x; dy[4 * o4 + 1] = v.y; dy[4 * o4 + 2] = v.z; dy[4 * o4 + 3] = v.w;
      }
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int o = 0; o < CONV_O; ++o) {
        acc[o] = (acc[o] - mean) * rstd;  // xhat
        a_dsc[o] = fmaf(dy[o], acc[o], a_dsc[o]);
        a_dbi[o] += dy[o];
        dy[o] *= sc[o];                   // dxhat
        m1 += dy[o];
        m2 = fmaf(dy[o], acc[o], m2);
      }
      m1 *= (1.0f / CONV_O);
      m2 *= (1.0f / CONV_O);
#pragma unroll
      for (int o4 = 0; o4 < CONV_O / 4; ++o4) {
        float4 dz;
        dz.x = rstd * (dy[4 * o4] - m1 - acc[4 * o4] * m2);
        dz.y = rstd * (dy[4 * o4 + 1] - m1 - acc[4 * o4 + 1] * m2);
        dz.z = rstd * (dy[4 * o4 + 2] - m1 - acc[4 * o4 + 2] * m2);
        dz.w = rstd * (dy[4 * o4 + 3] - m1 - acc[4 * o4 + 3] * m2);
        *reinterpret_cast<float4*>(&my_dz[pix * SDZ_LD + 4 * o4]) = dz;
      }
    }
    __syncwarp();
    // ---- d(conv bias): lane (tpar, o) sums dz over the pixels of its parity
#pragma unroll 8
    for (int p = tpar; p < CONV_PIX; p += 2) a_dcb += my_dz[p * SDZ_LD + o_b];
    // ---- phase B: sparse dW accumulation over the set input bits
#pragma unroll
    for (int c = 0; c < C; ++c) {
      for (int wi = 0; wi < Cfg::OBS_WORDS; ++wi) {
        // bits of word wi that belong to channel c: flat index f = wi*32 + b with f % C == c
        uint32_t bits = my_so[wi];
        if (C == 4) bits &= 0x11111111u << c;  // 32 % C == 0: channel c sits at a fixed bit phase in every word
        if (bits == 0u) continue;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1u;
          const int f = wi * 32 + b;
          const int q = f / C;
          if (f - q * C != c) continue;
          const int qy = q / 10, qx = q - qy * 10;
#pragma unroll
          for (int it = 0; it < 5; ++it) {
            const int py = qy - t_di[it], px = qx - t_dj[it];
            if ((unsigned)py < 8u && (unsigned)px < 8u) wacc[c][it] += my_dz[(py * 8 + px) * SDZ_LD + o_b];
          }
        }
      }
    }
  }
  // ---- reduce and publish
  __syncthreads();
  for (int i = tid; i < TAPS * CONV_O; i += blockDim.x) s_w[i] = 0.f;
  __syncthreads();
  const float inv255 = 1.0f / 255.0f;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int tap = it * 2 + tpar;
      if (tap < 9) atomicAdd(&s_w[(tap * C + c) * CONV_O + o_b], wacc[c][it] * inv255);
    }
  atomicAdd(&s_red[2 * CONV_O + o_b], a_dcb);
#pragma unroll
  for (int o = 0; o < CONV_O; ++o) {
    float v0 = a_dsc[o], v1 = a_dbi[o];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, sft);
      v1 += __shfl_xor_sync(0xffffffffu, v1, sft);
    }
    if (lane == 0) {
      atomicAdd(&s_red[o], v0);
      atomicAdd(&s_red[CONV_O + o], v1);
    }
  }
  __syncthreads();
  float* __restrict__ g = grads + (int64_t)seed * P;
  for (int i = tid; i < TAPS * CONV_O; i += blockDim.x) atomicAdd(g + L.conv_w + i, s_w[i]);
  if (tid < CONV_O) {
    atomicAdd(g + L.ln0_scale + tid, s_red[tid]);
    atomicAdd(g + L.ln0_bias + tid, s_red[CONV_O + tid]);
    atomicAdd(g + L.conv_b + tid, s_red[2 * CONV_O + tid]);
  }
}

// ---------------------------------------------------------------------------
// conv on the warp-level tensor-core path (mma.sync.m16n8k8 tf32, fp32 accumulate).
// The 3x3 conv is a skinny GEMM  Z[64 pixels x 16] = Xcol[64 x 9C] . W[9C x 16]  per sample whose A operand is
// {0,1}: every lane builds its A-fragment elements straight from the packed observation bits (no im2col in
// memory), B fragments (weights/255 split into tf32 hi + lo: two passes keep fp32 accuracy, x is exact) come from
// shared memory, and the 16 channels of a pixel end up in the 4 lanes of a quad, so LayerNorm is two shuffles.
// The weight gradient is the transposed skinny GEMM dW[9C x 16] = Xcol^T[9C x 64] . dZ[64 x 16], accumulated per
// sample in fresh MMA accumulators (16-long chains) and added to fp32 registers (FADD) across samples.
// One warp per sample.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16n8k8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int C>
struct ConvMma {
  static constexpr int TAPS = 9 * C;
  static constexpr int KS = (TAPS + 7) / 8;    // k-steps of the forward GEMM (taps)
  static constexpr int MT = (TAPS + 15) / 16;  // m-blocks of the weight-gradient GEMM (taps)
};

// im2col "patch" of one output pixel as bits: bit k = tap k = (di*3+dj)*C + c, i.e. obs bit
// ((y+di)*10 + x+dj)*C + c.  9C <= 90 bits -> PatchCfg::WORDS words; bits beyond 9C are zero.  Built once per
// sample into shared memory (patch[pixel][word]); the MMA fragment builders then test bits with a shift instead of
// re-deriving the observation bit address for every (pixel, tap) pair.
template <int C>
struct PatchCfg {
  static constexpr int WORDS = (9 * C + 31) / 32;
};

// The 9C patch bits of output pixel `pix` as PatchCfg::WORDS words.  The packed observation is pixel-major /
// channel-minor, so the three taps (dj = 0..2) x C channels of one patch row are 3C CONSECUTIVE bits of the input row:
// three funnel-shift extracts instead of nine per-pixel ones.
template <int C>
__device__ __forceinline__ void patch_bits(const uint32_t* __restrict__ so, int pix, uint32_t (&w)[PatchCfg<C>::WORDS]) {
  constexpr int W = PatchCfg<C>::WORDS;
#pragma unroll
  for (int k = 0; k < W; ++k) w[k] = 0u;
  const int y = pix >> 3, x = pix & 7;
#pragma unroll
  for (int di = 0; di < 3; ++di) {
    const int f0 = ((y + di) * 10 + x) * C;
    const uint32_t r = __funnelshift_r(so[f0 >> 5], so[(f0 >> 5) + 1], f0 & 31) & ((1u << (3 * C)) - 1u);
    const int o = 3 * C * di;  // compile-time after unrolling
    w[o >> 5] |= r << (o & 31);
    if ((o & 31) + 3 * C > 32) w[min((o >> 5) + 1, W - 1)] |= r >> (32 - (o & 31));
  }
}

template <int C>
__device__ __forceinline__ void build_patch(const uint32_t* __restrict__ so, int pix, uint32_t* __restrict__ out) {
  constexpr int W = PatchCfg<C>::WORDS;
  uint32_t w[W];
  patch_bits<C>(so, pix, w);
#pragma unroll
  for (int k = 0; k < W; ++k) out[k] = w[k];
}

// both output pixels of a lane (pix = lane, lane + 32) -> patch[64][WORDS]
template <int C>
__device__ __forceinline__ void build_patches(const uint32_t* __restrict__ so, uint32_t* __restrict__ patch, int lane) {
  build_patch<C>(so, lane, patch + lane * PatchCfg<C>::WORDS);
  build_patch<C>(so, lane + 32, patch + (lane + 32) * PatchCfg<C>::WORDS);
}

__device__ __forceinline__ uint32_t bit_f32(uint32_t word, int shift) {
  return ((word >> shift) & 1u) ? 0x3F800000u : 0u;
}

// B fragments of the conv weights (scaled by 1/255), hi and lo, laid out [ks][half][lane] as float2 (b0, b1).
// EXPC: the A operand is "exponent coded" (see ExpPatch): the activation of tap k is the power of two
// 2^(2^(k%8) - 127) instead of 1, so row k of B carries the inverse factor 2^(127 - 2^(k%8)) (exact scaling).
template <int C, bool EXPC = false>
__device__ __forceinline__ void conv_mma_load_weights(const float* __restrict__ prm, const pqn_net_layout_t& L,
                                                      float2* wb_hi, float2* wb_lo, float* cb, float* sc, float* bi) {
  using M = ConvMma<C>;
  const float inv255 = 1.0f / 255.0f;
  for (int i = threadIdx.x; i < M::KS * 2 * 32; i += blockDim.x) {
    const int ln = i & 31, h = (i >> 5) & 1, ks = i >> 6;
    const int o = h * 8 + (ln >> 2);
    const int t0 = ks * 8 + (ln & 3), t1 = t0 + 4;
    float w0 = t0 < M::TAPS ? __ldg(prm + L.conv_w + t0 * CONV_O + o) * inv255 : 0.f;
    float w1 = t1 < M::TAPS ? __ldg(prm + L.conv_w + t1 * CONV_O + o) * inv255 : 0.f;
    if (EXPC) {
      w0 *= __uint_as_float((254u - (1u << (t0 & 7))) << 23);
      w1 *= __uint_as_float((254u - (1u << (t1 & 7))) << 23);
    }
    const float h0 = __uint_as_float(__float_as_uint(w0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(w1) & 0xFFFFE000u);
    wb_hi[i] = make_float2(h0, h1);
    wb_lo[i] = make_float2(w0 - h0, w1 - h1);
  }
  if (threadIdx.x < CONV_O) {
    cb[threadIdx.x] = __ldg(prm + L.conv_b + threadIdx.x);
    sc[threadIdx.x] = __ldg(prm + L.ln0_scale + threadIdx.x);
    bi[threadIdx.x] = __ldg(prm + L.ln0_bias + threadIdx.x);
  }
}

// ---- exponent-coded im2col fragments (forward kernel) --------------------------------------------------------
// A tf32 MMA operand only has to be *some* exactly known value when the input bit is set, not 1.0: a word whose
// bits 23..30 (the fp32 exponent field) hold eight tap bits turns into an A element with ONE instruction,
//     a = word & (1 << (23 + j))      ->  0  or  2^(2^j - 127)        (a normal power of two, mantissa 0)
// and the matching row of B is pre-multiplied by 2^(127 - 2^j) (conv_mma_load_weights<C, true>), so every product
// is bit-for-bit the product of the plain 0/1 formulation.  Per output pixel the patch is stored as KS words,
// word ks = taps 8ks .. 8ks+7 at bits 23..30; lane (g, t) of the m16n8k8 fragment needs taps t and t+4.
template <int C>
struct ExpPatch {
  static constexpr int KS = ConvMma<C>::KS;
  static constexpr int LD = KS | 1;  // odd row stride: conflict-free for the 8 pixel rows a fragment load touches
};

template <int C>
__device__ __forceinline__ void build_exp_patch(const uint32_t* __restrict__ so, int pix, uint32_t* __restrict__ out) {
  constexpr int W = PatchCfg<C>::WORDS;
  uint32_t w[W];
  patch_bits<C>(so, pix, w);
#pragma unroll
  for (int ks = 0; ks < ExpPatch<C>::KS; ++ks)  // 8 | 32: a k-step never straddles two words; bits >= 9C are zero
    out[ks] = ((w[(8 * ks) >> 5] >> ((8 * ks) & 31)) & 0xFFu) << 23;
}

// conv pre-activation of the 32 pixels of m-blocks 2*mbp and 2*mbp+1 (two blocks share every B fragment load);
// z[i][h][0..3] is the C fragment of block mb = 2*mbp+i (pixel = 16*mb + g [+8]):
// z[i][h][0], z[i][h][1] -> pixel g, channels 8h+2t, 8h+2t+1 ; z[i][h][2], z[i][h][3] -> pixel g+8, same channels.
template <int C>
__device__ __forceinline__ void conv_mma_block2_exp(const uint32_t* __restrict__ xp, const float2* __restrict__ wb_hi,
                                                    const float2* __restrict__ wb_lo, const float* __restrict__ cb,
                                                    int mbp, int lane, float (&z)[2][2][4]) {
  using E = ExpPatch<C>;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t m_lo = 1u << (23 + t), m_hi = 1u << (27 + t);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      z[i][h][0] = z[i][h][2] = cb[8 * h + 2 * t];
      z[i][h][1] = z[i][h][3] = cb[8 * h + 2 * t + 1];
    }
  const uint32_t* r00 = xp + (32 * mbp + g) * E::LD;  // pixel rows g, g+8 of block 2mbp; +16, +24 of block 2mbp+1
#pragma unroll
  for (int ks = 0; ks < E::KS; ++ks) {
    uint32_t a[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t w0 = r00[(16 * i) * E::LD + ks], w1 = r00[(16 * i + 8) * E::LD + ks];
      a[i][0] = w0 & m_lo; a[i][1] = w1 & m_lo; a[i][2] = w0 & m_hi; a[i][3] = w1 & m_hi;
    }
    float2 bl[2], bh[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      bl[h] = wb_lo[(ks * 2 + h) * 32 + lane];
      bh[h] = wb_hi[(ks * 2 + h) * 32 + lane];
    }
    // four independent accumulators between the lo and the hi pass of the same one (no back-to-back dependency)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) mma_tf32_16n8k8(z[i][h], a[i], __float_as_uint(bl[h].x), __float_as_uint(bl[h].y));
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) mma_tf32_16n8k8(z[i][h], a[i], __float_as_uint(bh[h].x), __float_as_uint(bh[h].y));
  }
}

// LayerNorm statistics over the 16 channels of pixel rows g (z[.][0..1]) and g+8 (z[.][2..3]); quad reduction
__device__ __forceinline__ void ln16_quad(const float (&z)[2][4], float& mean0, float& rstd0, float& mean1,
                                          float& rstd1) {
  float s0 = z[0][0] + z[0][1] + z[1][0] + z[1][1];
  float q0 = z[0][0] * z[0][0] + z[0][1] * z[0][1] + z[1][0] * z[1][0] + z[1][1] * z[1][1];
  float s1 = z[0][2] + z[0][3] + z[1][2] + z[1][3];
  float q1 = z[0][2] * z[0][2] + z[0][3] * z[0][3] + z[1][2] * z[1][2] + z[1][3] * z[1][3];
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  mean0 = s0 * (1.0f / CONV_O); mean1 = s1 * (1.0f / CONV_O);
  // MUFU.RSQ (2 ulp) instead of the IEEE 1/sqrt sequence, whose slow-path branches cost more than the conv MMAs
  rstd0 = rsqrtf(fmaxf(q0 * (1.0f / CONV_O) - mean0 * mean0, 0.f) + LN_EPS);
  rstd1 = rsqrtf(fmaxf(q1 * (1.0f / CONV_O) - mean1 * mean1, 0.f) + LN_EPS);
}

constexpr int CONV_MMA_WARPS = 8;

// H16: the activation goes out as the two fp16 planes of the fp16-split tensor-core path (H1 = hi plane, H1LO = lo'
// plane, both __half[rows][1024]) instead of fp32 h1 -- the same 4 bytes per element, so the dense GEMMs read
// (hi, lo') straight through TMA and no operand conversion happens in the GEMM kernel.
template <int C, bool TRAIN, bool H16 = false>
__global__ void __launch_bounds__(CONV_MMA_WARPS * 32, 3)
    conv_fwd_mma_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                        const float* __restrict__ params, int64_t P, pqn_net_layout_t L, float* __restrict__ H1,
                        float* __restrict__ H1LO, float* __restrict__ XH1, float* __restrict__ RS1,
                        uint32_t* __restrict__ RB, float* __restrict__ bn_sums, int rows) {
  using Cfg = ConvCfg<C>;
  using M = ConvMma<C>;
  __shared__ float2 wb_hi[M::KS * 2 * 32], wb_lo[M::KS * 2 * 32];
  __shared__ float cb[CONV_O], sc[CONV_O], bi[CONV_O];
  __shared__ uint32_t so[CONV_MMA_WARPS][Cfg::SW];
  __shared__ uint32_t sxp[CONV_MMA_WARPS][CONV_PIX * ExpPatch<C>::LD];
  __shared__ float s_cnt[C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int seed = blockIdx.y;
  conv_mma_load_weights<C, true>(params + (int64_t)seed * P, L, wb_hi, wb_lo, cb, sc, bi);
  if (TRAIN && tid < C) s_cnt[tid] = 0.f;
  __syncthreads();
  int cnt[C];
  uint32_t cmask[C];  // bits of this lane's observation word (word index = lane) that belong to channel c
#pragma unroll
  for (int c = 0; c < C; ++c) {
    cnt[c] = 0;
    cmask[c] = 0u;
    if (TRAIN) {
      for (int b = 0; b < 32; ++b) {
        const int f = lane * 32 + b;
        if (f < Cfg::OBS_BITS && f % C == c) cmask[c] |= 1u << b;
      }
    }
  }
  uint32_t* __restrict__ my_so = so[warp];
  static_assert(Cfg::PW <= 32, "one packed observation word per lane");
  if (lane == 0) my_so[Cfg::PW] = 0u;  // pad word read by the funnel shift of the last pixel
  // software pipeline: the next sample's packed observation is fetched while the current one is processed
  const int row_stride = gridDim.x * CONV_MMA_WARPS;
  auto fetch = [&](int r) -> uint32_t {
    if (r >= rows || lane >= Cfg::PW) return 0u;
    const int64_t src = gather ? gather[(int64_t)seed * rows + r] : r;
    return __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + lane);
  };
  uint32_t pre = fetch(blockIdx.x * CONV_MMA_WARPS + warp);
  for (int row = blockIdx.x * CONV_MMA_WARPS + warp; row < rows; row += row_stride) {
    __syncwarp();
    if (lane < Cfg::PW) my_so[lane] = pre;
    if (TRAIN && bn_sums != nullptr) {
      // dummy input BatchNorm statistics: per-channel popcount of the 100 input pixels (x in {0,1})
#pragma unroll
      for (int c = 0; c < C; ++c) cnt[c] += __popc(pre & cmask[c]);
    }
    __syncwarp();
    pre = fetch(row + row_stride);
    build_exp_patch<C>(my_so, lane, sxp[warp] + lane * ExpPatch<C>::LD);
    build_exp_patch<C>(my_so, lane + 32, sxp[warp] + (lane + 32) * ExpPatch<C>::LD);
    __syncwarp();
    float* __restrict__ hrow = H16 ? nullptr : H1 + ((int64_t)seed * rows + row) * FLAT_CNN;
    float* __restrict__ lrow = (!H16 && H1LO) ? H1LO + ((int64_t)seed * rows + row) * FLAT_CNN : nullptr;
    __half* __restrict__ hrow16 = H16 ? reinterpret_cast<__half*>(H1) + ((int64_t)seed * rows + row) * FLAT_CNN : nullptr;
    __half* __restrict__ lrow16 = H16 ? reinterpret_cast<__half*>(H1LO) + ((int64_t)seed * rows + row) * FLAT_CNN : nullptr;
    float* __restrict__ xrow = (TRAIN && XH1) ? XH1 + ((int64_t)seed * rows + row) * FLAT_CNN : nullptr;
    float* __restrict__ rrow = (TRAIN && RS1) ? RS1 + ((int64_t)seed * rows + row) * CONV_PIX : nullptr;
    // packed ReLU mask for the dense dgrad epilogue: bit (pixel * 16 + channel) = (h1 > 0), 16 bits per pixel
    uint16_t* __restrict__ brow =
        (TRAIN && RB) ? reinterpret_cast<uint16_t*>(RB + ((int64_t)seed * rows + row) * (FLAT_CNN / 32)) : nullptr;
#pragma unroll 1
    for (int mbp = 0; mbp < 2; ++mbp) {
     float z2[2][2][4];
     conv_mma_block2_exp<C>(sxp[warp], wb_hi, wb_lo, cb, mbp, lane, z2);
#pragma unroll
     for (int mi = 0; mi < 2; ++mi) {
      const int mb = 2 * mbp + mi;
      float (&z)[2][4] = z2[mi];
      float mean0, rstd0, mean1, rstd1;
      ln16_quad(z, mean0, rstd0, mean1, rstd1);
      uint32_t rb0 = 0u, rb1 = 0u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = 8 * h + 2 * t;
        float2 v0, v1;
        v0.x = fmaxf((z[h][0] - mean0) * rstd0 * sc[o] + bi[o], 0.f);
        v0.y = fmaxf((z[h][1] - mean0) * rstd0 * sc[o + 1] + bi[o + 1], 0.f);
        v1.x = fmaxf((z[h][2] - mean1) * rstd1 * sc[o] + bi[o], 0.f);
        v1.y = fmaxf((z[h][3] - mean1) * rstd1 * sc[o + 1] + bi[o + 1], 0.f);
        const int p0 = 16 * mb + g, p1 = p0 + 8;
        if (H16) {
          __half2 h0, l0, h1v, l1;
          // post-ReLU values are >= 0: one-sided saturation is enough
          tc::split16x2<false>(fminf(v0.x, 65000.f), fminf(v0.y, 65000.f), h0, l0);
          tc::split16x2<false>(fminf(v1.x, 65000.f), fminf(v1.y, 65000.f), h1v, l1);
          *reinterpret_cast<__half2*>(hrow16 + p0 * CONV_O + o) = h0;
          *reinterpret_cast<__half2*>(lrow16 + p0 * CONV_O + o) = l0;
          *reinterpret_cast<__half2*>(hrow16 + p1 * CONV_O + o) = h1v;
          *reinterpret_cast<__half2*>(lrow16 + p1 * CONV_O + o) = l1;
        } else {
          *reinterpret_cast<float2*>(hrow + p0 * CONV_O + o) = v0;
          *reinterpret_cast<float2*>(hrow + p1 * CONV_O + o) = v1;
        }
        if (TRAIN) {
          rb0 |= ((v0.x > 0.f ? 1u : 0u) | (v0.y > 0.f ? 2u : 0u)) << o;
          rb1 |= ((v1.x > 0.f ? 1u : 0u) | (v1.y > 0.f ? 2u : 0u)) << o;
        }
        if (xrow) {  // saved for the backward pass (no conv recompute there)
          *reinterpret_cast<float2*>(xrow + p0 * CONV_O + o) =
              make_float2((z[h][0] - mean0) * rstd0, (z[h][1] - mean0) * rstd0);
          *reinterpret_cast<float2*>(xrow + p1 * CONV_O + o) =
              make_float2((z[h][2] - mean1) * rstd1, (z[h][3] - mean1) * rstd1);
          if (h == 0 && t == 0) { rrow[p0] = rstd0; rrow[p1] = rstd1; }
        }
        if (lrow) {
          *reinterpret_cast<float2*>(lrow + p0 * CONV_O + o) = make_float2(tc::tf32_lo(v0.x), tc::tf32_lo(v0.y));
          *reinterpret_cast<float2*>(lrow + p1 * CONV_O + o) = make_float2(tc::tf32_lo(v1.x), tc::tf32_lo(v1.y));
        }
      }
      if (TRAIN && brow) {
        rb0 |= __shfl_xor_sync(0xffffffffu, rb0, 1); rb1 |= __shfl_xor_sync(0xffffffffu, rb1, 1);
        rb0 |= __shfl_xor_sync(0xffffffffu, rb0, 2); rb1 |= __shfl_xor_sync(0xffffffffu, rb1, 2);
        if (t == 0) { brow[16 * mb + g] = (uint16_t)rb0; brow[16 * mb + g + 8] = (uint16_t)rb1; }
      }
     }
    }
  }
  if (TRAIN && bn_sums != nullptr) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int v = cnt[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(&s_cnt[c], (float)v);
    }
    __syncthreads();
    if (tid < C && s_cnt[tid] != 0.f) {
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + tid, s_cnt[tid]);
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + C + tid, s_cnt[tid]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// conv forward on fp16 warp-level MMA (mma.sync.m16n8k16, fp32 accumulate) -- the default conv path of round 2.
// Same structure as conv_fwd_mma_kernel (one sample per warp, quad-level LayerNorm), but
//   * the {0,1} im2col operand is fp16 and "exponent coded": per output pixel and k-step of 16 taps two words hold the
//     tap bits at the exponent bits 10..13 of the low half and 26..29 of the high half; lane t of the fragment masks
//     bit (10 + t) / (26 + t), which turns a set bit into the fp16 power of two 2^(2^t - 15) (one LOP3 per register
//     that carries TWO k values) and row k of B is pre-multiplied by the inverse power of two (exact), so every product
//     equals the plain 0/1 product;
//   * weights/255 are split into fp16 hi + lo (22 significant bits, like the tf32 hi/lo pair);
//   * K = 9C taps padded to 16 needs ceil(9C/16) k-steps (3 for C = 4) instead of ceil(9C/8) = 5 tf32 ones: 48 MMAs
//     and 48 fragment LOP3s per sample instead of 80 / 80 (the tf32 kernel's top stall was the mma.sync pipe).
// k order inside a k-step (free to choose, B is laid out to match): fragment column 2t <-> tap 16s + t,
// 2t+1 <-> 16s + 4 + t, 2t+8 <-> 16s + 8 + t, 2t+9 <-> 16s + 12 + t.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16_16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int C>
struct Conv16 {
  static constexpr int TAPS = 9 * C;
  static constexpr int KS = (TAPS + 15) / 16;          // k-steps of 16 taps
  static constexpr int ROW = 2 * KS < 8 ? 8 : 2 * KS;  // padded row: 8 or 12 words keep the 4-row LDS.64 groups apart
};

// Output channel of column n (0..7) of n-tile h.  NOT the natural 8h + n: with 4 (n / 2) + 2h + (n % 2) the accumulator
// columns (2t, 2t+1) of the two n-tiles are the four CONSECUTIVE channels 4t .. 4t+3 of a pixel, so a thread stores 16
// bytes of xhat / 8 bytes of each fp16 plane per pixel with one instruction and no lane exchange (the kernel's time
// follows its store instructions: same-box A/B in profiles/r2_conv_fwd_sensitivity.json).
__host__ __device__ constexpr int conv16_channel(int h, int n) { return 4 * (n >> 1) + 2 * h + (n & 1); }

// tap of fragment column kk (0..15) of k-step s, see the k order above
__host__ __device__ constexpr int conv16_tap(int s, int kk) {
  return 16 * s + (kk < 8 ? 0 : 8) + ((kk & 1) ? 4 : 0) + ((kk & 7) >> 1);
}

// B fragments of weights/255 as fp16 (hi, lo), pre-scaled by the inverse of the A coding: wb[s][h][lane] = uint4
// {b0_hi, b1_hi, b0_lo, b1_lo} (b0 = columns k = 2t, 2t+1; b1 = k = 2t+8, 2t+9; n = 8h + g)
template <int C>
__device__ __forceinline__ void conv16_load_weights(const float* __restrict__ prm, const pqn_net_layout_t& L, uint4* wb,
                                                    float* cb, float* sc, float* bi) {
  using M = Conv16<C>;
  const float inv255 = 1.0f / 255.0f;
  for (int i = threadIdx.x; i < M::KS * 2 * 32; i += blockDim.x) {
    const int ln = i & 31, h = (i >> 5) & 1, s = i >> 6;
    const int gg = ln >> 2, tt = ln & 3;
    const int o = conv16_channel(h, gg);
    const float scale = __uint_as_float((uint32_t)(127 + 15 - (1 << tt)) << 23);   // 2^(15 - 2^t), exact
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // q: 0 -> k=2t, 1 -> 2t+1, 2 -> 2t+8, 3 -> 2t+9
      const int tap = conv16_tap(s, 2 * tt + (q & 1) + (q >> 1) * 8);
      v[q] = tap < M::TAPS ? __ldg(prm + L.conv_w + tap * CONV_O + o) * inv255 * scale : 0.f;
      v[q] = fminf(fmaxf(v[q], -65000.f), 65000.f);
    }
    const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
    wb[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                       *reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
  if (threadIdx.x < CONV_O) {
    cb[threadIdx.x] = __ldg(prm + L.conv_b + threadIdx.x);
    sc[threadIdx.x] = __ldg(prm + L.ln0_scale + threadIdx.x);
    bi[threadIdx.x] = __ldg(prm + L.ln0_bias + threadIdx.x);
  }
}

// exponent-coded fp16 patch words of one output pixel: out[2s], out[2s+1] for k-step s (taps 16s..16s+7, 16s+8..16s+15).
// Only bits 10..13 and 26..29 are meaningful (the fragment mask picks one of them per half); tap 16s+j, j<4 sits at bit
// 10+j and tap 16s+4+j at bit 26+j.
template <int C>
__device__ __forceinline__ void build_patch16(const uint32_t* __restrict__ so, int pix, uint32_t* __restrict__ out) {
  constexpr int W = PatchCfg<C>::WORDS;
  uint32_t w[W];
  patch_bits<C>(so, pix, w);
#pragma unroll
  for (int b = 0; b < 2 * Conv16<C>::KS; ++b) {
    const int o = 8 * b;                                   // bit offset of this byte in the patch string
    uint32_t v = (o >> 5) < W ? w[(o >> 5) < W ? (o >> 5) : 0] : 0u;
    const int sh = o & 31;
    const uint32_t lo = sh >= 10 ? v >> (sh - 10) : v << (10 - sh);          // bits sh..sh+3   -> 10..13
    const uint32_t hi = sh + 4 <= 26 ? v << (26 - sh - 4) : v >> (sh + 4 - 26);  // bits sh+4..sh+7 -> 26..29
    out[b] = __byte_perm(lo, hi, 0x7610);                  // low half from lo, high half from hi
  }
}

// patch row of one pixel -> shared memory with 128-bit stores (ROW is 8 or 12 words; pad words are never read)
template <int C>
__device__ __forceinline__ void store_patch16(const uint32_t* __restrict__ so, int pix, uint32_t* __restrict__ row) {
  using M = Conv16<C>;
  uint32_t w[M::ROW];
#pragma unroll
  for (int k = 0; k < M::ROW; ++k) w[k] = 0u;
  build_patch16<C>(so, pix, w);
#pragma unroll
  for (int q = 0; q < M::ROW / 4; ++q)
    *reinterpret_cast<uint4*>(row + 4 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

template <int C>
__device__ __forceinline__ void conv16_block2(const uint32_t* __restrict__ xp, const uint4* __restrict__ wb,
                                              const float* __restrict__ cb, int mbp, int lane, float (&z)[2][2][4]) {
  using M = Conv16<C>;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t mask = (1u << (10 + t)) | (1u << (26 + t));
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      z[i][h][0] = z[i][h][2] = cb[4 * t + 2 * h];       // conv16_channel(h, 2t), (h, 2t + 1)
      z[i][h][1] = z[i][h][3] = cb[4 * t + 2 * h + 1];
    }
  const uint32_t* r00 = xp + (32 * mbp + g) * M::ROW;
#pragma unroll
  for (int s = 0; s < M::KS; ++s) {
    uint32_t a[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint2 w0 = *reinterpret_cast<const uint2*>(r00 + (16 * i) * M::ROW + 2 * s);       // pixel row g
      const uint2 w1 = *reinterpret_cast<const uint2*>(r00 + (16 * i + 8) * M::ROW + 2 * s);   // pixel row g + 8
      a[i][0] = w0.x & mask; a[i][1] = w1.x & mask; a[i][2] = w0.y & mask; a[i][3] = w1.y & mask;
    }
    uint4 b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) b[h] = wb[(s * 2 + h) * 32 + lane];
    // lo pass of all four accumulators, then the hi pass: no back-to-back dependent MMAs
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) mma_f16_16n8k16(z[i][h], a[i], b[h].z, b[h].w);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 2; ++i) mma_f16_16n8k16(z[i][h], a[i], b[h].x, b[h].y);
  }
}

__device__ __forceinline__ uint32_t cvt_f16x2_satfinite(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 4 warps per CTA, 5 CTAs per SM = 5 warps per scheduler: 96 registers per thread.  (With 8-warp CTAs x 3 the cap is 80
// registers and the epilogue spilled: ncu r2e, STL = 0.7 % of the instructions but 13 % of the stall samples; the
// kernel wants ~123 registers unconstrained.)
constexpr int CONV16_WARPS = 4;
constexpr int CONV16_CTAS_PER_SM = 5;

template <int C, bool TRAIN, bool H16>
__global__ void __launch_bounds__(CONV16_WARPS * 32, CONV16_CTAS_PER_SM)
    conv_fwd_mma16_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                          const float* __restrict__ params, int64_t P, pqn_net_layout_t L, float* __restrict__ H1,
                          float* __restrict__ H1LO, float* __restrict__ XH1, float* __restrict__ RS1,
                          uint32_t* __restrict__ RB, float* __restrict__ bn_sums, int rows) {
  using Cfg = ConvCfg<C>;
  using M = Conv16<C>;
  __shared__ __align__(16) uint4 wb[M::KS * 2 * 32];
  __shared__ float cb[CONV_O], sc[CONV_O], bi[CONV_O];
  __shared__ uint32_t so[CONV16_WARPS][Cfg::SW];
  __shared__ __align__(16) uint32_t sxp[CONV16_WARPS][CONV_PIX * M::ROW];
  __shared__ float s_cnt[C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int seed = blockIdx.y;
  conv16_load_weights<C>(params + (int64_t)seed * P, L, wb, cb, sc, bi);
  if (TRAIN && tid < C) s_cnt[tid] = 0.f;
  __syncthreads();
  int cnt[C];
  uint32_t cmask[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    cnt[c] = 0;
    cmask[c] = 0u;
    if (TRAIN) {
      for (int b = 0; b < 32; ++b) {
        const int f = lane * 32 + b;
        if (f < Cfg::OBS_BITS && f % C == c) cmask[c] |= 1u << b;
      }
    }
  }
  uint32_t* __restrict__ my_so = so[warp];
  static_assert(Cfg::PW <= 32, "one packed observation word per lane");
  if (lane == 0) my_so[Cfg::PW] = 0u;
  const int row_stride = gridDim.x * CONV16_WARPS;
  // two-deep prefetch: the gather index of row + 2 strides and the packed observation of row + 1 stride are in flight
  // while this row is computed, so the index -> observation load chain never stalls the in-order issue
  // unconditional loads from clamped addresses (rows past the end re-read the last row, lanes past the packed width
  // re-read word 0; neither is ever used): a predicated load with a default value made ptxas copy the result into the
  // loop-carried register ~125 instructions after the LDG, which stalled every warp on the load it had just issued
  // (13 % of all stall samples of the forward kernel, ncu r2m)
  auto fetch_index = [&](int r) -> int {
    const int rc = r < rows ? r : rows - 1;
    return gather ? __ldg(gather + (int64_t)seed * rows + rc) : rc;
  };
  auto fetch_obs = [&](int src) -> uint32_t {
    return __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + (lane < Cfg::PW ? lane : 0));
  };
  const int row0 = blockIdx.x * CONV16_WARPS + warp;
  uint32_t pre = fetch_obs(fetch_index(row0));
  int src_next = fetch_index(row0 + row_stride);
  for (int row = row0; row < rows; row += row_stride) {
    __syncwarp();
    if (lane < Cfg::PW) my_so[lane] = pre;
    if (TRAIN && bn_sums != nullptr) {
#pragma unroll
      for (int c = 0; c < C; ++c) cnt[c] += __popc(pre & cmask[c]);
    }
    __syncwarp();
    pre = fetch_obs(src_next);
    src_next = fetch_index(row + 2 * row_stride);
    store_patch16<C>(my_so, lane, sxp[warp] + lane * M::ROW);
    store_patch16<C>(my_so, lane + 32, sxp[warp] + (lane + 32) * M::ROW);
    __syncwarp();
    const int64_t grow = (int64_t)seed * rows + row;
    float* __restrict__ hrow = H16 ? nullptr : H1 + grow * FLAT_CNN;
    __half* __restrict__ hrow16 = H16 ? reinterpret_cast<__half*>(H1) + grow * FLAT_CNN : nullptr;
    __half* __restrict__ lrow16 = H16 ? reinterpret_cast<__half*>(H1LO) + grow * FLAT_CNN : nullptr;
    float* __restrict__ xrow = (TRAIN && XH1) ? XH1 + grow * FLAT_CNN : nullptr;
    float* __restrict__ rrow = (TRAIN && RS1) ? RS1 + grow * CONV_PIX : nullptr;
    uint16_t* __restrict__ brow = (TRAIN && RB) ? reinterpret_cast<uint16_t*>(RB + grow * (FLAT_CNN / 32)) : nullptr;
    // rstd and the ReLU masks of m-block t are kept by lane t of every quad and stored once per sample (2 + 2 store
    // instructions instead of 8 + 8 predicated ones)
    float keep_rs0 = 0.f, keep_rs1 = 0.f;
    uint32_t keep_rb = 0u;
#pragma unroll 1
    for (int mbp = 0; mbp < 2; ++mbp) {
      float z2[2][2][4];
      conv16_block2<C>(sxp[warp], wb, cb, mbp, lane, z2);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mb = 2 * mbp + mi;
        float (&z)[2][4] = z2[mi];
        float mean0, rstd0, mean1, rstd1;
        ln16_quad(z, mean0, rstd0, mean1, rstd1);
        const float nm0 = -mean0 * rstd0, nm1 = -mean1 * rstd1;   // xhat = z * rstd - mean * rstd: one FFMA
        uint32_t rb0 = 0u, rb1 = 0u;
        // the thread's channels of pixel rows p0 = 16 mb + g and p1 = p0 + 8: 4t .. 4t+3 (n-tile h -> 4t + 2h, + 1)
        const int p0 = 16 * mb + g, p1 = p0 + 8, o4 = 4 * t;
        float x0[4], x1[4], v0[4], v1[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int o = o4 + 2 * h + c;
            x0[2 * h + c] = fmaf(z[h][c], rstd0, nm0);
            x1[2 * h + c] = fmaf(z[h][2 + c], rstd1, nm1);
            v0[2 * h + c] = fmaxf(fmaf(x0[2 * h + c], sc[o], bi[o]), 0.f);
            v1[2 * h + c] = fmaxf(fmaf(x1[2 * h + c], sc[o], bi[o]), 0.f);
          }
        }
        if (H16) {
          // hi = fp16(h) (saturating: no inf), lo = fp16((h - hi) * 2^11); 8-byte stores
          uint32_t hw0[2], hw1[2], lw0[2], lw1[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            hw0[h] = cvt_f16x2_satfinite(v0[2 * h], v0[2 * h + 1]);
            hw1[h] = cvt_f16x2_satfinite(v1[2 * h], v1[2 * h + 1]);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hw0[h]));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&hw1[h]));
            const __half2 l0 = __floats2half2_rn((v0[2 * h] - f0.x) * tc::TC_LO_SCALE, (v0[2 * h + 1] - f0.y) * tc::TC_LO_SCALE);
            const __half2 l1 = __floats2half2_rn((v1[2 * h] - f1.x) * tc::TC_LO_SCALE, (v1[2 * h + 1] - f1.y) * tc::TC_LO_SCALE);
            lw0[h] = *reinterpret_cast<const uint32_t*>(&l0); lw1[h] = *reinterpret_cast<const uint32_t*>(&l1);
          }
          *reinterpret_cast<uint2*>(hrow16 + p0 * CONV_O + o4) = make_uint2(hw0[0], hw0[1]);
          *reinterpret_cast<uint2*>(lrow16 + p0 * CONV_O + o4) = make_uint2(lw0[0], lw0[1]);
          *reinterpret_cast<uint2*>(hrow16 + p1 * CONV_O + o4) = make_uint2(hw1[0], hw1[1]);
          *reinterpret_cast<uint2*>(lrow16 + p1 * CONV_O + o4) = make_uint2(lw1[0], lw1[1]);
        } else {
          *reinterpret_cast<float4*>(hrow + p0 * CONV_O + o4) = make_float4(v0[0], v0[1], v0[2], v0[3]);
          *reinterpret_cast<float4*>(hrow + p1 * CONV_O + o4) = make_float4(v1[0], v1[1], v1[2], v1[3]);
        }
        if (TRAIN) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            rb0 |= (v0[c] > 0.f ? 1u : 0u) << (o4 + c);
            rb1 |= (v1[c] > 0.f ? 1u : 0u) << (o4 + c);
          }
        }
        if (xrow) {
          *reinterpret_cast<float4*>(xrow + p0 * CONV_O + o4) = make_float4(x0[0], x0[1], x0[2], x0[3]);
          *reinterpret_cast<float4*>(xrow + p1 * CONV_O + o4) = make_float4(x1[0], x1[1], x1[2], x1[3]);
          if (t == mb) { keep_rs0 = rstd0; keep_rs1 = rstd1; }
        }
        if (TRAIN && brow) {
          uint32_t rb = rb0 | (rb1 << 16);      // both pixel rows in one register: two shuffles instead of four
          rb |= __shfl_xor_sync(0xffffffffu, rb, 1);
          rb |= __shfl_xor_sync(0xffffffffu, rb, 2);
          if (t == mb) keep_rb = rb;
        }
      }
    }
    if (TRAIN) {
      if (rrow) { rrow[16 * t + g] = keep_rs0; rrow[16 * t + g + 8] = keep_rs1; }
      if (brow) { brow[16 * t + g] = (uint16_t)keep_rb; brow[16 * t + g + 8] = (uint16_t)(keep_rb >> 16); }
    }
  }
  if (TRAIN && bn_sums != nullptr) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int v = cnt[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(&s_cnt[c], (float)v);
    }
    __syncthreads();
    if (tid < C && s_cnt[tid] != 0.f) {
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + tid, s_cnt[tid]);
      atomicAdd(bn_sums + (int64_t)seed * 2 * C + C + tid, s_cnt[tid]);
    }
  }
}

constexpr int CDZ_LD = 16;  // staged dz / dy / xhat row stride (floats)
// Column swizzle of the [64 pixels][16 channels] shared-memory tiles of the conv backward: rows p and p+2 are 32 banks
// apart, so the column is XORed with 8 on every second row pair.  The float2 accesses of a fragment (8 pixel rows x
// 4 column pairs per half-warp pair) and the B-fragment reads (4 pixel rows x 8 columns) then touch every bank once.
__device__ __forceinline__ int cswz(int p, int col) { return col ^ (((p >> 1) & 1) << 3); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// per-warp shared-memory slice of the conv backward kernel (floats)
template <int C>
struct ConvBwdSmem {
  static constexpr int DY = 0;                                  // [64][16] upstream gradient of the sample
  static constexpr int XH = DY + FLAT_CNN;                      // [64][16] saved LayerNorm xhat
  static constexpr int RS = XH + FLAT_CNN;                      // [64]     saved rstd
  static constexpr int DZ = RS + CONV_PIX;                      // [64][CDZ_LD] dz stage (aliases the packed obs row)
  static constexpr int PATCH = DZ + CONV_PIX * CDZ_LD;          // [64][PatchCfg::WORDS] im2col bit patches
  static constexpr int WARP_FLOATS = (PATCH + CONV_PIX * PatchCfg<C>::WORDS + 3) / 4 * 4;
  static constexpr int WARPS = 8;
  static constexpr int BYTES = (WARPS * WARP_FLOATS + 4 * CONV_O) * 4;  // + sc[16] + s_red[48]
};

// Backward of conv3x3 + LayerNorm(16) + ReLU for one sample per warp, from the xhat / rstd the training forward
// saved (ReLU' is already folded into DY1 by the dense dgrad epilogue):
//   phase A  LayerNorm backward per pixel -> dz[pixel][16] (staged in shared memory), d(scale), d(bias), d(conv bias)
//   phase B  dW[tap][o] += sum_pixels x[pixel, tap] dz[pixel][o] on mma.sync (A = im2col bits, B = dz as hi + lo)
// The sample's DY1 / xhat / rstd rows (8.25 KB) are fetched with cp.async while phase B of the previous sample
// runs, so the HBM latency is off the dependent path (it was 42% of all stall samples before).
template <int C>
__global__ void __launch_bounds__(ConvBwdSmem<C>::WARPS * 32, 2)
    conv_bwd_mma_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                        const float* __restrict__ params, int64_t P, pqn_net_layout_t L, const float* __restrict__ DY1,
                        const float* __restrict__ XH1, const float* __restrict__ RS1, float* __restrict__ part,
                        int rows) {
  using Cfg = ConvCfg<C>;
  using M = ConvMma<C>;
  using SM = ConvBwdSmem<C>;
  constexpr int PWD = PatchCfg<C>::WORDS;
  extern __shared__ __align__(16) float smem_bwd[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int seed = blockIdx.y;
  float* my = smem_bwd + warp * SM::WARP_FLOATS;
  float* my_dy = my + SM::DY;
  float* my_xh = my + SM::XH;
  float* my_rs = my + SM::RS;
  float* my_dz = my + SM::DZ;
  uint32_t* my_so = reinterpret_cast<uint32_t*>(my + SM::DZ);  // only needed until the patches are built
  uint32_t* my_patch = reinterpret_cast<uint32_t*>(my + SM::PATCH);
  float* sc = smem_bwd + SM::WARPS * SM::WARP_FLOATS;
  float* s_red = sc + CONV_O;
  float* s_w = smem_bwd;  // block-level dW reduction buffer, aliases warp 0's slice; only used after the row loop
  static_assert(Cfg::SW <= CONV_PIX * CDZ_LD, "obs staging aliases the dz stage");
  static_assert(M::MT * 16 * CONV_O <= SM::WARP_FLOATS * SM::WARPS, "dW reduction buffer aliases the warp slices");
  static_assert(Cfg::PW <= 32, "one packed observation word per lane");
  if (tid < CONV_O) sc[tid] = __ldg(params + (int64_t)seed * P + L.ln0_scale + tid);
  if (tid < 3 * CONV_O) s_red[tid] = 0.f;
  __syncthreads();
  // lane-private accumulators: columns {2t, 2t+1, 8+2t, 8+2t+1}
  float a_dsc[4] = {0.f, 0.f, 0.f, 0.f}, a_dbi[4] = {0.f, 0.f, 0.f, 0.f}, a_dcb[4] = {0.f, 0.f, 0.f, 0.f};
  float wrun[M::MT][2][4];
#pragma unroll
  for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) wrun[mt][h][j] = 0.f;

  const int row_stride = gridDim.x * SM::WARPS;
  auto fetch_obs = [&](int r) -> uint32_t {
    if (r >= rows || lane >= Cfg::PW) return 0u;
    const int64_t src = gather ? gather[(int64_t)seed * rows + r] : r;
    return __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + lane);
  };
  auto fetch_rows = [&](int r) {  // async copy of the sample's dy / xhat / rstd rows into this warp's slice
    if (r < rows) {
      const int64_t gr = (int64_t)seed * rows + r;
      const float* dsrc = DY1 + gr * FLAT_CNN;
      const float* xsrc = XH1 + gr * FLAT_CNN;
#pragma unroll
      for (int i = 0; i < FLAT_CNN / 4 / 32; ++i) {
        const int q = i * 32 + lane, prow = q >> 2;              // 16-byte chunk q = (pixel row, column quad)
        const int dst = prow * CONV_O + cswz(prow, (q & 3) * 4);  // the XOR moves whole chunks
        cp_async16(my_dy + dst, dsrc + q * 4);
        cp_async16(my_xh + dst, xsrc + q * 4);
      }
      if (lane < CONV_PIX / 4) cp_async16(my_rs + lane * 4, RS1 + gr * CONV_PIX + lane * 4);
    }
    cp_async_commit();
  };
  int row = blockIdx.x * SM::WARPS + warp;
  uint32_t pre = fetch_obs(row);
  fetch_rows(row);
  for (; row < rows; row += row_stride) {
    __syncwarp();
    if (lane < Cfg::PW) my_so[lane] = pre;
    if (lane == 0) my_so[Cfg::PW] = 0u;  // pad word read by the funnel shift of the last pixel
    __syncwarp();
    pre = fetch_obs(row + row_stride);
    build_patches<C>(my_so, my_patch, lane);
    cp_async_wait_all();
    __syncwarp();
    // ---- phase A: LayerNorm backward, stage dz
#pragma unroll 1
    for (int mb = 0; mb < 4; ++mb) {
      const int p0 = 16 * mb + g, p1 = p0 + 8;
      float z[2][4];  // xhat in C-fragment layout
      float2 dyv[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        dyv[h][0] = *reinterpret_cast<const float2*>(my_dy + p0 * CONV_O + cswz(p0, 8 * h + 2 * t));
        dyv[h][1] = *reinterpret_cast<const float2*>(my_dy + p1 * CONV_O + cswz(p1, 8 * h + 2 * t));
        const float2 a0 = *reinterpret_cast<const float2*>(my_xh + p0 * CONV_O + cswz(p0, 8 * h + 2 * t));
        const float2 a1 = *reinterpret_cast<const float2*>(my_xh + p1 * CONV_O + cswz(p1, 8 * h + 2 * t));
        z[h][0] = a0.x; z[h][1] = a0.y; z[h][2] = a1.x; z[h][3] = a1.y;
      }
      const float rstd0 = my_rs[p0], rstd1 = my_rs[p1];
      float dxh[2][4];
      float m1a = 0.f, m2a = 0.f, m1b = 0.f, m2b = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = 8 * h + 2 * t;
        const float dy4[4] = {dyv[h][0].x, dyv[h][0].y, dyv[h][1].x, dyv[h][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 2 * h + (j & 1);    // index into the lane's 4 columns; z[h][j] holds xhat
          a_dsc[col] = fmaf(dy4[j], z[h][j], a_dsc[col]);
          a_dbi[col] += dy4[j];
          dxh[h][j] = dy4[j] * sc[o + (j & 1)];
          if (j < 2) { m1a += dxh[h][j]; m2a = fmaf(dxh[h][j], z[h][j], m2a); }
          else { m1b += dxh[h][j]; m2b = fmaf(dxh[h][j], z[h][j], m2b); }
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        m1a += __shfl_xor_sync(0xffffffffu, m1a, o); m2a += __shfl_xor_sync(0xffffffffu, m2a, o);
        m1b += __shfl_xor_sync(0xffffffffu, m1b, o); m2b += __shfl_xor_sync(0xffffffffu, m2b, o);
      }
      m1a *= (1.0f / CONV_O); m2a *= (1.0f / CONV_O); m1b *= (1.0f / CONV_O); m2b *= (1.0f / CONV_O);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = 8 * h + 2 * t;
        float dzv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float rstd = j < 2 ? rstd0 : rstd1, m1 = j < 2 ? m1a : m1b, m2 = j < 2 ? m2a : m2b;
          dzv[j] = rstd * (dxh[h][j] - m1 - z[h][j] * m2);
          a_dcb[2 * h + (j & 1)] += dzv[j];
        }
        *reinterpret_cast<float2*>(my_dz + p0 * CDZ_LD + cswz(p0, o)) = make_float2(dzv[0], dzv[1]);
        *reinterpret_cast<float2*>(my_dz + p1 * CDZ_LD + cswz(p1, o)) = make_float2(dzv[2], dzv[3]);
      }
    }
    __syncwarp();
    fetch_rows(row + row_stride);  // overlaps phase B
    // ---- phase B: dW[tap][o] += sum_pixels x[pixel, tap] * dz[pixel][o]   (fresh accumulators per sample)
    float wacc[M::MT][2][4];
#pragma unroll
    for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) wacc[mt][h][j] = 0.f;
#pragma unroll 1
    for (int kk = 0; kk < 8; ++kk) {
      // B fragments: b0 = dz[pixel 8kk + t][o = 8h + g], b1 = dz[pixel 8kk + t + 4][o]
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r0 = 8 * kk + t, r1 = r0 + 4;
        const float v0 = my_dz[r0 * CDZ_LD + cswz(r0, 8 * h + g)], v1 = my_dz[r1 * CDZ_LD + cswz(r1, 8 * h + g)];
        bh[h][0] = __float_as_uint(v0) & 0xFFFFE000u; bh[h][1] = __float_as_uint(v1) & 0xFFFFE000u;
        bl[h][0] = __float_as_uint(v0 - __uint_as_float(bh[h][0])); bl[h][1] = __float_as_uint(v1 - __uint_as_float(bh[h][1]));
      }
      // A fragments: x[tap, pixel]: a0 = (tap g, pixel 8kk+t), a1 = (tap g+8, same), a2 = (tap g, pixel+4), a3
      const int pa = 8 * kk + t, pb = pa + 4;
      uint32_t qa[PWD], qb[PWD];
#pragma unroll
      for (int k = 0; k < PWD; ++k) { qa[k] = my_patch[pa * PWD + k]; qb[k] = my_patch[pb * PWD + k]; }
#pragma unroll
      for (int mt = 0; mt < M::MT; ++mt) {
        // taps 16mt .. 16mt+15 live in one patch word (32 % 16 == 0, 16 mt < 9C <= 32 PWD); bits beyond 9C are zero
        const uint32_t wa = qa[(16 * mt) >> 5], wb = qb[(16 * mt) >> 5];
        const int sh = ((16 * mt) & 31) + g;
        uint32_t a[4];
        a[0] = bit_f32(wa, sh); a[1] = bit_f32(wa, sh + 8); a[2] = bit_f32(wb, sh); a[3] = bit_f32(wb, sh + 8);
        if (__ballot_sync(0xffffffffu, (a[0] | a[1] | a[2] | a[3]) != 0u) == 0u) continue;
        // lo pass of both column halves, then the hi pass: no back-to-back MMAs on one accumulator
        mma_tf32_16n8k8(wacc[mt][0], a, bl[0][0], bl[0][1]);
        mma_tf32_16n8k8(wacc[mt][1], a, bl[1][0], bl[1][1]);
        mma_tf32_16n8k8(wacc[mt][0], a, bh[0][0], bh[0][1]);
        mma_tf32_16n8k8(wacc[mt][1], a, bh[1][0], bh[1][1]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) wrun[mt][h][j] += wacc[mt][h][j];
  }
  cp_async_wait_all();
  // ---- reduce and publish (deterministic): wrun[mt][h][j] is dW[tap = 16mt + g (+8 for j>=2)][o = 8h + 2t + (j&1)].
  // The 8 warps add their registers to the block accumulator one after the other, lanes of a warp own distinct
  // elements; the CTA's partial vector goes to part[seed][cta][TAPS*16 + 48] and conv_bwd_final_kernel sums the CTAs
  // in index order.
  __syncthreads();  // all warps are done with their slices
  for (int w = 0; w < SM::WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tap = 16 * mt + g + (j >= 2 ? 8 : 0);
            float* dst = &s_w[tap * CONV_O + 8 * h + 2 * t + (j & 1)];   // tap < 16 * MT: inside the buffer
            *dst = (w == 0 ? 0.f : *dst) + wrun[mt][h][j];
          }
    }
    __syncthreads();
  }
  // per-channel sums: lanes with the same t hold the same columns -> xor-shuffle tree (fixed), then warps in order
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    float v0 = a_dsc[col], v1 = a_dbi[col], v2 = a_dcb[col];
#pragma unroll
    for (int sft = 4; sft <= 16; sft <<= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, sft);
      v1 += __shfl_xor_sync(0xffffffffu, v1, sft);
      v2 += __shfl_xor_sync(0xffffffffu, v2, sft);
    }
    a_dsc[col] = v0; a_dbi[col] = v1; a_dcb[col] = v2;
  }
  for (int w = 0; w < SM::WARPS; ++w) {
    if (warp == w && g == 0) {
#pragma unroll
      for (int col = 0; col < 4; ++col) {
        const int o = 8 * (col >> 1) + 2 * t + (col & 1);
        s_red[o] = (w == 0 ? 0.f : s_red[o]) + a_dsc[col];
        s_red[CONV_O + o] = (w == 0 ? 0.f : s_red[CONV_O + o]) + a_dbi[col];
        s_red[2 * CONV_O + o] = (w == 0 ? 0.f : s_red[2 * CONV_O + o]) + a_dcb[col];
      }
    }
    __syncthreads();
  }
  float* __restrict__ o = part + ((int64_t)seed * gridDim.x + blockIdx.x) * (M::TAPS * CONV_O + 3 * CONV_O);
  // x = bit / 255 (pqn_minatar.py:66): the A operand was the raw bit, so scale here
  for (int i = tid; i < M::TAPS * CONV_O; i += blockDim.x) o[i] = s_w[i] * (1.0f / 255.0f);
  if (tid < 3 * CONV_O) o[M::TAPS * CONV_O + tid] = s_red[tid];
}

// ---------------------------------------------------------------------------------------------------------------
// conv backward with the weight gradient on fp16 warp-level MMA (mma.sync.m16n8k16, fp32 accumulate) -- the default
// backward of the fp16 conv path.  Same phases as conv_bwd_mma_kernel; what changes is phase B,
// dW[tap][o] = sum_pixels x[pixel, tap] dz[pixel][o] with M = taps, N = 16 channels, K = 64 pixels in 4 k-steps of 16:
//   * A (the {0,1} im2col bits, tap-major) is "exponent coded" like the forward: for every PAIR of pixels (2j, 2j+1)
//     and group q of four taps one word holds the tap bits of pixel 2j at bits 10..13 and of pixel 2j+1 at bits 26..29;
//     a fragment register (row = tap, its two k values = the two pixels of a pair) is that word AND the lane's mask
//     (1 << (10 + g%4)) | (1 << (26 + g%4)): a set bit becomes the fp16 power of two 2^(2^(g%4) - 15).  The factor
//     depends only on the accumulator ROW, which is fixed per thread, so it is undone by one exact multiplication
//     when the accumulators are published -- 48 LOP3 per sample instead of 288 shift / and / select.
//   * B = dz * gs as fp16 hi + lo (lo = fp16(v - hi): 22 significant bits where it matters; gs is a power of two that
//     lifts the gradient-sized values out of the fp16 subnormals), written by phase A directly in fragment order:
//     dzT[plane][channel][pixel pair], the pairs (t, t + 4) of a k-step adjacent so that (b0, b1) is one LDS.64; the
//     k-step block of a channel row is XOR-swizzled so that the phase-A stores and the fragment loads are conflict free.
//   * 48 MMAs per sample (3 m-tiles x 2 n-tiles x 4 k-steps x {hi, lo}) instead of 96 tf32 ones.
// Phase A works on the pixel pair (16 mb + 2g, + 1) per thread (it was (g, g + 8)): a thread's two dz values of one
// channel are exactly one B word.  The DY / xhat stage has a row + column swizzle for that access pattern.
// ---------------------------------------------------------------------------------------------------------------
template <int C>
struct ConvBwd16 {
  static constexpr int MT = ConvMma<C>::MT;
  static constexpr int NG = 4 * MT;                             // tap groups of 4 per pixel pair
  static constexpr int NGP = (NG % 8 == 4) ? NG : NG + 4;       // pair stride in words: the 4 t-lanes hit distinct banks
  static constexpr int DY = 0;                                  // [64][16] upstream gradient (swizzled)
  static constexpr int XH = DY + FLAT_CNN;                      // [64][16] saved xhat (swizzled)
  static constexpr int RS = XH + FLAT_CNN;                      // [64] saved rstd
  static constexpr int DZT = RS + CONV_PIX;                     // 2 planes x [16 ch][32 pair words]; aliases the obs row
  static constexpr int PW = DZT + 2 * CONV_O * 32;              // [32 pairs][NGP] exponent-coded pair words
  static constexpr int WARP_FLOATS = (PW + 32 * NGP + 3) / 4 * 4;
  static constexpr int WARPS = (C == 4) ? 8 : 6;                // keeps two CTAs per SM for the wider observations
  static constexpr int BYTES = (WARPS * WARP_FLOATS + 4 * CONV_O) * 4;  // + sc[16] + s_red[48]
};

// float offset of (pixel row p, channel column col) in the swizzled [64][16] DY / xhat stage: the rows 2g of one
// fragment group would all start on bank 0, so odd (p / 4) swaps the two rows of a pair and odd (p / 2) swaps the
// column halves -- the four g of a half-warp then cover the 32 banks once
__device__ __forceinline__ int cswz16(int p, int col) {
  return ((p ^ ((p >> 2) & 1)) << 4) + (col ^ (((p >> 1) & 1) << 3));
}
// word offset of (channel ch, k-step ks, position pos) in a dzT plane [16][32]
__device__ __forceinline__ int dzt_word(int ch, int ks, int pos) {
  const int c7 = ch & 7;
  return (ch << 5) + ((ks ^ ((c7 + (c7 >> 2)) & 3)) << 3) + pos;
}

template <int C>
__global__ void __launch_bounds__(ConvBwd16<C>::WARPS * 32, 2)
    conv_bwd_mma16_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                          const float* __restrict__ params, int64_t P, pqn_net_layout_t L, const float* __restrict__ DY1,
                          const float* __restrict__ XH1, const float* __restrict__ RS1, float* __restrict__ part,
                          int rows, float gs) {
  using Cfg = ConvCfg<C>;
  using M = ConvMma<C>;
  using SM = ConvBwd16<C>;
  constexpr int PWD = PatchCfg<C>::WORDS;
  extern __shared__ __align__(16) float smem_bwd[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int seed = blockIdx.y;
  float* my = smem_bwd + warp * SM::WARP_FLOATS;
  float* my_dy = my + SM::DY;
  float* my_xh = my + SM::XH;
  float* my_rs = my + SM::RS;
  uint32_t* my_dzt = reinterpret_cast<uint32_t*>(my + SM::DZT);   // plane 0 = hi, plane 1 = lo: 512 words each
  uint32_t* my_so = reinterpret_cast<uint32_t*>(my + SM::DZT);    // packed obs row: only needed until the pair words exist
  uint32_t* my_pw = reinterpret_cast<uint32_t*>(my + SM::PW);
  float* sc = smem_bwd + SM::WARPS * SM::WARP_FLOATS;
  float* s_red = sc + CONV_O;
  float* s_w = smem_bwd;  // block-level dW reduction buffer, aliases warp 0's slice; only used after the row loop
  static_assert(Cfg::SW <= 2 * CONV_O * 32, "obs staging aliases the dz planes");
  static_assert(M::MT * 16 * CONV_O <= SM::WARP_FLOATS * SM::WARPS, "dW reduction buffer aliases the warp slices");
  static_assert(Cfg::PW <= 32, "one packed observation word per lane");
  if (tid < CONV_O) sc[tid] = __ldg(params + (int64_t)seed * P + L.ln0_scale + tid);
  if (tid < 3 * CONV_O) s_red[tid] = 0.f;
  __syncthreads();
  float a_dsc[4] = {0.f, 0.f, 0.f, 0.f}, a_dbi[4] = {0.f, 0.f, 0.f, 0.f}, a_dcb[4] = {0.f, 0.f, 0.f, 0.f};
  float wrun[M::MT][2][4];
#pragma unroll
  for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) wrun[mt][h][j] = 0.f;
  const uint32_t amask = (1u << (10 + (g & 3))) | (1u << (26 + (g & 3)));
  const int pos_g = g < 4 ? 2 * g : 2 * (g - 4) + 1;   // pair g of a k-step sits next to pair g + 4

  const int row_stride = gridDim.x * SM::WARPS;
  // unconditional loads from clamped addresses (rows past the end re-read the last row, lanes past the packed width
  // re-read word 0; neither is ever used): a predicated load with a default value made ptxas copy the result into the
  // loop-carried register ~125 instructions after the LDG, which stalled every warp on the load it had just issued
  // (13 % of all stall samples of the forward kernel, ncu r2m)
  auto fetch_index = [&](int r) -> int {
    const int rc = r < rows ? r : rows - 1;
    return gather ? __ldg(gather + (int64_t)seed * rows + rc) : rc;
  };
  auto fetch_obs = [&](int src) -> uint32_t {
    return __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + (lane < Cfg::PW ? lane : 0));
  };
  auto fetch_rows = [&](int r) {  // async copy of the sample's dy / xhat / rstd rows into this warp's slice
    if (r < rows) {
      const int64_t gr = (int64_t)seed * rows + r;
      const float* dsrc = DY1 + gr * FLAT_CNN;
      const float* xsrc = XH1 + gr * FLAT_CNN;
#pragma unroll
      for (int i = 0; i < FLAT_CNN / 4 / 32; ++i) {
        const int q = i * 32 + lane, prow = q >> 2;              // 16-byte chunk q = (pixel row, column quad)
        const int dst = cswz16(prow, (q & 3) * 4);               // the swizzles move whole chunks
        cp_async16(my_dy + dst, dsrc + q * 4);
        cp_async16(my_xh + dst, xsrc + q * 4);
      }
      if (lane < CONV_PIX / 4) cp_async16(my_rs + lane * 4, RS1 + gr * CONV_PIX + lane * 4);
    }
    cp_async_commit();
  };
  int row = blockIdx.x * SM::WARPS + warp;
  uint32_t pre = fetch_obs(fetch_index(row));
  int src_next = fetch_index(row + row_stride);
  fetch_rows(row);
  for (; row < rows; row += row_stride) {
    __syncwarp();
    if (lane < Cfg::PW) my_so[lane] = pre;
    if (lane == 0) my_so[Cfg::PW] = 0u;  // pad word read by the funnel shift of the last pixel
    __syncwarp();
    pre = fetch_obs(src_next);
    src_next = fetch_index(row + 2 * row_stride);
    {  // exponent-coded pair words of pixel pair `lane`
      uint32_t w0[PWD], w1[PWD];
      patch_bits<C>(my_so, 2 * lane, w0);
      patch_bits<C>(my_so, 2 * lane + 1, w1);
      uint32_t pw[SM::NGP];
#pragma unroll
      for (int k = 0; k < SM::NGP; ++k) pw[k] = 0u;
#pragma unroll
      for (int q = 0; q < SM::NG; ++q) {
        // memory order inside an m-tile: groups (0, 2, 1, 3), so the words of fragment rows g and g + 8 are adjacent
        const int ql = q & 3, slot = (q & ~3) + (ql == 1 ? 2 : (ql == 2 ? 1 : ql));
        const int bit = 4 * q, wi = bit >> 5, sh = bit & 31;
        if (wi < PWD) {
          const uint32_t v0 = w0[wi < PWD ? wi : 0], v1 = w1[wi < PWD ? wi : 0];
          const uint32_t lo = sh >= 10 ? v0 >> (sh >= 10 ? sh - 10 : 0) : v0 << (sh < 10 ? 10 - sh : 0);
          const uint32_t hi = sh <= 26 ? v1 << (sh <= 26 ? 26 - sh : 0) : v1 >> (sh > 26 ? sh - 26 : 0);
          pw[slot] = __byte_perm(lo, hi, 0x7610);
        }
      }
      uint32_t* dstw = my_pw + lane * SM::NGP;
#pragma unroll
      for (int k = 0; k < SM::NGP / 4; ++k)
        *reinterpret_cast<uint4*>(dstw + 4 * k) = make_uint4(pw[4 * k], pw[4 * k + 1], pw[4 * k + 2], pw[4 * k + 3]);
    }
    cp_async_wait_all();
    __syncwarp();
    // ---- phase A: LayerNorm backward of the pixel pair (16 mb + 2g, + 1); dz * gs -> fp16 (hi, lo) planes
#pragma unroll 1
    for (int mb = 0; mb < 4; ++mb) {
      const int p0 = 16 * mb + 2 * g, p1 = p0 + 1;
      float z[2][4];
      float2 dyv[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        dyv[h][0] = *reinterpret_cast<const float2*>(my_dy + cswz16(p0, 8 * h + 2 * t));
        dyv[h][1] = *reinterpret_cast<const float2*>(my_dy + cswz16(p1, 8 * h + 2 * t));
        const float2 a0 = *reinterpret_cast<const float2*>(my_xh + cswz16(p0, 8 * h + 2 * t));
        const float2 a1 = *reinterpret_cast<const float2*>(my_xh + cswz16(p1, 8 * h + 2 * t));
        z[h][0] = a0.x; z[h][1] = a0.y; z[h][2] = a1.x; z[h][3] = a1.y;
      }
      // rstd * gs: dz comes out pre-scaled for the fp16 planes (gs is a power of two; the conv-bias sum is unscaled at
      // the end), which saves a multiplication per element
      const float rstd0 = my_rs[p0] * gs, rstd1 = my_rs[p1] * gs;
      float dxh[2][4];
      float m1a = 0.f, m2a = 0.f, m1b = 0.f, m2b = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = 8 * h + 2 * t;
        const float dy4[4] = {dyv[h][0].x, dyv[h][0].y, dyv[h][1].x, dyv[h][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 2 * h + (j & 1);
          a_dsc[col] = fmaf(dy4[j], z[h][j], a_dsc[col]);
          a_dbi[col] += dy4[j];
          dxh[h][j] = dy4[j] * sc[o + (j & 1)];
          if (j < 2) { m1a += dxh[h][j]; m2a = fmaf(dxh[h][j], z[h][j], m2a); }
          else { m1b += dxh[h][j]; m2b = fmaf(dxh[h][j], z[h][j], m2b); }
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        m1a += __shfl_xor_sync(0xffffffffu, m1a, o); m2a += __shfl_xor_sync(0xffffffffu, m2a, o);
        m1b += __shfl_xor_sync(0xffffffffu, m1b, o); m2b += __shfl_xor_sync(0xffffffffu, m2b, o);
      }
      m1a *= (1.0f / CONV_O); m2a *= (1.0f / CONV_O); m1b *= (1.0f / CONV_O); m2b *= (1.0f / CONV_O);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float dzv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float rstd = j < 2 ? rstd0 : rstd1, m1 = j < 2 ? m1a : m1b, m2 = j < 2 ? m2a : m2b;
          dzv[j] = rstd * (dxh[h][j] - m1 - z[h][j] * m2);
          a_dcb[2 * h + (j & 1)] += dzv[j];
        }
        // B words: (pixel p0, pixel p1) of channel o (j = 0, 2) and of channel o + 1 (j = 1, 3); k-step mb, pair g
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float v0 = dzv[c], v1 = dzv[2 + c];
          const uint32_t hw = cvt_f16x2_satfinite(v0, v1);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw));
          const __half2 lw = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          const int wd = dzt_word(8 * h + 2 * t + c, mb, pos_g);
          my_dzt[wd] = hw;
          my_dzt[CONV_O * 32 + wd] = *reinterpret_cast<const uint32_t*>(&lw);
        }
      }
    }
    __syncwarp();
    fetch_rows(row + row_stride);  // overlaps phase B
    // ---- phase B: dW[tap][o] += sum_pixels x[pixel, tap] * dz[pixel][o]   (fresh accumulators per sample)
    float wacc[M::MT][2][4];
#pragma unroll
    for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) wacc[mt][h][j] = 0.f;
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {
      // B fragments: b0 = dz[pixels 16ks + 2t, + 1][o = 8h + g], b1 = the same of pixels + 8: one LDS.64 per plane
      uint2 bhi[2], blo[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int wd = dzt_word(8 * h + g, ks, 2 * t);
        bhi[h] = *reinterpret_cast<const uint2*>(my_dzt + wd);
        blo[h] = *reinterpret_cast<const uint2*>(my_dzt + CONV_O * 32 + wd);
      }
      const uint32_t* pa = my_pw + (8 * ks + t) * SM::NGP + 2 * (g >> 2);   // pixel pair 8ks + t: rows (g, g + 8)
      const uint32_t* pb = pa + 4 * SM::NGP;                               // pixel pair 8ks + t + 4
#pragma unroll
      for (int mt = 0; mt < M::MT; ++mt) {
        const uint2 wa = *reinterpret_cast<const uint2*>(pa + 4 * mt), wb = *reinterpret_cast<const uint2*>(pb + 4 * mt);
        uint32_t a[4];
        a[0] = wa.x & amask; a[1] = wa.y & amask; a[2] = wb.x & amask; a[3] = wb.y & amask;
        // lo pass of both column halves, then the hi pass: no back-to-back MMAs on one accumulator
        mma_f16_16n8k16(wacc[mt][0], a, blo[0].x, blo[0].y);
        mma_f16_16n8k16(wacc[mt][1], a, blo[1].x, blo[1].y);
        mma_f16_16n8k16(wacc[mt][0], a, bhi[0].x, bhi[0].y);
        mma_f16_16n8k16(wacc[mt][1], a, bhi[1].x, bhi[1].y);
      }
    }
#pragma unroll
    for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) wrun[mt][h][j] += wacc[mt][h][j];
  }
  cp_async_wait_all();
  // ---- reduce and publish (deterministic), as in conv_bwd_mma_kernel; the A coding 2^(2^(g%4) - 15) of this thread's
  // accumulator rows and the dz scale gs are undone here (exact powers of two)
  const float inv_gs = 1.0f / gs;
  const float unscale = __uint_as_float((uint32_t)(127 + 15 - (1 << (g & 3))) << 23) * inv_gs;
  __syncthreads();  // all warps are done with their slices
  for (int w = 0; w < SM::WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int mt = 0; mt < M::MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tap = 16 * mt + g + (j >= 2 ? 8 : 0);
            float* dst = &s_w[tap * CONV_O + 8 * h + 2 * t + (j & 1)];   // tap < 16 * MT: inside the buffer
            *dst = (w == 0 ? 0.f : *dst) + wrun[mt][h][j] * unscale;
          }
    }
    __syncthreads();
  }
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    float v0 = a_dsc[col], v1 = a_dbi[col], v2 = a_dcb[col];
#pragma unroll
    for (int sft = 4; sft <= 16; sft <<= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, sft);
      v1 += __shfl_xor_sync(0xffffffffu, v1, sft);
      v2 += __shfl_xor_sync(0xffffffffu, v2, sft);
    }
    a_dsc[col] = v0; a_dbi[col] = v1; a_dcb[col] = v2;
  }
  for (int w = 0; w < SM::WARPS; ++w) {
    if (warp == w && g == 0) {
#pragma unroll
      for (int col = 0; col < 4; ++col) {
        const int o = 8 * (col >> 1) + 2 * t + (col & 1);
        s_red[o] = (w == 0 ? 0.f : s_red[o]) + a_dsc[col];
        s_red[CONV_O + o] = (w == 0 ? 0.f : s_red[CONV_O + o]) + a_dbi[col];
        s_red[2 * CONV_O + o] = (w == 0 ? 0.f : s_red[2 * CONV_O + o]) + a_dcb[col] * inv_gs;
      }
    }
    __syncthreads();
  }
  float* __restrict__ o = part + ((int64_t)seed * gridDim.x + blockIdx.x) * (M::TAPS * CONV_O + 3 * CONV_O);
  for (int i = tid; i < M::TAPS * CONV_O; i += blockDim.x) o[i] = s_w[i] * (1.0f / 255.0f);
  if (tid < 3 * CONV_O) o[M::TAPS * CONV_O + tid] = s_red[tid];
}

// Sums conv_bwd_mma_kernel's per-CTA partials in a fixed order into the gradients.
// grid = (ceil(n / (256 / SL)), S), block = 256
template <int SL>
__global__ void conv_bwd_final_kernel(const float* __restrict__ part, int nctas, int taps16, float* __restrict__ grads,
                                      int64_t P, pqn_net_layout_t L) {
  const int seed = blockIdx.y;
  const int stride = taps16 + 3 * CONV_O;
  const int i = blockIdx.x * (256 / SL) + threadIdx.x % (256 / SL);
  const float v = ordered_partial_sum<SL>(part + (int64_t)seed * nctas * stride + i, nctas, stride, i < stride);
  if (i >= stride || threadIdx.x >= 256 / SL) return;
  float* __restrict__ gout = grads + (int64_t)seed * P;
  if (i < taps16) gout[L.conv_w + i] = v;
  else if (i < taps16 + CONV_O) gout[L.ln0_scale + (i - taps16)] = v;
  else if (i < taps16 + 2 * CONV_O) gout[L.ln0_bias + (i - taps16 - CONV_O)] = v;
  else gout[L.conv_b + (i - taps16 - 2 * CONV_O)] = v;
}

// ---------------------------------------------------------------------------
// conv forward on tcgen05: per tile of 2 samples (128 output pixels) the 128 producer threads (thread = pixel)
// write their im2col row ({0,1} floats, taps padded to a multiple of 32) straight into shared memory in the
// K-major SWIZZLE_128B operand layout, one elected thread issues tcgen05.mma M=128 x N=16 x K=8 per 8 taps
// against the (hi, lo)-split weights/255 held in shared memory, and the same 128 threads read their pixel's 16
// channels back from TMEM for LayerNorm + ReLU and the stores.  A tiles and TMEM accumulators are double
// buffered: tile i+1 is produced while tile i's MMAs run.
// ---------------------------------------------------------------------------
template <int C>
struct ConvTc {
  static constexpr int TAPS = 9 * C;
  static constexpr int KB = (TAPS + 31) / 32;           // k-blocks of 32 taps (one 128-byte swizzle row each)
  static constexpr int KS = (TAPS + 7) / 8;             // MMA k-steps
  static constexpr int A_TILE = 128 * 128;              // bytes per k-block tile (128 rows x 128 B)
  static constexpr int A_BUF = KB * A_TILE;
  static constexpr int B_TILE = 16 * 128;               // 16 output channels x 128 B
  static constexpr int OFF_A = 0;                       // 2 buffers
  static constexpr int OFF_BHI = 2 * A_BUF;
  static constexpr int OFF_BLO = OFF_BHI + KB * B_TILE;
  static constexpr int OFF_MISC = OFF_BLO + KB * B_TILE;  // barriers, tmem slot, obs words, consts
  static constexpr int SMEM = OFF_MISC + 1024 + 1024 /*align slack*/;
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int C, bool TRAIN>
__global__ void __launch_bounds__(160)
    conv_fwd_tc_kernel(const uint32_t* __restrict__ obs, int64_t obs_rows_per_seed, const int32_t* __restrict__ gather,
                       const float* __restrict__ params, int64_t P, pqn_net_layout_t L, float* __restrict__ H1,
                       float* __restrict__ H1LO, float* __restrict__ XH1, float* __restrict__ RS1,
                       float* __restrict__ bn_sums, int rows, int tiles_per_seed, int ctas_per_seed) {
  using Cfg = ConvCfg<C>;
  using T = ConvTc<C>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (sbase - tc::smem_u32(smem_raw));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm + T::OFF_MISC);  // [2]
  uint64_t* d_full = a_full + 2;                                      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);
  float* cb = reinterpret_cast<float*>(sm + T::OFF_MISC + 64);       // [16] conv bias, LN scale, LN bias
  float* sc = cb + CONV_O;
  float* bi = sc + CONV_O;
  float* s_cnt = bi + CONV_O;                                         // [C]
  uint32_t* sobs = reinterpret_cast<uint32_t*>(sm + T::OFF_MISC + 320);  // [2 buf][2 samples][SW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seed = blockIdx.y;
  const float* __restrict__ prm = params + (int64_t)seed * P;

  // ---- one-time setup: weights (hi, lo) as K-major SW128 B tiles, constants, barriers, TMEM
  {
    const float inv255 = 1.0f / 255.0f;
    for (int i = tid; i < T::KB * 16 * 32; i += blockDim.x) {
      const int kb = i / (16 * 32), n = (i / 32) % 16, kk = i % 32;
      const int tap = kb * 32 + kk;
      const float w = tap < T::TAPS ? __ldg(prm + L.conv_w + tap * CONV_O + n) * inv255 : 0.f;
      const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
      const int off = kb * T::B_TILE + n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
      *reinterpret_cast<float*>(sm + T::OFF_BHI + off) = hi;
      *reinterpret_cast<float*>(sm + T::OFF_BLO + off) = w - hi;
    }
    if (tid < CONV_O) {
      cb[tid] = __ldg(prm + L.conv_b + tid);
      sc[tid] = __ldg(prm + L.ln0_scale + tid);
      bi[tid] = __ldg(prm + L.ln0_bias + tid);
    }
    if (tid < C) s_cnt[tid] = 0.f;
    if (tid == 128) {
      for (int b = 0; b < 2; ++b) { tc::mbar_init(&a_full[b], 128); tc::mbar_init(&d_full[b], 1); }
      tc::fence_barrier_init();
    }
    if (warp == 4) { tc::tmem_alloc(tmem_slot, 32); tc::tmem_relinquish(); }
    fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
  }
  const uint32_t tmem_base = *tmem_slot;
  const int n_iters = (tiles_per_seed - (int)blockIdx.x + ctas_per_seed - 1) / ctas_per_seed;  // tiles of this CTA

  if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_tf32(128, 16, 0, 0);
      for (int it = 0; it < n_iters; ++it) {
        const int b = it & 1;
        tc::mbar_wait(&a_full[b], (it >> 1) & 1);
        tc::tcgen05_fence_after();
        const uint32_t d = tmem_base + b * 16;
        const uint32_t abase = sbase + T::OFF_A + b * T::A_BUF;
#pragma unroll
        for (int ks = 0; ks < T::KS; ++ks) {
          const int kb = ks >> 2, kq = ks & 3;
          const uint64_t da = tc::make_sdesc<0>(abase + kb * T::A_TILE, kq);
          const uint64_t dlo = tc::make_sdesc<0>(sbase + T::OFF_BLO + kb * T::B_TILE, kq);
          const uint64_t dhi = tc::make_sdesc<0>(sbase + T::OFF_BHI + kb * T::B_TILE, kq);
          tc::umma_tf32(d, da, dlo, idesc, ks > 0 ? 1u : 0u);
          tc::umma_tf32(d, da, dhi, idesc, 1u);
        }
        tc::umma_commit(&d_full[b]);
      }
    }
  } else {
    // ===================== producers / epilogue: thread = output pixel of the 2-sample tile =====================
    const int sl = tid >> 6, pix = tid & 63, y = pix >> 3, x = pix & 7;
    int cnt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) cnt[c] = 0;

    // observation word of this thread for tile `it` (issued one iteration ahead to hide the HBM latency)
    auto fetch_word = [&](int it) -> uint32_t {
      const int row = (blockIdx.x + it * ctas_per_seed) * 2 + sl;
      if (it < n_iters && row < rows && pix < Cfg::PW) {
        const int64_t src = gather ? gather[(int64_t)seed * rows + row] : row;
        return __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * Cfg::PW + pix);
      }
      return 0u;
    };
    uint32_t w_next = fetch_word(0);

    auto produce = [&](int it) {
      const int b = it & 1;
      const int tile = blockIdx.x + it * ctas_per_seed;
      const int row = tile * 2 + sl;
      uint32_t* so = sobs + (b * 2 + sl) * Cfg::SW;
      if (pix < Cfg::SW) so[pix] = w_next;
      w_next = fetch_word(it + 1);
      bar_sync_named(1, 128);
      // im2col row of this pixel: taps (di,dj,c) -> {0,1}
      float f[T::KB * 32];
#pragma unroll
      for (int k = 0; k < T::KB * 32; ++k) f[k] = 0.f;
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        const uint32_t nib = pixel_bits<C>(so, (y + r / 3) * 10 + x + r % 3);
#pragma unroll
        for (int c = 0; c < C; ++c) f[r * C + c] = ((nib >> c) & 1u) ? 1.0f : 0.0f;
      }
      uint8_t* arow = sm + T::OFF_A + b * T::A_BUF + tid * 128;
#pragma unroll
      for (int kb = 0; kb < T::KB; ++kb)
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<float4*>(arow + kb * T::A_TILE + ((c16 ^ (tid & 7)) << 4)) =
              make_float4(f[kb * 32 + 4 * c16], f[kb * 32 + 4 * c16 + 1], f[kb * 32 + 4 * c16 + 2], f[kb * 32 + 4 * c16 + 3]);
      if (TRAIN && bn_sums != nullptr && row < rows) {
        const uint32_t b0 = pixel_bits<C>(so, pix);
        const uint32_t b1 = (pix + 64 < 100) ? pixel_bits<C>(so, pix + 64) : 0u;
#pragma unroll
        for (int c = 0; c < C; ++c) cnt[c] += (int)((b0 >> c) & 1u) + (int)((b1 >> c) & 1u);
      }
      fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc::mbar_arrive(&a_full[b]);
    };

    auto epilogue = [&](int it) {
      const int b = it & 1;
      const int tile = blockIdx.x + it * ctas_per_seed;
      const int row = tile * 2 + sl;
      tc::mbar_wait(&d_full[b], (it >> 1) & 1);
      tc::tcgen05_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + b * 16 + ((uint32_t)(warp * 32) << 16), v);
      tc::tmem_ld_wait();
      tc::tcgen05_fence_before();
      float z[CONV_O];
#pragma unroll
      for (int o = 0; o < CONV_O; ++o) z[o] = __uint_as_float(v[o]) + cb[o];
      float mean, rstd;
      ln16(z, mean, rstd);
      if (row < rows) {
        const int64_t base = ((int64_t)seed * rows + row) * FLAT_CNN + pix * CONV_O;
#pragma unroll
        for (int o4 = 0; o4 < CONV_O / 4; ++o4) {
          float xh[4], hv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xh[j] = (z[4 * o4 + j] - mean) * rstd;
            hv[j] = fmaxf(xh[j] * sc[4 * o4 + j] + bi[4 * o4 + j], 0.f);
          }
          *reinterpret_cast<float4*>(H1 + base + 4 * o4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          if (H1LO)
            *reinterpret_cast<float4*>(H1LO + base + 4 * o4) =
                make_float4(tc::tf32_lo(hv[0]), tc::tf32_lo(hv[1]), tc::tf32_lo(hv[2]), tc::tf32_lo(hv[3]));
          if (TRAIN && XH1) *reinterpret_cast<float4*>(XH1 + base + 4 * o4) = make_float4(xh[0], xh[1], xh[2], xh[3]);
        }
        if (TRAIN && RS1) RS1[((int64_t)seed * rows + row) * CONV_PIX + pix] = rstd;
      }
    };

    if (n_iters > 0) produce(0);
    for (int it = 1; it < n_iters; ++it) {
      produce(it);
      epilogue(it - 1);
    }
    if (n_iters > 0) epilogue(n_iters - 1);

    if (TRAIN && bn_sums != nullptr) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        int vsum = cnt[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if (lane == 0 && vsum) atomicAdd(&s_cnt[c], (float)vsum);
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (TRAIN && bn_sums != nullptr && tid < C && s_cnt[tid] != 0.f) {
    atomicAdd(bn_sums + (int64_t)seed * 2 * C + tid, s_cnt[tid]);
    atomicAdd(bn_sums + (int64_t)seed * 2 * C + C + tid, s_cnt[tid]);
  }
  if (warp == 4) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem_base, 32);
  }
}

template <int C, bool TRAIN>
static int launch_conv_fwd_tc_t(int S, cudaStream_t st, const uint32_t* obs, int64_t orps, const int32_t* gather,
                                const float* params, int64_t P, const pqn_net_layout_t& L, float* h1, float* h1lo,
                                float* xh1, float* rs1, float* bn, int rows) {
  auto kfn = conv_fwd_tc_kernel<C, TRAIN>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTc<C>::SMEM) != cudaSuccess)
    return check_launch("conv_fwd_tc(cudaFuncSetAttribute)");
  const int tiles = (rows + 1) / 2;
  // all CTAs co-resident (3 per SM by shared memory): a second partial wave would double the time
  int per_seed = (148 * 3) / S;
  if (per_seed > tiles) per_seed = tiles;
  if (per_seed < 1) per_seed = 1;
  {
    LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st);
    kfn<<<dim3(per_seed, S), 160, ConvTc<C>::SMEM, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, bn, rows,
                                                         tiles, per_seed);
  }
  return 0;
}

// MLP input gather (minibatch rows of float obs); the input BatchNorm sums come from nrm::colsum2 of the result.
__global__ void gather_rows_kernel(const float* __restrict__ obs, int64_t obs_rows_per_seed,
                                   const int32_t* __restrict__ gather, float* __restrict__ out, int rows, int D) {
  const int seed = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (int64_t)rows * D) return;
  const int r = (int)(g / D), j = (int)(g - (int64_t)r * D);
  const int64_t src = gather ? gather[(int64_t)seed * rows + r] : r;
  const float v = __ldg(obs + ((int64_t)seed * obs_rows_per_seed + src) * D + j);
  out[((int64_t)seed * rows + r) * D + j] = v;
}

// ---------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------
struct Workspace {
  // CNN
  float *h1, *h2, *xhat2, *rstd2, *dz2;
  float *h1_lo, *dz2_lo, *w1_lo;  // 3xTF32 "lo" operands of the tcgen05 path
  float *cxhat, *crstd;           // conv LayerNorm xhat / rstd saved by the training forward (MMA conv path)
  uint32_t* relu_bits;            // packed (h1 > 0) mask, 1024 bits per row (MMA conv path -> tcgen05 dgrad epilogue)
  float *rb_part, *cb_part;       // per-CTA partial vectors of the deterministic row_bwd / conv_bwd reductions
  float* wg_part;                 // split-K partial outputs of the tensor-core weight gradient (small S: few output tiles)
  // MLP
  float *xg, *h0, *xhat0, *rstd0, *hh1, *xhat1, *rstd1, *dzl, *dh0;
  float *m16_h0, *m16_w, *m16_dz;   // fp16 (hi, lo') planes of h0 / Dense_1 kernel / dz1 for the tensor-core hidden layer
};

// split-K of the tensor-core weight gradient: when S * m_tiles * n_tiles output tiles cannot fill the SMs (one seed of
// the MLP has 4 tiles, of the CNN 8), the K = rows range is divided so that one CTA per SM runs; the partial tiles
// (at most WGRAD_SPLIT_TILES of them) are added in split order by wgrad_split_reduce_kernel (deterministic)
constexpr int64_t WGRAD_SPLIT_TILES = 4 * 148 + 16;   // tensor-core split-K: <= SMs; FFMA row splits: tiles * S * splits < 4 * 148
static int wgrad_ksplit(int tiles_total, int k_blocks) {
  const int sms = device_sm_count();                 // persistent kernel, one CTA per SM: one wave of split tiles
  if (tiles_total >= sms) return 1;
  int ks = sms / tiles_total;
  if (ks > k_blocks / 8) ks = k_blocks / 8;          // keep >= 8 k-blocks per CTA
  if (ks < 1) ks = 1;
  while (ks > 1 && (int64_t)(ks - 1) * ((k_blocks + ks - 1) / ks) >= k_blocks) --ks;   // no empty split
  return ks;
}
template <int SL>
__global__ void wgrad_split_reduce_kernel(const float* __restrict__ part, int ksplit, int64_t split_stride, int64_t n_per_seed,
                                          float* __restrict__ out, int64_t out_seed_stride) {
  const int seed = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * (256 / SL) + threadIdx.x % (256 / SL);
  const float v = ordered_partial_sum<SL>(part + (int64_t)seed * n_per_seed + i, ksplit, split_stride, i < n_per_seed);
  if (i < n_per_seed && threadIdx.x < 256 / SL) out[(int64_t)seed * out_seed_stride + i] = v;
}
// out[seed][i] = sum over k < ksplit of part[k * split_stride + seed * n + i], in k-slice order
static void launch_split_reduce(const float* part, int ksplit, int64_t split_stride, int64_t n, int S, float* out,
                                int64_t out_seed_stride, cudaStream_t st) {
  LaunchScope _ls(K_GRAD_FINAL, st);
  if (final_slices(ksplit) == 32)
    wgrad_split_reduce_kernel<32><<<dim3((unsigned)((n + 7) / 8), S), 256, 0, st>>>(part, ksplit, split_stride, n, out, out_seed_stride);
  else
    wgrad_split_reduce_kernel<8><<<dim3((unsigned)((n + 31) / 32), S), 256, 0, st>>>(part, ksplit, split_stride, n, out, out_seed_stride);
}

// The register-tiled FFMA weight gradient + (splits > 1) its ordered reduction.  `part` needs splits * S * Kin * N floats
// (<= WGRAD_SPLIT_TILES tiles of 128 x 128: wgrad_splits keeps tiles * S * splits below 4 * 148).
static void run_wgrad_ffma(const float* X, int64_t x_seed_stride, int ldx, const float* DZ, int64_t dz_seed_stride, int N,
                           float* grads, int64_t P, int64_t off_w, int rows, int Kin, int S, int splits, float* part,
                           cudaStream_t st) {
  { LaunchScope _ls(K_WGRAD, st);
    wgrad_kernel<<<dim3((unsigned)((Kin + 127) / 128), (unsigned)(N / 128), (unsigned)(S * splits)), GT, 0, st>>>(
        X, x_seed_stride, ldx, DZ, dz_seed_stride, N, grads, P, off_w, rows, Kin, splits, part); }
  if (splits > 1) {
    const int64_t n = (int64_t)Kin * N;
    launch_split_reduce(part, splits, (int64_t)S * n, n, S, grads + off_w, P, st);
  }
}

// Weight gradient of a layer with a *thin* input (the first MLP layer: Kin = observation features <= 8):
// dW[k][n] = sum_rows X[row][k] * dZ[row][n].  A 128 x 128 register tile would spend 94 % of its FFMAs on padding, so
// here every thread owns four output columns (256 / (N/4) row groups) and Kin float4 accumulators, streams its rows
// of dZ coalesced (16-byte loads) and reads the X rows of the CTA's tile from shared memory (broadcast).  Per-CTA partials
// part[chunk][seed][Kin][N] are then added in chunk order by wgrad_split_reduce_kernel: no float atomics.
// grid = (chunks, S), block = 256
constexpr int THIN_KMAX = 8, THIN_ROWS = 128;
__global__ void __launch_bounds__(256) wgrad_thin_kernel(const float* __restrict__ X, int64_t x_seed_stride, int Kin,
                                                         const float* __restrict__ DZ, int64_t dz_seed_stride, int N,
                                                         float* __restrict__ part, int rows, int rows_per_chunk) {
  __shared__ float xs[THIN_ROWS * THIN_KMAX];
  __shared__ float4 red[256];
  const int seed = blockIdx.y, S = gridDim.y;
  const int cols4 = N >> 2;                                             // threads per row (float4 columns); N in {128, 256}
  const int groups = 256 / cols4, c4 = threadIdx.x % cols4, grp = threadIdx.x / cols4;
  const float* __restrict__ Xs = X + (int64_t)seed * x_seed_stride;
  const float4* __restrict__ Zs = reinterpret_cast<const float4*>(DZ + (int64_t)seed * dz_seed_stride);
  const int r_begin = blockIdx.x * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
  float4 acc[THIN_KMAX];
#pragma unroll
  for (int k = 0; k < THIN_KMAX; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = r_begin; r0 < r_end; r0 += THIN_ROWS) {
    const int nr = min(THIN_ROWS, r_end - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * Kin; i += 256) {
      const int r = i / Kin, k = i - r * Kin;
      xs[r * THIN_KMAX + k] = __ldg(Xs + (int64_t)(r0 + r) * Kin + k);
    }
    __syncthreads();
#pragma unroll 4
    for (int r = grp; r < nr; r += groups) {
      const float4 z = __ldg(Zs + (int64_t)(r0 + r) * cols4 + c4);
#pragma unroll
      for (int k = 0; k < THIN_KMAX; ++k)
        if (k < Kin) {
          const float x = xs[r * THIN_KMAX + k];
          acc[k].x = fmaf(x, z.x, acc[k].x); acc[k].y = fmaf(x, z.y, acc[k].y);
          acc[k].z = fmaf(x, z.z, acc[k].z); acc[k].w = fmaf(x, z.w, acc[k].w);
        }
    }
  }
  // add the row groups in group order (one k at a time through shared memory), then store this CTA's partial
  float4* __restrict__ out = reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * S + seed) * Kin * N);
#pragma unroll
  for (int k = 0; k < THIN_KMAX; ++k) {
    if (k >= Kin) break;
    __syncthreads();
    red[threadIdx.x] = acc[k];
    __syncthreads();
    if (grp == 0) {
      float4 v = red[c4];
      for (int g = 1; g < groups; ++g) {
        const float4 o = red[g * cols4 + c4];
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      out[k * cols4 + c4] = v;
    }
  }
}

// upper bound of the CTAs (all seeds) of the wave-sized grids of conv_mma_ctas(): <= 6 waves of <= 4 CTAs/SM, + S
static int64_t part_ctas(int S) { return 6 * 4 * (int64_t)device_sm_count() + 2 * (int64_t)S; }
static int64_t row_bwd_part_floats(int N, int A);
static int wgrad_splits(int tiles, int S, int rows);
static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// first-layer weight gradient dW0[D][H] = X^T . dZ: the thin deterministic kernel when D <= 8 (every shipped classic-control
// env), the register-tiled FFMA kernel otherwise.  `part` is the row_bwd partial buffer (free again at this point of
// the stream; chunks * S * D * H floats are far below its size).
static void run_wgrad_first(const float* X, const float* DZ, float* grads, int64_t P, int64_t off_w, int S, int rows,
                            int D, int H, float* part, float* wg_part, cudaStream_t st) {
  if (D <= THIN_KMAX && (H == 128 || H == 256)) {
    int chunks = (2 * device_sm_count()) / S;
    const int max_chunks = (rows + THIN_ROWS - 1) / THIN_ROWS;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int per = (rows + chunks - 1) / chunks;
    per = (per + THIN_ROWS - 1) / THIN_ROWS * THIN_ROWS;
    chunks = (rows + per - 1) / per;
    const int64_t n = (int64_t)D * H;
    { LaunchScope _ls(K_WGRAD, st);
      wgrad_thin_kernel<<<dim3(chunks, S), 256, 0, st>>>(X, (int64_t)rows * D, D, DZ, (int64_t)rows * H, H, part, rows, per); }
    launch_split_reduce(part, chunks, (int64_t)S * n, n, S, grads + off_w, P, st);
    return;
  }
  run_wgrad_ffma(X, (int64_t)rows * D, D, DZ, (int64_t)rows * H, H, grads, P, off_w, rows, D, S,
                 wgrad_splits((D + 127) / 128 * (H / 128), S, rows), wg_part, st);
}

static int64_t carve(const pqn_net_desc_t* d, int32_t S, int64_t rows, char* base, Workspace* w) {
  int64_t off = 0;
  auto take = [&](int64_t nfloats) -> float* {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (nfloats * 4 + 255) / 256 * 256;
    return p;
  };
  const int64_t R = (int64_t)S * rows;
  Workspace tmp;
  Workspace* ww = w ? w : &tmp;
  if (d->kind == PQN_NET_MINATAR_CNN) {
    ww->h1 = take(R * FLAT_CNN);
    ww->h2 = take(R * HID_CNN);
    ww->xhat2 = take(R * HID_CNN);
    ww->rstd2 = take(R);
    ww->dz2 = take(R * HID_CNN);
    ww->h1_lo = take(R * FLAT_CNN);
    ww->dz2_lo = take(R * HID_CNN);
    ww->w1_lo = take((int64_t)S * FLAT_CNN * HID_CNN);
    ww->cxhat = take(R * FLAT_CNN);
    ww->crstd = take(R * CONV_PIX);
    ww->relu_bits = reinterpret_cast<uint32_t*>(take(R * (FLAT_CNN / 32)));
    ww->rb_part = take(part_ctas(S) * row_bwd_part_floats(HID_CNN, d->num_actions));
    ww->cb_part = take(part_ctas(S) * (int64_t)(9 * d->in_c * CONV_O + 3 * CONV_O));
    ww->wg_part = take(WGRAD_SPLIT_TILES * 128 * 128);
  } else {
    const int H = d->hidden;
    ww->xg = take(R * d->in_c);
    ww->h0 = take(R * H);
    ww->xhat0 = take(R * H);
    ww->rstd0 = take(R);
    ww->hh1 = take(R * H);
    ww->xhat1 = take(R * H);
    ww->rstd1 = take(R);
    ww->dzl = take(R * H);
    ww->dh0 = take(R * H);
    ww->rb_part = take(part_ctas(S) * row_bwd_part_floats(H, d->num_actions));
    ww->cb_part = nullptr;
    ww->wg_part = take(WGRAD_SPLIT_TILES * 128 * 128);
    ww->m16_h0 = take(R * H);                       // 2 planes x 2 bytes = 4 bytes per element
    ww->m16_w = take((int64_t)S * H * H);
    ww->m16_dz = take(R * H);
  }
  return off;
}

template <int MODE>
static void launch_dense(int BN, dim3 grid, cudaStream_t st, const float* X, int64_t xss, int ldx, const float* params,
                         int64_t P, int64_t ow, int64_t ob, int64_t osc, int64_t obi, int64_t ohw, int64_t ohb, int A,
                         float* H, float* XH, float* RS, float* Q, int rows, int K) {
  if (BN == 128)
    { LaunchScope _ls(K_DENSE_FWD, st); dense_fwd_kernel<128, MODE><<<grid, GT, 0, st>>>(X, xss, ldx, params, P, ow, ob, osc, obi, ohw, ohb, A, H, XH, RS, Q, rows, K); }
  else
    { LaunchScope _ls(K_DENSE_FWD, st); dense_fwd_kernel<256, MODE><<<grid, GT, 0, st>>>(X, xss, ldx, params, P, ow, ob, osc, obi, ohw, ohb, A, H, XH, RS, Q, rows, K); }
}


template <int C>
static int launch_conv_bwd_mma(dim3 grid, cudaStream_t st, const uint32_t* obs, int64_t orps, const int32_t* gather,
                               const float* params, int64_t P, const pqn_net_layout_t& L, const float* dy1,
                               const float* xh1, const float* rs1, float* grads, float* part, int rows) {
  auto kfn = conv_bwd_mma_kernel<C>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvBwdSmem<C>::BYTES) != cudaSuccess)
    return check_launch("conv_bwd_mma(cudaFuncSetAttribute)");
  kfn<<<grid, ConvBwdSmem<C>::WARPS * 32, ConvBwdSmem<C>::BYTES, st>>>(obs, orps, gather, params, P, L, dy1, xh1, rs1,
                                                                       part, rows);
  const int n = 9 * C * CONV_O + 3 * CONV_O;
  if (final_slices((int)grid.x) == 32)
    conv_bwd_final_kernel<32><<<dim3(cdiv(n, 8), grid.y), 256, 0, st>>>(part, (int)grid.x, 9 * C * CONV_O, grads, P, L);
  else
    conv_bwd_final_kernel<8><<<dim3(cdiv(n, 32), grid.y), 256, 0, st>>>(part, (int)grid.x, 9 * C * CONV_O, grads, P, L);
  return 0;
}

// the fp16 variant (conv path 1); gs = power-of-two scale of dz before the fp16 split
template <int C>
static int launch_conv_bwd_mma16(dim3 grid, cudaStream_t st, const uint32_t* obs, int64_t orps, const int32_t* gather,
                                 const float* params, int64_t P, const pqn_net_layout_t& L, const float* dy1,
                                 const float* xh1, const float* rs1, float* grads, float* part, int rows, float gs) {
  auto kfn = conv_bwd_mma16_kernel<C>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvBwd16<C>::BYTES) != cudaSuccess)
    return check_launch("conv_bwd_mma16(cudaFuncSetAttribute)");
  kfn<<<grid, ConvBwd16<C>::WARPS * 32, ConvBwd16<C>::BYTES, st>>>(obs, orps, gather, params, P, L, dy1, xh1, rs1, part,
                                                                   rows, gs);
  const int n = 9 * C * CONV_O + 3 * CONV_O;
  if (final_slices((int)grid.x) == 32)
    conv_bwd_final_kernel<32><<<dim3(cdiv(n, 8), grid.y), 256, 0, st>>>(part, (int)grid.x, 9 * C * CONV_O, grads, P, L);
  else
    conv_bwd_final_kernel<8><<<dim3(cdiv(n, 32), grid.y), 256, 0, st>>>(part, (int)grid.x, 9 * C * CONV_O, grads, P, L);
  return 0;
}

// dynamic shared memory of row_bwd_kernel<N, HEAD> (floats: eight warp-private [3N] reduction slices; HEAD: [A][N]
// weights, eight warp-private [A][N] gradient slices, [8][A] + [8][2] scalars)
static size_t row_bwd_smem(int N, int A, bool head) {
  return (size_t)(24 * N + (head ? A * N + 8 * A + 16 + 8 * A * N : 0)) * sizeof(float);
}
static int64_t row_bwd_part_floats(int N, int A) { return 3 * (int64_t)N + (int64_t)A * N + A + 2; }

template <int N, bool HEAD, typename... Args>
static int launch_row_bwd(dim3 grid, int A, cudaStream_t st, Args... args) {
  const size_t sm = row_bwd_smem(N, A, HEAD);
  if (sm > 48 * 1024 &&
      cudaFuncSetAttribute(row_bwd_kernel<N, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
    return check_launch("row_bwd(cudaFuncSetAttribute)");
  LaunchScope _ls(K_ROW_BWD, st);
  row_bwd_kernel<N, HEAD><<<grid, 256, sm, st>>>(args...);
  return 0;
}
static void launch_row_bwd_final(const float* part, dim3 rbg, int N, int A, bool head, float* grads, int64_t P,
                                 int64_t off_dscale, int64_t off_dbias, int64_t off_db, int64_t off_hw, int64_t off_hb,
                                 float* loss_sum, float* qsa_sum, cudaStream_t st) {
  const int stride = 3 * N + (head ? A * N + A + 2 : 0);
  LaunchScope _ls(K_GRAD_FINAL, st);
  if (final_slices((int)rbg.x) == 32)
    row_bwd_final_kernel<32><<<dim3(cdiv(stride, 8), rbg.y), 256, 0, st>>>(part, (int)rbg.x, N, A, head ? 1 : 0, grads, P, off_dscale,
                                                                           off_dbias, off_db, off_hw, off_hb, loss_sum, qsa_sum);
  else
    row_bwd_final_kernel<8><<<dim3(cdiv(stride, 32), rbg.y), 256, 0, st>>>(part, (int)rbg.x, N, A, head ? 1 : 0, grads, P, off_dscale,
                                                                           off_dbias, off_db, off_hw, off_hb, loss_sum, qsa_sum);
}

// row_bwd + its fixed-order finalize.  off_scale = LayerNorm scale of this layer (also where d scale goes), off_bias its
// bias slot, off_db the bias of the dense layer before it.
static int run_row_bwd(int N, bool head, dim3 rbg, int A, cudaStream_t st, const float* Hh, const float* XHAT,
                       const float* RSTD, const float* DH, float* DZ, float* DZLO, __half* DZ16H, __half* DZ16L,
                       float gscale, const float* params, float* grads, int64_t P, int64_t off_scale, int64_t off_bias,
                       int64_t off_db, int64_t off_hw, int64_t off_hb, const int32_t* gather, const int32_t* action,
                       const float* target, int64_t trps, float* loss_sum, float* qsa_sum, float* part, int rows) {
  int rc = 0;
#define PQN_RB(NN, HH)                                                                                              \
  rc = launch_row_bwd<NN, HH>(rbg, A, st, Hh, XHAT, RSTD, DH, DZ, DZLO, DZ16H, DZ16L, gscale, params, grads, P,     \
                              off_scale, off_scale, off_bias, off_db, off_hw, off_hb, A, gather, action, target, trps, \
                              part, rows)
  if (N == 128 && head) PQN_RB(128, true);
  else if (N == 128) PQN_RB(128, false);
  else if (N == 256 && head) PQN_RB(256, true);
  else if (N == 256) PQN_RB(256, false);
  else return set_error(PQN_E_UNSUPPORTED, "row_bwd width %d", N);
#undef PQN_RB
  if (rc) return rc;
  launch_row_bwd_final(part, rbg, N, A, head, grads, P, off_scale, off_bias, off_db, off_hw, off_hb, loss_sum, qsa_sum, st);
  return 0;
}

// CTAs per seed for the warp-per-sample conv kernels (grid = per_seed x S).  `resident` = CTAs the GPU holds at once
// (SMs x CTAs/SM): the grid is sized to fill whole waves of that many CTAs -- 1280 CTAs on 296 slots would run a
// fifth, 32%-full wave -- while staying near 4 waves so that per-CTA setup (weight fragments) stays amortised.
static unsigned conv_mma_ctas(int S, int rows, int ctas_per_sm) {
  const int sms = device_sm_count();
  const int resident = sms * ctas_per_sm;
  const int maxc = (rows + CONV_MMA_WARPS - 1) / CONV_MMA_WARPS;
  int best = 1;
  double best_eff = 0.0;
  const int lim = (6 * resident + S - 1) / S;
  for (int per_seed = 1; per_seed <= lim && per_seed <= maxc; ++per_seed) {
    const int total = per_seed * S;
    const int waves = (total + resident - 1) / resident;
    if (waves > 6) break;
    const double eff = (double)total / ((double)waves * resident);
    if (eff > best_eff + 0.01 || (eff > best_eff - 0.01 && waves <= 4)) { best_eff = eff > best_eff ? eff : best_eff; best = per_seed; }
  }
  return (unsigned)best;
}

template <bool TRAIN>
static int launch_conv_fwd(int C, dim3 grid, cudaStream_t st, const uint32_t* obs, int64_t orps, const int32_t* gather,
                           const float* params, int64_t P, const pqn_net_layout_t& L, float* h1, float* h1lo, float* bn,
                           int rows, float* xh1 = nullptr, float* rs1 = nullptr, uint32_t* rb = nullptr,
                           bool h16 = false) {
  if (g_conv_mma == 2) {
    const int S = (int)grid.y;
    switch (C) {
      case 4: return launch_conv_fwd_tc_t<4, TRAIN>(S, st, obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, bn, rows);
      case 6: return launch_conv_fwd_tc_t<6, TRAIN>(S, st, obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, bn, rows);
      case 7: return launch_conv_fwd_tc_t<7, TRAIN>(S, st, obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, bn, rows);
      case 10: return launch_conv_fwd_tc_t<10, TRAIN>(S, st, obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, bn, rows);
      default: return -1;
    }
  }
  if (g_conv_mma == 1) {  // fp16 mma.sync conv (default); h16: h1 / h1lo are the fp16 (hi, lo') planes
    const dim3 mg(conv_mma_ctas((int)grid.y, rows, CONV16_CTAS_PER_SM), grid.y);
    LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st);
#define PQN_CONV16(CC)                                                                                              \
  if (h16) conv_fwd_mma16_kernel<CC, TRAIN, true><<<mg, CONV16_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows); \
  else conv_fwd_mma16_kernel<CC, TRAIN, false><<<mg, CONV16_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows)
    switch (C) {
      case 4: PQN_CONV16(4); break;
      case 6: PQN_CONV16(6); break;
      case 7: PQN_CONV16(7); break;
      case 10: PQN_CONV16(10); break;
      default: return -1;
    }
#undef PQN_CONV16
    return 0;
  }
  if (g_conv_mma == 3) {  // the round-1 tf32 mma.sync conv (kept as an A/B reference)
    const dim3 mg(conv_mma_ctas((int)grid.y, rows, 3), grid.y);
    LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st);
    switch (C) {
      case 4: conv_fwd_mma_kernel<4, TRAIN><<<mg, CONV_MMA_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows); break;
      case 6: conv_fwd_mma_kernel<6, TRAIN><<<mg, CONV_MMA_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows); break;
      case 7: conv_fwd_mma_kernel<7, TRAIN><<<mg, CONV_MMA_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows); break;
      case 10: conv_fwd_mma_kernel<10, TRAIN><<<mg, CONV_MMA_WARPS * 32, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, xh1, rs1, rb, bn, rows); break;
      default: return -1;
    }
    return 0;
  }
  switch (C) {
    case 4: { LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st); conv_fwd_kernel<4, TRAIN><<<grid, 256, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, bn, rows); } break;
    case 6: { LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st); conv_fwd_kernel<6, TRAIN><<<grid, 256, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, bn, rows); } break;
    case 7: { LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st); conv_fwd_kernel<7, TRAIN><<<grid, 256, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, bn, rows); } break;
    case 10: { LaunchScope _ls(TRAIN ? K_CONV_FWD : K_CONV_FWD_INFER, st); conv_fwd_kernel<10, TRAIN><<<grid, 256, 0, st>>>(obs, orps, gather, params, P, L, h1, h1lo, bn, rows); } break;
    default: return -1;
  }
  return 0;
}

// CTAs per seed for conv_bwd: ~4 waves of 2 CTAs/SM over all seeds, at most one sample-group per CTA
static unsigned conv_bwd_ctas(int S, int rows) {
  int per_seed = (148 * 2 * 4 + S - 1) / S;
  const int maxc = (rows + CONV_BWD_WARPS - 1) / CONV_BWD_WARPS;
  if (per_seed > maxc) per_seed = maxc;
  if (per_seed < 1) per_seed = 1;
  return (unsigned)per_seed;
}

static int wgrad_splits(int tiles, int S, int rows) {
  int s = (2 * 148 + tiles * S - 1) / (tiles * S);
  const int maxs = (rows + 255) / 256;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

// lo = x - trunc_tf32(x) for the per-seed weight block W1 (source rows at stride P)
__global__ void split_lo_strided_kernel(const float* __restrict__ src, int64_t src_seed_stride, float* __restrict__ lo,
                                        int64_t n_per_seed) {
  const int seed = blockIdx.y;
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n_per_seed) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)seed * src_seed_stride) + i4);
  reinterpret_cast<float4*>(lo + (int64_t)seed * n_per_seed)[i4] =
      make_float4(tc::tf32_lo(v.x), tc::tf32_lo(v.y), tc::tf32_lo(v.z), tc::tf32_lo(v.w));
}

static void launch_split_w1(const float* params, int64_t P, int64_t off_w, float* w1_lo, int S, cudaStream_t st) {
  const int64_t n = (int64_t)FLAT_CNN * HID_CNN;
  LaunchScope _ls(K_TC_SPLIT, st);
  split_lo_strided_kernel<<<dim3(cdiv(n / 4, 256), S), 256, 0, st>>>(params + off_w, P, w1_lo, n);
}

// ---- fp16-split tensor-core path (g_use_tc == 2) ------------------------------------------------------------------
// planes inside the workspace (no extra memory: they alias the fp32 "lo" tensors of the tf32 path, same byte size):
//   h1 planes  : w.h1_lo  region  -> __half hi[R][1024], lo'[R][1024]
//   dz2 planes : w.dz2_lo region  -> __half hi[R][128],  lo'[R][128]      (dz2 * gscale)
//   W1 planes  : w.w1_lo  region  -> __half hi[S][1024][128], lo'[S][1024][128]
struct Planes16 { __half *h1_hi, *h1_lo, *dz_hi, *dz_lo, *w1_hi, *w1_lo; };
static Planes16 planes16(const Workspace& w, int S, int64_t rows) {
  const int64_t R = (int64_t)S * rows;
  Planes16 p;
  p.h1_hi = reinterpret_cast<__half*>(w.h1_lo); p.h1_lo = p.h1_hi + R * FLAT_CNN;
  p.dz_hi = reinterpret_cast<__half*>(w.dz2_lo); p.dz_lo = p.dz_hi + R * HID_CNN;
  p.w1_hi = reinterpret_cast<__half*>(w.w1_lo); p.w1_lo = p.w1_hi + (int64_t)S * FLAT_CNN * HID_CNN;
  return p;
}
// dz2 is pre-scaled by a power of two so that gradient-sized values (|dz2| ~ |diff| / rows) sit in the middle of the
// fp16 range: gscale = 16 * 2^ceil(log2(rows)); the GEMM epilogues multiply by 1 / gscale (exact).
static float grad_scale(int64_t rows) {
  float s = 16.f;
  while (rows > 1) { s *= 2.f; rows = (rows + 1) >> 1; }
  return s;
}

__global__ void split16_strided_kernel(const float* __restrict__ src, int64_t src_seed_stride, __half* __restrict__ hi,
                                       __half* __restrict__ lo, int64_t n_per_seed) {
  const int seed = blockIdx.y;
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n_per_seed) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)seed * src_seed_stride) + i4);
  __half2 h0, h1, l0, l1;
  tc::split16x2(v.x, v.y, h0, l0);
  tc::split16x2(v.z, v.w, h1, l1);
  reinterpret_cast<uint2*>(hi + (int64_t)seed * n_per_seed)[i4] =
      make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  reinterpret_cast<uint2*>(lo + (int64_t)seed * n_per_seed)[i4] =
      make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
}

static void launch_split16_w1(const float* params, int64_t P, int64_t off_w, const Planes16& pl, int S, cudaStream_t st) {
  const int64_t n = (int64_t)FLAT_CNN * HID_CNN;
  LaunchScope _ls(K_TC_SPLIT, st);
  split16_strided_kernel<<<dim3(cdiv(n / 4, 256), S), 256, 0, st>>>(params + off_w, P, pl.w1_hi, pl.w1_lo, n);
}
// fp32 h1 (conv paths that do not write planes themselves) -> planes
static void launch_split16_h1(const float* h1, const Planes16& pl, int64_t n, cudaStream_t st) {
  LaunchScope _ls(K_TC_SPLIT, st);
  split16_strided_kernel<<<dim3(cdiv(n / 4, 256), 1), 256, 0, st>>>(h1, 0, pl.h1_hi, pl.h1_lo, n);
}

static int tc16_dense_fwd(int epi, const float* params, int64_t P, const pqn_net_layout_t& L, const Workspace& w,
                          const Planes16& pl, int A, float* q, int S, int rows, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], pl.h1_hi, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[1], pl.h1_lo, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[2], pl.w1_hi, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[3], pl.w1_lo, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 64))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = 1; gs.k_blocks = FLAT_CNN / tc::TC_BK16; gs.split3 = 1;
  tc::EpiParams ep = {};
  ep.params = params; ep.P = P; ep.off_b = L.d0_b; ep.off_scale = L.ln1_scale; ep.off_bias = L.ln1_bias;
  ep.off_hw = L.head_w; ep.off_hb = L.head_b; ep.A = A; ep.rows = rows;
  ep.H = w.h2; ep.XHAT = w.xhat2; ep.RSTD = w.rstd2; ep.Q = q;
  return tc::launch_gemm16(0, 1, epi, t, gs, ep, st, epi == tc::EPI_LN_HEAD ? K_TC_FWD_HEAD : K_TC_FWD);
}

static int tc16_wgrad(float* grads, int64_t P, const pqn_net_layout_t& L, const Planes16& pl, int S, int rows,
                      float gscale, float* wg_part, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], pl.h1_hi, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[1], pl.h1_lo, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[2], pl.dz_hi, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[3], pl.dz_lo, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 64))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = FLAT_CNN; gs.m_tiles = FLAT_CNN / 128; gs.n_tiles = 1;
  gs.k_blocks = (rows + tc::TC_BK16 - 1) / tc::TC_BK16; gs.split3 = 1;
  gs.k_split = wgrad_ksplit(S * gs.m_tiles * gs.n_tiles, gs.k_blocks);
  tc::EpiParams ep = {};
  ep.ld_out = HID_CNN; ep.out_scale = 1.0f / gscale;
  const int64_t n = (int64_t)FLAT_CNN * HID_CNN;
  if (gs.k_split > 1) { ep.out = wg_part; ep.out_seed_stride = n; ep.split_stride = (int64_t)S * n; }
  else { ep.out = grads + L.d0_w; ep.out_seed_stride = P; }
  if ((rc = tc::launch_gemm16(1, 1, tc::EPI_STORE, t, gs, ep, st, K_TC_WGRAD))) return rc;
  if (gs.k_split > 1) {
    launch_split_reduce(wg_part, gs.k_split, (int64_t)S * n, n, S, grads + L.d0_w, P, st);
  }
  return 0;
}

// dY1 = relu_mask * (dZ2 . W1^T) into w.h1 (fp32; the fp16 path has no fp32 h1, so this is not in place unless the
// mask itself is the fp32 h1 of a conv path that wrote one)
static int tc16_dgrad(const Workspace& w, const Planes16& pl, int S, int rows, bool have_bits, float gscale,
                      cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], pl.dz_hi, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[1], pl.dz_lo, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[2], pl.w1_hi, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[3], pl.w1_lo, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 128))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = FLAT_CNN / 128; gs.k_blocks = HID_CNN / tc::TC_BK16;
  gs.split3 = 1;
  tc::EpiParams ep = {};
  ep.out = w.h1; ep.mask = w.h1; ep.ld_out = FLAT_CNN; ep.out_seed_stride = (int64_t)rows * FLAT_CNN;
  ep.relu_bits = w.relu_bits; ep.rows = rows; ep.out_scale = 1.0f / gscale;
  return tc::launch_gemm16(0, 0, have_bits ? tc::EPI_RELU_BITS : tc::EPI_RELU_MASK, t, gs, ep, st, K_TC_DGRAD);
}

// ---- generic fp16-split GEMMs on planes (hi at p, lo' at p + plane_elems), used by the MLP hidden layer --------------
// out[S][rows][N] = A[S][rows][K] . B[S][K][N]          (A K-major, B MN-major; raw store, no bias)
static int tc16_mm_store(const __half* a, int64_t a_plane, const __half* b, int64_t b_plane, float* out, int S, int rows, int K,
                         int N, cudaStream_t st, int kid) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], a, K, rows, S, K, (uint64_t)rows * K, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[1], a + a_plane, K, rows, S, K, (uint64_t)rows * K, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[2], b, N, K, S, N, (uint64_t)K * N, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[3], b + b_plane, N, K, S, N, (uint64_t)K * N, 64))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = N / 128; gs.k_blocks = (K + tc::TC_BK16 - 1) / tc::TC_BK16;
  gs.split3 = 1;
  tc::EpiParams ep = {};
  ep.out = out; ep.ld_out = N; ep.out_seed_stride = (int64_t)rows * N;
  return tc::launch_gemm16(0, 1, tc::EPI_STORE, t, gs, ep, st, kid);
}
// dW[S(P)][M][N] = A[S][rows][M]^T . DZ[S][rows][N] * out_scale      (both operands MN-major, K = rows)
static int tc16_mm_wgrad(const __half* a, int64_t a_plane, const __half* dz, int64_t dz_plane, float* out, int64_t out_seed_stride,
                         int S, int rows, int M, int N, float out_scale, float* wg_part, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], a, M, rows, S, M, (uint64_t)rows * M, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[1], a + a_plane, M, rows, S, M, (uint64_t)rows * M, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[2], dz, N, rows, S, N, (uint64_t)rows * N, 64))) return rc;
  if ((rc = tc::make_tmap16(&t[3], dz + dz_plane, N, rows, S, N, (uint64_t)rows * N, 64))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = M; gs.m_tiles = M / 128; gs.n_tiles = N / 128; gs.k_blocks = (rows + tc::TC_BK16 - 1) / tc::TC_BK16; gs.split3 = 1;
  gs.k_split = wgrad_ksplit(S * gs.m_tiles * gs.n_tiles, gs.k_blocks);
  tc::EpiParams ep = {};
  ep.ld_out = N; ep.out_scale = out_scale;
  const int64_t n = (int64_t)M * N;
  if (gs.k_split > 1) { ep.out = wg_part; ep.out_seed_stride = n; ep.split_stride = (int64_t)S * n; }
  else { ep.out = out; ep.out_seed_stride = out_seed_stride; }
  if ((rc = tc::launch_gemm16(1, 1, tc::EPI_STORE, t, gs, ep, st, K_TC_WGRAD))) return rc;
  if (gs.k_split > 1) {
    launch_split_reduce(wg_part, gs.k_split, (int64_t)S * n, n, S, out, out_seed_stride, st);
  }
  return 0;
}
// out[S][rows][Kp] = (mask > 0) * (DZ[S][rows][N] . W[S][Kp][N]^T) * out_scale       (both K-major, K = N)
static int tc16_mm_dgrad(const __half* dz, int64_t dz_plane, const __half* wgt, int64_t w_plane, const float* mask, float* out,
                         int S, int rows, int N, int Kp, float out_scale, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap16(&t[0], dz, N, rows, S, N, (uint64_t)rows * N, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[1], dz + dz_plane, N, rows, S, N, (uint64_t)rows * N, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[2], wgt, N, Kp, S, N, (uint64_t)Kp * N, 128))) return rc;
  if ((rc = tc::make_tmap16(&t[3], wgt + w_plane, N, Kp, S, N, (uint64_t)Kp * N, 128))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = Kp / 128; gs.k_blocks = (N + tc::TC_BK16 - 1) / tc::TC_BK16;
  gs.split3 = 1;
  tc::EpiParams ep = {};
  ep.out = out; ep.mask = mask; ep.ld_out = Kp; ep.out_seed_stride = (int64_t)rows * Kp; ep.rows = rows; ep.out_scale = out_scale;
  return tc::launch_gemm16(0, 0, tc::EPI_RELU_MASK, t, gs, ep, st, K_TC_DGRAD);
}
static void split16_rows(const float* src, int64_t src_seed_stride, int64_t n_per_seed, int S, __half* hi, __half* lo,
                         cudaStream_t st) {
  LaunchScope _ls(K_TC_SPLIT, st);
  split16_strided_kernel<<<dim3(cdiv(n_per_seed / 4, 256), S), 256, 0, st>>>(src, src_seed_stride, hi, lo, n_per_seed);
}

// Z = H1 . W1 on the tcgen05 path with the LayerNorm/ReLU(/head) epilogue.  epi = EPI_LN_TRAIN or EPI_LN_HEAD.
static int tc_dense_fwd(int epi, const float* params, int64_t P, const pqn_net_layout_t& L, const Workspace& w, int A,
                        float* q, int S, int rows, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap(&t[0], w.h1, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 128, 0))) return rc;
  if ((rc = tc::make_tmap(&t[1], w.h1_lo, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 128, 0))) return rc;
  if ((rc = tc::make_tmap(&t[2], params + L.d0_w, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)P, 32, 1))) return rc;
  if ((rc = tc::make_tmap(&t[3], w.w1_lo, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 32, 1))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = 1; gs.k_blocks = FLAT_CNN / tc::TC_BK; gs.split3 = 1;
  gs.a_lo_inline = 1;
  tc::EpiParams ep = {};
  ep.params = params; ep.P = P; ep.off_b = L.d0_b; ep.off_scale = L.ln1_scale; ep.off_bias = L.ln1_bias;
  ep.off_hw = L.head_w; ep.off_hb = L.head_b; ep.A = A; ep.rows = rows;
  ep.H = w.h2; ep.XHAT = w.xhat2; ep.RSTD = w.rstd2; ep.Q = q;
  return tc::launch_gemm(0, 1, epi, t, gs, ep, st, epi == tc::EPI_LN_HEAD ? K_TC_FWD_HEAD : K_TC_FWD);
}

// dW1 = H1^T . dZ2  -> grads[d0_w]
static int tc_wgrad(float* grads, int64_t P, const pqn_net_layout_t& L, const Workspace& w, int S, int rows,
                    cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap(&t[0], w.h1, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 32, 1))) return rc;
  if ((rc = tc::make_tmap(&t[1], w.h1_lo, FLAT_CNN, rows, S, FLAT_CNN, (uint64_t)rows * FLAT_CNN, 32, 1))) return rc;
  if ((rc = tc::make_tmap(&t[2], w.dz2, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 32, 1))) return rc;
  if ((rc = tc::make_tmap(&t[3], w.dz2_lo, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 32, 1))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = FLAT_CNN; gs.m_tiles = FLAT_CNN / 128; gs.n_tiles = 1; gs.k_blocks = (rows + tc::TC_BK - 1) / tc::TC_BK;
  gs.split3 = 1;
  gs.a_lo_inline = 1;
  tc::EpiParams ep = {};
  ep.out = grads + L.d0_w; ep.ld_out = HID_CNN; ep.out_seed_stride = P;
  return tc::launch_gemm(1, 1, tc::EPI_STORE, t, gs, ep, st, K_TC_WGRAD);
}

// dY1 = relu_mask(H1) * (dZ2 . W1^T), written in place over H1
static int tc_dgrad(const float* params, int64_t P, const pqn_net_layout_t& L, const Workspace& w, int S, int rows,
                    bool have_bits, cudaStream_t st) {
  CUtensorMap t[4];
  int rc;
  if ((rc = tc::make_tmap(&t[0], w.dz2, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 128, 0))) return rc;
  if ((rc = tc::make_tmap(&t[1], w.dz2_lo, HID_CNN, rows, S, HID_CNN, (uint64_t)rows * HID_CNN, 128, 0))) return rc;
  if ((rc = tc::make_tmap(&t[2], params + L.d0_w, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)P, 128, 0))) return rc;
  if ((rc = tc::make_tmap(&t[3], w.w1_lo, HID_CNN, FLAT_CNN, S, HID_CNN, (uint64_t)FLAT_CNN * HID_CNN, 128, 0))) return rc;
  tc::GemmShape gs = {};
  gs.S = S; gs.M = rows; gs.m_tiles = (rows + 127) / 128; gs.n_tiles = FLAT_CNN / 128; gs.k_blocks = HID_CNN / tc::TC_BK;
  gs.split3 = 1;
  tc::EpiParams ep = {};
  ep.out = w.h1; ep.mask = w.h1; ep.ld_out = FLAT_CNN; ep.out_seed_stride = (int64_t)rows * FLAT_CNN;
  ep.relu_bits = w.relu_bits; ep.rows = rows;
  // the packed mask (16 B per row and tile) replaces re-reading the 2 GB of activations it was derived from
  return tc::launch_gemm(0, 0, have_bits ? tc::EPI_RELU_BITS : tc::EPI_RELU_MASK, t, gs, ep, st, K_TC_DGRAD);
}

#include "pqn_norm.cuh"
#include "pqn_rnn.cuh"

static inline bool modular_net(const pqn_net_desc_t* d) { return d->norm_type != PQN_NORM_LAYER || d->norm_input != 0; }

}  // namespace pqn

using namespace pqn;

extern "C" {

int64_t pqn_net_stats_floats(const pqn_net_desc_t* d) {
  if (check_desc(d, "pqn_net_stats_floats")) return -1;
  return nrm::stats_floats(d);
}

int pqn_rnn_step(const pqn_net_desc_t* d, const float* params, float* hs, const float* obs, int64_t obs_rows_per_seed,
                 const uint8_t* last_done, const int32_t* last_action, float* q, int32_t S, int32_t E, void* workspace,
                 void* stream) {
  int rc = check_desc(d, "pqn_rnn_step");
  if (rc) return rc;
  if ((rc = rnn::check_rnn(d, "pqn_rnn_step"))) return rc;
  if (!params || !hs || !obs || !last_done || !last_action || !q || !workspace || S <= 0 || E <= 0 || S > 65535)
    return set_error(PQN_E_INVALID, "pqn_rnn_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  pqn_net_layout_t L;
  make_layout(d, &L);
  rnn::RnnWs w;
  rnn::carve_rnn(d, S, E, (char*)workspace, &w);
  rnn::rnn_trunk(d, L, params, obs, obs_rows_per_seed * d->in_c, S, E, false, w, st);
  if ((rc = rnn::rnn_scan_fwd<false>(d, L, params, last_action, last_done, hs, hs, S, 1, E, w, st))) return rc;
  { LaunchScope _ls(K_RNN_MISC, st);
    nrm::head_fwd_kernel<<<dim3(cdiv(E, 8), S), 256, 0, st>>>(w.y, E, d->hidden, params, L.total, L.head_w, L.head_b,
                                                               d->num_actions, q); }
  return check_launch("pqn_rnn_step");
}

int pqn_rnn_loss_grad(const pqn_net_desc_t* d, const float* params, const float* hs0, const float* obs,
                      const uint8_t* last_done, const int32_t* last_action, const int32_t* action, const float* reward,
                      const uint8_t* done, float* grads, float* loss_sum, float* qsa_sum, int32_t S, int32_t T, int32_t B,
                      float gamma, float lambda, void* workspace, void* stream) {
  int rc = check_desc(d, "pqn_rnn_loss_grad");
  if (rc) return rc;
  if ((rc = rnn::check_rnn(d, "pqn_rnn_loss_grad"))) return rc;
  if (!params || !hs0 || !obs || !last_done || !last_action || !action || !reward || !done || !grads || !loss_sum ||
      !qsa_sum || !workspace || S <= 0 || T < 2 || B <= 0 || B > 1024 || S > 65535)
    return set_error(PQN_E_INVALID, "pqn_rnn_loss_grad: bad argument (T >= 2, B <= 1024)");
  cudaStream_t st = (cudaStream_t)stream;
  pqn_net_layout_t L;
  make_layout(d, &L);
  const int64_t P = L.total;
  const int H = d->hidden, A = d->num_actions, D = d->in_c, rows = T * B;
  rnn::RnnWs w;
  rnn::carve_rnn(d, S, rows, (char*)workspace, &w);
  if (cudaMemsetAsync(grads, 0, (size_t)S * P * sizeof(float), st) != cudaSuccess)
    return check_launch("pqn_rnn_loss_grad(memset)");
  const int64_t gs = (int64_t)S * rows * H;
  // ---- forward over the window
  rnn::rnn_trunk(d, L, params, obs, (int64_t)rows * D, S, rows, true, w, st);
  if ((rc = rnn::rnn_scan_fwd<true>(d, L, params, last_action, last_done, hs0, w.dhl /*scratch carry out*/, S, T, B, w, st)))
    return rc;
  { LaunchScope _ls(K_RNN_MISC, st);
    nrm::head_fwd_kernel<<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.y, rows, H, params, P, L.head_w, L.head_b, A, w.q); }
  // ---- targets, loss, dq
  { LaunchScope _ls(K_RNN_MISC, st);
    const int bt = ((B + 31) / 32) * 32;
    rnn::rnn_targets_kernel<<<S, bt, 2 * bt * sizeof(float), st>>>(w.q, action, reward, done, T, B, A, gamma, lambda, w.dq,
                                                                    loss_sum, qsa_sum); }
  // ---- head backward
  { LaunchScope _ls(K_RNN_MISC, st);
    if (H == 128) rnn::rnn_head_bwd_kernel<128><<<dim3(nrm::RED_BLOCKS, S), 128, 0, st>>>(w.y, w.dq, rows, A, params, P, L.head_w, w.dy, w.part);
    else rnn::rnn_head_bwd_kernel<256><<<dim3(nrm::RED_BLOCKS, S), 256, 0, st>>>(w.y, w.dq, rows, A, params, P, L.head_w, w.dy, w.part); }
  { LaunchScope _ls(K_RNN_MISC, st);
    rnn::rnn_head_bwd_final_kernel<<<S, 256, 0, st>>>(w.part, nrm::RED_BLOCKS, H, A, grads, P, L.head_w, L.head_b); }
  // ---- BPTT through the GRU
  const int64_t wts = (int64_t)S * H * H;
  { LaunchScope _ls(K_RNN_MISC, st);
    rnn::gru_transpose_kernel<<<dim3(32, S, 3), 256, 0, st>>>(params, P, L, H, w.wt, wts); }
  { LaunchScope _ls(K_RNN_SCAN, st);
    const dim3 grid(cdiv(B, rnn::RB), S);
    if (H == 128) rnn::gru_scan_bwd_kernel<128><<<grid, 128, 0, st>>>(w.dy, last_done, w.h0, w.rg, w.zg, w.ng, w.hn, w.wt, wts, w.da, gs, w.dhn, T, B);
    else rnn::gru_scan_bwd_kernel<256><<<grid, 256, 0, st>>>(w.dy, last_done, w.h0, w.rg, w.zg, w.ng, w.hn, w.wt, wts, w.da, gs, w.dhn, T, B); }
  // ---- weight gradients of the GRU (batched over the window)
  const float* xl = w.h[d->layers - 1];   // trunk output = first H columns of the GRU input
  const int64_t iw[3] = {L.gru_ir_w, L.gru_iz_w, L.gru_in_w}, ib[3] = {L.gru_ir_b, L.gru_iz_b, L.gru_in_b};
  const int64_t hw[3] = {L.gru_hr_w, L.gru_hz_w, L.gru_hn_w};
  const int tiles = (H / 128) * (H / 128);
  const int sp = wgrad_splits(tiles, S, rows);
  nrm::NormWs nw = {};
  nw.part = w.part;
  for (int g = 0; g < 3; ++g) {
    const float* da = w.da + g * gs;
    run_wgrad_ffma(xl, (int64_t)rows * H, H, da, (int64_t)rows * H, H, grads, P, iw[g], rows, H, S, sp, w.wgp, st);
    nrm::colsum2(da, da, S, rows, H, H, nw, w.sums, grads, P, ib[g], -1, st);                         // d b_ig
    const float* dh = g == 2 ? w.dhn : da;                                                             // hn uses d(hn)
    run_wgrad_ffma(w.h0, (int64_t)rows * H, H, dh, (int64_t)rows * H, H, grads, P, hw[g], rows, H, S, sp, w.wgp, st);
    // d x_L (+)= da_g W_ig[:H]^T, masked by the trunk's ReLU
    { LaunchScope _ls(K_DGRAD, st); dgrad_kernel<<<dim3(cdiv(rows, 128), H / 128, S), GT, 0, st>>>(da, (int64_t)rows * H, H, params, P, iw[g], xl, w.dx, (int64_t)rows * H, rows, H, g > 0 ? 1 : 0); }
  }
  nrm::colsum2(w.dhn, w.dhn, S, rows, H, H, nw, w.sums, grads, P, L.gru_hn_b, -1, st);                 // d b_hn
  { LaunchScope _ls(K_RNN_MISC, st);
    if (H == 128) rnn::rnn_onehot_grad_kernel<128><<<dim3(S, 3), 128, 0, st>>>(w.da, gs, last_action, rows, A, grads, P, L);
    else rnn::rnn_onehot_grad_kernel<256><<<dim3(S, 3), 256, 0, st>>>(w.da, gs, last_action, rows, A, grads, P, L); }
  // ---- trunk backward (LayerNorm backward -> weight gradient -> input gradient of the layer below)
  const dim3 rbg(conv_mma_ctas(S, rows, 4), S);
  const int64_t offw[2] = {L.d0_w, L.d1_w}, offb[2] = {L.d0_b, L.d1_b}, offg[2] = {L.ln0_scale, L.ln1_scale},
                offbi[2] = {L.ln0_bias, L.ln1_bias};
  float* dcur = w.dx;
  for (int l = d->layers - 1; l >= 0; --l) {
    if ((rc = run_row_bwd(H, false, rbg, A, st, nullptr, w.xh[l], w.rs[l], dcur, dcur, nullptr, nullptr, nullptr, 1.0f, params,
                          grads, P, offg[l], offbi[l], offb[l], 0, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, w.rbp,
                          rows))) return rc;
    const float* xprev = l == 0 ? obs : w.h[l - 1];
    const int kin = l == 0 ? D : H;
    const int spl = wgrad_splits((kin + 127) / 128 * (H / 128), S, rows);
    run_wgrad_ffma(xprev, (int64_t)rows * kin, kin, dcur, (int64_t)rows * H, H, grads, P, offw[l], rows, kin, S, spl, w.wgp, st);
    if (l > 0) {
      { LaunchScope _ls(K_DGRAD, st); dgrad_kernel<<<dim3(cdiv(rows, 128), H / 128, S), GT, 0, st>>>(dcur, (int64_t)rows * H, H, params, P, offw[l], w.h[l - 1], w.dhl, (int64_t)rows * H, rows, H, 0); }
      dcur = w.dhl;
    }
  }
  return check_launch("pqn_rnn_loss_grad");
}

int pqn_set_conv_mma_path(int on) {
  g_conv_mma = on < 0 ? 0 : (on > 3 ? 3 : on);
  return PQN_OK;
}

int pqn_set_tensor_core_path(int on) {
  g_use_tc = on < 0 ? 0 : (on > 2 ? 2 : on);
  return PQN_OK;
}

int pqn_net_layout(const pqn_net_desc_t* d, pqn_net_layout_t* out) {
  int rc = check_desc(d, "pqn_net_layout");
  if (rc) return rc;
  if (!out) return set_error(PQN_E_INVALID, "pqn_net_layout: out is NULL");
  make_layout(d, out);
  return PQN_OK;
}

int64_t pqn_net_workspace_bytes(const pqn_net_desc_t* d, int32_t S, int64_t rows) {
  if (check_desc(d, "pqn_net_workspace_bytes")) return -1;
  if (d->kind == PQN_NET_RNN) return rnn::carve_rnn(d, S, rows, nullptr, nullptr);
  if (modular_net(d)) return nrm::carve_norm(d, S, rows, nullptr, nullptr);
  return carve(d, S, rows, nullptr, nullptr);
}

int pqn_qnet_forward(const pqn_net_desc_t* d, const float* params, const float* batch_stats, const void* obs,
                     const int32_t* gather, int64_t obs_rows_per_seed, float* q, int32_t S, int64_t rows, void* workspace,
                     void* stream) {
  int rc = check_desc(d, "pqn_qnet_forward");
  if (rc) return rc;
  if (!params || !obs || !q || !workspace || S <= 0 || rows <= 0 || S > 65535)
    return set_error(PQN_E_INVALID, "pqn_qnet_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  pqn_net_layout_t L;
  make_layout(d, &L);
  if (modular_net(d)) {
    nrm::NormWs nw;
    nrm::carve_norm(d, S, rows, (char*)workspace, &nw);
    return nrm::norm_forward(d, L, params, const_cast<float*>(batch_stats), obs, gather, obs_rows_per_seed, q, S, (int)rows,
                            0, nullptr, nw, st);
  }
  Workspace w;
  carve(d, S, rows, (char*)workspace, &w);
  const int A = d->num_actions;
  if (d->kind == PQN_NET_MINATAR_CNN) {
    const bool use_tc = g_use_tc && A <= PQN_TC_MAX_A;
    const bool f16 = use_tc && g_use_tc == 2;
    const Planes16 pl = planes16(w, S, rows);
    const bool conv16 = f16 && g_conv_mma == 1;   // the mma.sync conv writes the fp16 planes itself
    launch_conv_fwd<false>(d->in_c, dim3(cdiv(rows, 4), S), st, (const uint32_t*)obs, obs_rows_per_seed, gather, params,
                           L.total, L, conv16 ? (float*)pl.h1_hi : w.h1, conv16 ? (float*)pl.h1_lo : nullptr, nullptr,
                           (int)rows, nullptr, nullptr, nullptr, conv16);
    if (f16) {
      if (!conv16) launch_split16_h1(w.h1, pl, (int64_t)S * rows * FLAT_CNN, st);
      launch_split16_w1(params, L.total, L.d0_w, pl, S, st);
      if ((rc = tc16_dense_fwd(tc::EPI_LN_HEAD, params, L.total, L, w, pl, A, q, S, (int)rows, st))) return rc;
    } else if (use_tc) {
      launch_split_w1(params, L.total, L.d0_w, w.w1_lo, S, st);
      if ((rc = tc_dense_fwd(tc::EPI_LN_HEAD, params, L.total, L, w, A, q, S, (int)rows, st))) return rc;
    } else {
      launch_dense<2>(128, dim3(cdiv(rows, 128), S), st, w.h1, rows * FLAT_CNN, FLAT_CNN, params, L.total, L.d0_w,
                      L.d0_b, L.ln1_scale, L.ln1_bias, L.head_w, L.head_b, A, nullptr, nullptr, nullptr, q, (int)rows,
                      FLAT_CNN);
    }
  } else {
    const int D = d->in_c, H = d->hidden;
    const float* x = (const float*)obs;
    int64_t xss = obs_rows_per_seed * D;
    if (gather) {
      { LaunchScope _ls(K_GATHER_ROWS, st); gather_rows_kernel<<<dim3(cdiv(rows * D, 256), S), 256, 0, st>>>(x, obs_rows_per_seed, gather, w.xg,
                                                                      (int)rows, D); }
      x = w.xg;
      xss = rows * D;
    }
    const int BM = (H == 128) ? 128 : 64;
    if (d->layers == 1) {
      launch_dense<2>(H, dim3(cdiv(rows, BM), S), st, x, xss, D, params, L.total, L.d0_w, L.d0_b, L.ln0_scale,
                      L.ln0_bias, L.head_w, L.head_b, A, nullptr, nullptr, nullptr, q, (int)rows, D);
    } else {
      launch_dense<0>(H, dim3(cdiv(rows, BM), S), st, x, xss, D, params, L.total, L.d0_w, L.d0_b, L.ln0_scale,
                      L.ln0_bias, 0, 0, A, w.h0, nullptr, nullptr, nullptr, (int)rows, D);
      if (g_use_tc == 2) {
        // hidden layer (K = N = H) on tcgen05: fp16-split planes of h0 and of the Dense_1 kernel, raw product, then
        // bias + LayerNorm + ReLU and the Q head in row kernels
        const int64_t R = (int64_t)S * rows;
        __half* hp = reinterpret_cast<__half*>(w.m16_h0);
        __half* wp = reinterpret_cast<__half*>(w.m16_w);
        split16_rows(w.h0, 0, R * H, 1, hp, hp + R * H, st);
        split16_rows(params + L.d1_w, L.total, (int64_t)H * H, S, wp, wp + (int64_t)S * H * H, st);
        if ((rc = tc16_mm_store(hp, R * H, wp, (int64_t)S * H * H, w.hh1, S, (int)rows, H, H, st, K_TC_FWD_HEAD))) return rc;
        { LaunchScope _ls(K_NORM_FWD, st);
          if (H == 128) nrm::ln_fwd_kernel<128><<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.hh1, rows, params, L.total, L.ln1_scale, L.ln1_bias, nullptr, nullptr, w.hh1, L.d1_b);
          else nrm::ln_fwd_kernel<256><<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.hh1, rows, params, L.total, L.ln1_scale, L.ln1_bias, nullptr, nullptr, w.hh1, L.d1_b); }
        { LaunchScope _ls(K_NORM_FWD, st);
          nrm::head_fwd_kernel<<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.hh1, (int)rows, H, params, L.total, L.head_w, L.head_b, A, q); }
      } else
      launch_dense<2>(H, dim3(cdiv(rows, BM), S), st, w.h0, rows * H, H, params, L.total, L.d1_w, L.d1_b, L.ln1_scale,
                      L.ln1_bias, L.head_w, L.head_b, A, nullptr, nullptr, nullptr, q, (int)rows, H);
    }
  }
  return check_launch("pqn_qnet_forward");
}

int pqn_qnet_loss_grad(const pqn_net_desc_t* d, const float* params, float* batch_stats, const void* obs,
                       const int32_t* gather, int64_t obs_rows_per_seed, const int32_t* action, const float* target,
                       int64_t tr_rows_per_seed, float* grads, float* loss_sum, float* qsa_sum, float* bn_sums,
                       int32_t S, int64_t rows, void* workspace, void* stream) {
  int rc = check_desc(d, "pqn_qnet_loss_grad");
  if (rc) return rc;
  if (!params || !obs || !action || !target || !grads || !loss_sum || !qsa_sum || !workspace || S <= 0 || rows <= 0 ||
      S > 65535)
    return set_error(PQN_E_INVALID, "pqn_qnet_loss_grad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  pqn_net_layout_t L;
  make_layout(d, &L);
  const int64_t P = L.total;
  Workspace w;
  carve(d, S, rows, (char*)workspace, &w);
  const int A = d->num_actions;
  const int R = (int)rows;
  if (cudaMemsetAsync(grads, 0, (size_t)S * P * sizeof(float), st) != cudaSuccess)
    return check_launch("pqn_qnet_loss_grad(memset)");
  if (modular_net(d)) {
    nrm::NormWs nw;
    nrm::carve_norm(d, S, rows, (char*)workspace, &nw);
    return nrm::norm_loss_grad(d, L, params, batch_stats, obs, gather, obs_rows_per_seed, action, target, tr_rows_per_seed,
                              grads, loss_sum, qsa_sum, bn_sums, S, R, nw, st);
  }

  if (d->kind == PQN_NET_MINATAR_CNN) {
    const uint32_t* ob = (const uint32_t*)obs;
    const bool use_tc = g_use_tc && A <= PQN_TC_MAX_A;
    const bool f16 = use_tc && g_use_tc == 2;
    const Planes16 pl = planes16(w, S, rows);
    const bool conv16 = f16 && g_conv_mma == 1;
    const float gscale = f16 ? grad_scale(rows) : 1.0f;
    launch_conv_fwd<true>(d->in_c, dim3(cdiv(rows, 4), S), st, ob, obs_rows_per_seed, gather, params, P, L,
                          conv16 ? (float*)pl.h1_hi : w.h1, conv16 ? (float*)pl.h1_lo : nullptr, bn_sums, R,
                          g_conv_mma ? w.cxhat : nullptr, g_conv_mma ? w.crstd : nullptr,
                          (g_conv_mma == 1 || g_conv_mma == 3) ? w.relu_bits : nullptr, conv16);
    if (f16) {
      if (!conv16) launch_split16_h1(w.h1, pl, (int64_t)S * rows * FLAT_CNN, st);
      launch_split16_w1(params, P, L.d0_w, pl, S, st);
      if ((rc = tc16_dense_fwd(tc::EPI_LN_TRAIN, params, P, L, w, pl, A, nullptr, S, R, st))) return rc;
    } else if (use_tc) {
      launch_split_w1(params, P, L.d0_w, w.w1_lo, S, st);
      if ((rc = tc_dense_fwd(tc::EPI_LN_TRAIN, params, P, L, w, A, nullptr, S, R, st))) return rc;
    } else {
      launch_dense<1>(128, dim3(cdiv(rows, 128), S), st, w.h1, rows * FLAT_CNN, FLAT_CNN, params, P, L.d0_w, L.d0_b,
                      L.ln1_scale, L.ln1_bias, 0, 0, A, w.h2, w.xhat2, w.rstd2, nullptr, R, FLAT_CNN);
    }
    const dim3 rbg(conv_mma_ctas(S, R, 4), S);
    if ((rc = run_row_bwd(128, true, rbg, A, st, w.h2, w.xhat2, w.rstd2, nullptr, w.dz2, (use_tc && !f16) ? w.dz2_lo : nullptr,
                          f16 ? pl.dz_hi : nullptr, f16 ? pl.dz_lo : nullptr, gscale, params, grads, P, L.ln1_scale,
                          L.ln1_bias, L.d0_b, L.head_w, L.head_b, gather, action, target, tr_rows_per_seed, loss_sum,
                          qsa_sum, w.rb_part, R))) return rc;
    if (f16) {
      if ((rc = tc16_wgrad(grads, P, L, pl, S, R, gscale, w.wg_part, st))) return rc;
      if ((rc = tc16_dgrad(w, pl, S, R, g_conv_mma == 1 || g_conv_mma == 3, gscale, st))) return rc;
    } else if (use_tc) {
      if ((rc = tc_wgrad(grads, P, L, w, S, R, st))) return rc;
      if ((rc = tc_dgrad(params, P, L, w, S, R, g_conv_mma == 1 || g_conv_mma == 3, st))) return rc;
    } else {
      const int splits = wgrad_splits(FLAT_CNN / 128, S, R);
      run_wgrad_ffma(w.h1, rows * FLAT_CNN, FLAT_CNN, w.dz2, rows * HID_CNN, HID_CNN, grads, P, L.d0_w, R, FLAT_CNN, S, splits, w.wg_part, st);
      { LaunchScope _ls(K_DGRAD, st); dgrad_kernel<<<dim3(cdiv(rows, 128), FLAT_CNN / 128, S), GT, 0, st>>>(w.dz2, rows * HID_CNN, HID_CNN, params, P,
                                                                            L.d0_w, w.h1, w.h1, rows * FLAT_CNN, R,
                                                                            FLAT_CNN); }
    }
    dim3 cg(conv_bwd_ctas(S, R), S);
    if (g_conv_mma) {
      const dim3 mg(conv_mma_ctas(S, R, 2), S);
      LaunchScope _ls(K_CONV_BWD, st);
      // dz of the conv is rstd (<= 316) times a dense-layer-sized gradient: one sixteenth of the dense gradient scale
      // keeps it a factor ~200 below the fp16 maximum (conversions saturate) and far above the subnormals
      const float gs16 = grad_scale((int)rows) * (1.0f / 16.0f);
      if (g_conv_mma == 1) {
        switch (d->in_c) {
          case 4: rc = launch_conv_bwd_mma16<4>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R, gs16); break;
          case 6: rc = launch_conv_bwd_mma16<6>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R, gs16); break;
          case 7: rc = launch_conv_bwd_mma16<7>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R, gs16); break;
          case 10: rc = launch_conv_bwd_mma16<10>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R, gs16); break;
        }
      } else
      switch (d->in_c) {
        case 4: rc = launch_conv_bwd_mma<4>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R); break;
        case 6: rc = launch_conv_bwd_mma<6>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R); break;
        case 7: rc = launch_conv_bwd_mma<7>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R); break;
        case 10: rc = launch_conv_bwd_mma<10>(mg, st, ob, obs_rows_per_seed, gather, params, P, L, w.h1, w.cxhat, w.crstd, grads, w.cb_part, R); break;
      }
      if (rc) return rc;
    } else
    switch (d->in_c) {
      case 4: { LaunchScope _ls(K_CONV_BWD, st); conv_bwd_kernel<4><<<cg, 256, 0, st>>>(ob, obs_rows_per_seed, gather, params, P, L, w.h1, grads, R); } break;
      case 6: { LaunchScope _ls(K_CONV_BWD, st); conv_bwd_kernel<6><<<cg, 256, 0, st>>>(ob, obs_rows_per_seed, gather, params, P, L, w.h1, grads, R); } break;
      case 7: { LaunchScope _ls(K_CONV_BWD, st); conv_bwd_kernel<7><<<cg, 256, 0, st>>>(ob, obs_rows_per_seed, gather, params, P, L, w.h1, grads, R); } break;
      case 10: { LaunchScope _ls(K_CONV_BWD, st); conv_bwd_kernel<10><<<cg, 256, 0, st>>>(ob, obs_rows_per_seed, gather, params, P, L, w.h1, grads, R); } break;
    }
  } else {
    const int D = d->in_c, H = d->hidden;
    const int BM = (H == 128) ? 128 : 64;
    { LaunchScope _ls(K_GATHER_ROWS, st); gather_rows_kernel<<<dim3(cdiv(rows * D, 256), S), 256, 0, st>>>((const float*)obs, obs_rows_per_seed, gather, w.xg, R, D); }
    if (bn_sums) {
      // input BatchNorm statistics (sum x, sum x^2 per feature): deterministic two-stage column sums of the gathered
      // rows (the per-element float atomics this replaces were 18 % of an Acrobot update at 65,536 envs)
      nrm::NormWs nw = {};
      nw.part = w.rb_part;
      nrm::colsum2(w.xg, w.xg, S, R, D, D, nw, bn_sums, nullptr, 0, -1, -1, st);
    }
    launch_dense<1>(H, dim3(cdiv(rows, BM), S), st, w.xg, rows * D, D, params, P, L.d0_w, L.d0_b, L.ln0_scale,
                    L.ln0_bias, 0, 0, A, w.h0, w.xhat0, w.rstd0, nullptr, R, D);
    const dim3 rbg(conv_mma_ctas(S, R, 4), S);
    if (d->layers == 2 && g_use_tc == 2) {
      // hidden layer on tcgen05 (fp16-split planes): forward product, weight gradient and input gradient
      const int64_t RR = (int64_t)S * rows;
      __half* hp = reinterpret_cast<__half*>(w.m16_h0);
      __half* wp = reinterpret_cast<__half*>(w.m16_w);
      __half* zp = reinterpret_cast<__half*>(w.m16_dz);
      const float gscale = grad_scale(rows);
      split16_rows(w.h0, 0, RR * H, 1, hp, hp + RR * H, st);
      split16_rows(params + L.d1_w, P, (int64_t)H * H, S, wp, wp + (int64_t)S * H * H, st);
      if ((rc = tc16_mm_store(hp, RR * H, wp, (int64_t)S * H * H, w.hh1, S, R, H, H, st, K_TC_FWD))) return rc;
      { LaunchScope _ls(K_NORM_FWD, st);
        if (H == 128) nrm::ln_fwd_kernel<128><<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.hh1, rows, params, P, L.ln1_scale, L.ln1_bias, w.xhat1, w.rstd1, w.hh1, L.d1_b);
        else nrm::ln_fwd_kernel<256><<<dim3(cdiv(rows, 8), S), 256, 0, st>>>(w.hh1, rows, params, P, L.ln1_scale, L.ln1_bias, w.xhat1, w.rstd1, w.hh1, L.d1_b); }
      if ((rc = run_row_bwd(H, true, rbg, A, st, w.hh1, w.xhat1, w.rstd1, nullptr, w.dzl, nullptr, zp, zp + RR * H, gscale, params, grads, P, L.ln1_scale, L.ln1_bias, L.d1_b, L.head_w, L.head_b,
                            gather, action, target, tr_rows_per_seed, loss_sum, qsa_sum, w.rb_part, R))) return rc;
      if ((rc = tc16_mm_wgrad(hp, RR * H, zp, RR * H, grads + L.d1_w, P, S, R, H, H, 1.0f / gscale, w.wg_part, st))) return rc;
      if ((rc = tc16_mm_dgrad(zp, RR * H, wp, (int64_t)S * H * H, w.h0, w.dh0, S, R, H, H, 1.0f / gscale, st))) return rc;
      if ((rc = run_row_bwd(H, false, rbg, A, st, nullptr, w.xhat0, w.rstd0, w.dh0, w.dh0, nullptr, nullptr, nullptr, 1.0f, params, grads, P, L.ln0_scale, L.ln0_bias, L.d0_b, 0, 0,
                            nullptr, nullptr, nullptr, 0, nullptr, nullptr, w.rb_part, R))) return rc;
      run_wgrad_first(w.xg, w.dh0, grads, P, L.d0_w, S, R, D, H, w.rb_part, w.wg_part, st);
    } else if (d->layers == 2) {
      launch_dense<1>(H, dim3(cdiv(rows, BM), S), st, w.h0, rows * H, H, params, P, L.d1_w, L.d1_b, L.ln1_scale,
                      L.ln1_bias, 0, 0, A, w.hh1, w.xhat1, w.rstd1, nullptr, R, H);
      if ((rc = run_row_bwd(H, true, rbg, A, st, w.hh1, w.xhat1, w.rstd1, nullptr, w.dzl, nullptr, nullptr, nullptr, 1.0f, params, grads, P, L.ln1_scale, L.ln1_bias, L.d1_b, L.head_w, L.head_b,
                            gather, action, target, tr_rows_per_seed, loss_sum, qsa_sum, w.rb_part, R))) return rc;
      const int tiles = (H / 128) * (H / 128);
      const int splits = wgrad_splits(tiles, S, R);
      run_wgrad_ffma(w.h0, rows * H, H, w.dzl, rows * H, H, grads, P, L.d1_w, R, H, S, splits, w.wg_part, st);
      { LaunchScope _ls(K_DGRAD, st); dgrad_kernel<<<dim3(cdiv(rows, 128), H / 128, S), GT, 0, st>>>(w.dzl, rows * H, H, params, P, L.d1_w, w.h0,
                                                                     w.dh0, rows * H, R, H); }
      if ((rc = run_row_bwd(H, false, rbg, A, st, nullptr, w.xhat0, w.rstd0, w.dh0, w.dh0, nullptr, nullptr, nullptr, 1.0f, params, grads, P, L.ln0_scale, L.ln0_bias, L.d0_b, 0, 0,
                            nullptr, nullptr, nullptr, 0, nullptr, nullptr, w.rb_part, R))) return rc;
      run_wgrad_first(w.xg, w.dh0, grads, P, L.d0_w, S, R, D, H, w.rb_part, w.wg_part, st);
    } else {
      if ((rc = run_row_bwd(H, true, rbg, A, st, w.h0, w.xhat0, w.rstd0, nullptr, w.dzl, nullptr, nullptr, nullptr, 1.0f, params, grads, P, L.ln0_scale, L.ln0_bias, L.d0_b, L.head_w, L.head_b,
                            gather, action, target, tr_rows_per_seed, loss_sum, qsa_sum, w.rb_part, R))) return rc;
      run_wgrad_first(w.xg, w.dzl, grads, P, L.d0_w, S, R, D, H, w.rb_part, w.wg_part, st);
    }
  }
  return check_launch("pqn_qnet_loss_grad");
}

}  // extern "C"
