// PTX wrappers and constants of the tcgen05 / TMEM / TMA GEMM path (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pqn {
namespace tc {

constexpr int TC_THREADS = 256;          // warp0 TMA, warp1 MMA, warps 2-5 epilogue (warp2 also owns TMEM alloc),
                                         // warps 6-7 operand converters (in-kernel A_lo, see GemmShape::a_lo_inline)
constexpr int TC_CONV_THREADS = 64;
constexpr int TC_BK = 32;                // fp32 (tf32 path) elements per k-block = one 128-byte swizzle row
constexpr int TC_BK16 = 64;              // fp16 (f16 path) elements per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_PROMOTE = 4;             // k-blocks per in-TMEM main chain (16 MMAs) before promotion to registers
constexpr int TC_TILE_BYTES = 128 * TC_BK * 4;   // 16 KB: 128 rows x 128 bytes for either element type
// fp16 split: x = hi + lo' * 2^-11 with hi = fp16(x), lo' = fp16((x - hi) * 2^11): 22 significant bits, the scaled lo'
// stays a normal fp16 number down to |x| ~ 1e-7.  The two cross products accumulate in the `corr` TMEM accumulator
// in units of 2^-11.
constexpr float TC_LO_SCALE = 2048.0f, TC_LO_INV = 1.0f / 2048.0f;
constexpr int TC_A_HI = 0, TC_A_LO = TC_TILE_BYTES, TC_B_HI = 2 * TC_TILE_BYTES, TC_B_LO = 3 * TC_TILE_BYTES;
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;  // 64 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                              4 * 32 * 36 * 4 /*epilogue store staging, 4 warps x [32][36] floats*/ +
                              (384 + 8 * 128 + 8) * 4 /*per-seed epilogue parameters (LN scale/bias, Q-head)*/;
#define PQN_TC_MAX_A 8

enum Epilogue : int { EPI_STORE = 0, EPI_LN_TRAIN = 1, EPI_LN_HEAD = 2, EPI_RELU_MASK = 3, EPI_RELU_BITS = 4 };

struct GemmShape {
  int S;         // batch (seeds)
  int M;         // rows of D that exist (rows >= M are not stored)
  int m_tiles, n_tiles, k_blocks;
  int split3;    // 1: split precision (3 tensor-core products per k-step), 0: single pass
  int k_split;   // >1: the k-blocks of every output tile are divided over k_split CTAs; partial ks goes to
                 // out + ks * EpiParams::split_stride (EPI_STORE only) and a reduce kernel adds them in order
  int a_lo_inline;  // split3 only: 1 = no A_lo tensor in memory; the converter warps derive A_lo = A - trunc_tf32(A)
                    // from the A tile TMA staged in shared memory (halves the HBM traffic of the A operand)
};

struct EpiParams {
  int64_t split_stride;  // elements between the partial outputs of a split-K launch
  float out_scale;  // F16 kernels: result = (main + corr * 2^-11) * out_scale (undoes the operand pre-scaling); 0 => 1
  // EPI_STORE / EPI_RELU_MASK
  float* out;
  const float* mask;
  const uint32_t* relu_bits;  // EPI_RELU_BITS: [S][rows][ld_out / 32] words, bit c of a row = (activation c > 0)
  int64_t ld_out, out_seed_stride;
  // EPI_LN_*
  const float* params;
  int64_t P, off_b, off_scale, off_bias, off_hw, off_hb;
  int A, rows;
  float *H, *XHAT, *RSTD, *Q;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_lo(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // what the tensor core reads of x
  return x - hi;                                                       // exact in fp32
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bits 4-5 = 1), a/b format TF32 (= 2,
// bits 7-9 / 10-12), a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with fp16 operands (a/b format F16 = 0), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) for a 128 x 32 fp32 operand tile in SWIZZLE_128B
// atoms (8 rows x 128 B = 1024 B), k-step `ks` selects 8 of the 32 k values:
//   K-major  tile: row r (MN index) at r*128 B; 8-row atoms every 1024 B (SBO); the k-step advances 32 B in the row.
//   MN-major tile (32-bit types need the 32-byte-base swizzle, LayoutType SWIZZLE_128B_BASE32B = 1, written by TMA's
//                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4 boxes of [32 k rows][32 mn] at 4096 B (LBO between MN
//                  atoms); atoms are [4 k][128 B] = 512 B (SBO between k atoms); one k-step (8 k) = 1024 B.
template <int MN>
__device__ __forceinline__ uint64_t make_sdesc(uint32_t tile_addr, int ks) {
  const uint32_t addr = MN ? tile_addr + ks * 1024 : tile_addr + ks * 32;
  const uint64_t lbo = MN ? (4096u >> 4) : 1u;
  const uint64_t sbo = MN ? (512u >> 4) : (1024u >> 4);
  const uint64_t layout = MN ? 1ull : 2ull;
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) /*version*/ | (layout << 61);
}

// The same for a 128 x 64 fp16 operand tile (k-step = 16 k values):
//   K-major : identical byte geometry (128-byte rows, 8-row atoms, 32 B per k-step);
//   MN-major: the ordinary SWIZZLE_128B canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 2 TMA boxes
//             of [64 k rows][64 mn = 128 B] at 8192 B (LBO between MN atoms), k atoms of 8 rows = 1024 B (SBO); one
//             k-step (16 k rows) = 2048 B.
template <int MN>
__device__ __forceinline__ uint64_t make_sdesc16(uint32_t tile_addr, int ks) {
  const uint32_t addr = MN ? tile_addr + ks * 2048 : tile_addr + ks * 32;
  const uint64_t lbo = MN ? (8192u >> 4) : 1u;
  const uint64_t sbo = 1024u >> 4;
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) /*version*/ | (2ull << 61);
}

// fp16 split of an fp32 value (see TC_LO_SCALE); saturates instead of overflowing to inf
__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -65000.0f), 65000.0f);
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * TC_LO_SCALE);
}
// SAT = false: the caller guarantees |x| < 65504 (e.g. LayerNorm outputs: |xhat| <= sqrt(n - 1))
template <bool SAT = true>
__device__ __forceinline__ void split16x2(float x0, float x1, __half2& hi, __half2& lo) {
  if (SAT) {
    x0 = fminf(fmaxf(x0, -65000.0f), 65000.0f);
    x1 = fminf(fmaxf(x1, -65000.0f), 65000.0f);
  }
  hi = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn((x0 - hf.x) * TC_LO_SCALE, (x1 - hf.y) * TC_LO_SCALE);
}

// host-side pieces used by other translation units (pqn_net.cu)
int make_tmap(CUtensorMap* tm, const float* base, uint64_t inner, uint64_t mid, uint64_t seeds, uint64_t mid_stride_elems,
              uint64_t seed_stride_elems, uint32_t box_mid, int mn_major);
// fp16 tensor [seeds][mid][inner] with a {64, box_mid, 1} box (128 bytes x box_mid), SWIZZLE_128B for both majors
int make_tmap16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t mid, uint64_t seeds, uint64_t mid_stride_elems,
                uint64_t seed_stride_elems, uint32_t box_mid);
int launch_gemm(int a_mn, int b_mn, int epi, const CUtensorMap* t, const GemmShape& gs, const EpiParams& ep, cudaStream_t st,
                int kernel_id = -1);
// the fp16-split kernel: t = {A_hi, A_lo', B_hi, B_lo'} fp16 maps, gs.k_blocks counts 64-element k-blocks
int launch_gemm16(int a_mn, int b_mn, int epi, const CUtensorMap* t, const GemmShape& gs, const EpiParams& ep,
                  cudaStream_t st, int kernel_id = -1);

}  // namespace tc
}  // namespace pqn
