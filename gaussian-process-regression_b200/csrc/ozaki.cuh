// ozaki.cuh -- the variance pass v = L^-1 K_star on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators).
//
// FP64 has no tcgen05 kind on sm_100a: the DMMA path (gemm.cuh) is pinned at 37 TFLOP/s, and the variance pass
// (n^2 m flops, R/GPRclass.R:162-164) already runs at that ceiling.  The only way past it is to do the O(n^2 m)
// product of the blocked substitution,  R_i = Ks_i - L[i, <i] V[<i],  in exact integer arithmetic on the 4.5 POP/s
// INT8 pipe (Ozaki splitting) and keep FP64 for the O(n m 128) diagonal solves:
//
//   * every operand entry is written as a fixed-point number with S signed 8-bit digits,
//       x = 2^(e - 8S + 2) * sum_t d_t 256^(S-1-t),   d_t in [-128, 127],   |x| < 2^e,
//     e = per-row exponent for L (from the row maximum), per-test-point exponent for V (from |v| <= sqrt(k**),
//     which holds because the predictive variance k** - v'v is non-negative);
//   * digit products d_a d_b are exact in int32 and are accumulated over k by the tensor core; all pairs with the
//     same order o = a + b share one TMEM accumulator (S accumulators of 128 x 64 int32 = S * 64 TMEM columns);
//     pairs with a + b >= S are dropped (they sit below 2^(-8S+2) of the leading term);
//   * every KC = 16384 k-values (int32 headroom: 7 pairs * 16384 * 2^14 < 2^31; 8192 for S = 8) the accumulators are drained:
//       R -= 2^(eL-6) 2^(eV-6) sum_o acc_o 256^-o     (Horner in FP64, one rounding per term).
//   S = 7 carries 54 bits below the row / column maximum: |var - var_fp64| ~ 1e-14 at n = 1024 (tools/oz_sim.py).
//
// Data layout: digits are stored pre-tiled in the canonical no-swizzle K-major UMMA layout, so that one k-step of one
// tile (all S digit planes) is ONE contiguous block in global memory and one TMA bulk copy into shared memory:
//   tile(rows R, 32 k):  byte offset of (r, k) = ((r / 8) * 2 + k / 16) * 128 + (r % 8) * 16 + k % 16
//   (8 x 16 B core matrices; LBO = 128 B between the two k halves, SBO = 256 B between 8-row groups)
//   Ls: [row block i][k block kb][digit s][128 x 32]      Vs: [64-point tile tc][k block kb][digit s][64 x 32]
//
// Kernel structure (one CTA per 128 L-rows x 64 test points, 192 threads):
//   warp 0   producer: two bulk copies per k-step into a ring of shared-memory stages (full/empty mbarriers)
//   warp 1   allocates TMEM; one lane issues the MMAs of a k-step, tcgen05.commit releases the stage
//   warps 2-5 drain TMEM (tcgen05.ld 32x32b), combine the orders, scale and update R in place.
// Kernels in this file, newest first (all produce the same digits; GPRC_OPT_INT8_TILE selects):
//   update_stack_kernel<S, 2>  DEFAULT.  Stacked digit planes (one MMA = plane a of L x up to four planes of V, N <= 256:
//                              10 MMAs per k-step for S = 7 instead of 28) on clusters of two CTAs that share every V
//                              stage through TMA multicast (two block rows of L per launch).  91.7 % of the INT8 pipe.
//   update_stack_kernel<S, 1>  the same without clusters (one block row per launch; also the odd last block row).
//   update_kernel<S, TS>       round 1: one MMA per digit pair (TS: the L tiles staged in tensor memory -- tools/oz_test only).
//   update128_kernel<S>        round 1: 128 x 128 tiles, the orders in two passes.
#pragma once
#include "common.cuh"

namespace gprc {
namespace oz {

constexpr int BM = 128;    // L rows per tile = one block row of the substitution
constexpr int BN = 64;     // test points per tile
constexpr int BK = 32;     // k per stage = K of one tcgen05.mma kind::i8
constexpr int A_TILE = BM * BK;  // 4096 B
constexpr int B_TILE = BN * BK;  // 2048 B
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int DRAIN_STAGING_BYTES = 4 * 32 * 9 * 8;  // per drain warp: 32 rows x 8 columns of FP64, row pitch 9

template <int S>
struct Cfg {
  static_assert(S >= 1 && S <= 8, "1..8 digits");
  static constexpr int STAGE_BYTES = S * (A_TILE + B_TILE);
  static constexpr int STAGES = (217 * 1024) / STAGE_BYTES > 8 ? 8 : (217 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + DRAIN_STAGING_BYTES + 256;
  // drain interval: the order with the most pairs (S of them) must stay below 2^31: S * KC * 2^14 < 2^31
  static constexpr int KC = (S <= 7) ? 16384 : 8192;
};

__host__ __device__ __forceinline__ int tile_off(int r, int k) {
  return ((r >> 3) * 2 + (k >> 4)) * 128 + (r & 7) * 16 + (k & 15);
}

// ---------------------------------------------------------------------------------------------------------------
// digits
// ---------------------------------------------------------------------------------------------------------------
// X = rint(x 2^((8S-2) - e)); balanced digits through the bias trick: Y = X + 0x80..80 (lower S-1 bytes), lower digit
// t = byte_t(Y) - 128, top digit = Y >> 8(S-1).  sum_t digit_t 256^t = X exactly.
template <int S>
__host__ __device__ __forceinline__ long long to_fixed(double x, int e, bool* overflow) {
  const double s = ldexp(x, -e);  // |s| < 1 nominally
  if (!(fabs(s) < 1.9)) {
    *overflow = true;
    return 0;
  }
  return llrint(ldexp(s, 8 * S - 2));
}
template <int S>
__host__ __device__ __forceinline__ long long biased(long long X) {
  long long bias = 0;
#pragma unroll
  for (int t = 0; t < S - 1; ++t) bias |= 0x80LL << (8 * t);
  return X + bias;
}
// digit plane s (0 = most significant) of a biased value
template <int S>
__host__ __device__ __forceinline__ int digit(long long Y, int s) {
  const int t = S - 1 - s;
  if (t == S - 1) return (int)(Y >> (8 * t));
  return (int)((Y >> (8 * t)) & 0xFF) - 128;
}

// exponent e with |x| < 2^e (0 for x = 0)
__host__ __device__ __forceinline__ int exponent_of(double mx) {
  if (!(mx > 0.0)) return 0;
  int e;
  frexp(mx, &e);
  return e;
}

// ---------------------------------------------------------------------------------------------------------------
// L: row maxima over the strictly-lower block part, exponents and scales
// ---------------------------------------------------------------------------------------------------------------
// grid = n_pad / 128 row blocks, 256 threads (row = tid % 128, two interleaved k phases)
__global__ void __launch_bounds__(256) rowmax_kernel(const double* __restrict__ L, long ld, int* __restrict__ erow,
                                                     double* __restrict__ scale_row) {
  __shared__ double red[256];
  const int i = blockIdx.x, r = threadIdx.x & 127, ph = threadIdx.x >> 7;
  const double* p = L + (long)i * BM + r;
  const long kend = (long)i * BM;
  double mx = 0.0;
  long k = ph;
#pragma unroll 1
  for (; k + 14 < kend; k += 16) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = fabs(p[(k + 2 * u) * ld]);
#pragma unroll
    for (int u = 0; u < 8; ++u) mx = fmax(mx, v[u]);
  }
  for (; k < kend; k += 2) mx = fmax(mx, fabs(p[k * ld]));
  red[threadIdx.x] = mx;
  __syncthreads();
  if (ph == 0) {
    mx = fmax(red[r], red[r + 128]);
    const int e = exponent_of(mx);
    erow[(long)i * BM + r] = e;
    scale_row[(long)i * BM + r] = ldexp(1.0, e - 6);
  }
}

// grid (4 * nt, nt): blockIdx.y = row block i, blockIdx.x = k block kb (< 4 i); 256 threads = 128 rows x 2 k halves
template <int S>
__global__ void __launch_bounds__(256) split_l_kernel(const double* __restrict__ L, long ld, const int* __restrict__ erow,
                                                      int8_t* __restrict__ Ls, int KB, int* __restrict__ error) {
  const int i = blockIdx.y, kb = blockIdx.x;
  if (kb >= 4 * i) return;
  const int m = threadIdx.x & 127, kh = threadIdx.x >> 7;
  const long row = (long)i * BM + m;
  const int e = erow[row];
  const double* p = L + row + ((long)kb * BK + kh * 16) * ld;
  long long Y[16];
  bool ovf = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) Y[j] = biased<S>(to_fixed<S>(p[(long)j * ld], e, &ovf));
  if (ovf) atomicOr(error, 1);
  int8_t* tile = Ls + ((long)i * KB + kb) * (long)(S * A_TILE) + tile_off(m, kh * 16);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (uint32_t)(digit<S>(Y[q * 4 + b], s) & 0xFF) << (8 * b);
      w[q] = v;
    }
    *reinterpret_cast<uint4*>(tile + s * A_TILE) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// V: per-test-point exponents from k**, digits of a freshly solved block row
// ---------------------------------------------------------------------------------------------------------------
// shrink > 0 (GPRC_OPT_INT8_TEST_SHRINK, tests only) lowers every exponent so that V leaves its range and the overflow
// flag / FP64 redo of the chunk is exercised
__global__ void colscale_kernel(const double* __restrict__ kss, long mcur, long mpad, int* __restrict__ ecol,
                                double* __restrict__ scale_col, int shrink = 0) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= mpad) return;
  int e = 0;
  if (t < mcur) {
    const double b = sqrt(fmax(kss[t], 0.0)) * (1.0 + 1e-9);
    e = exponent_of(b) - shrink;
  }
  ecol[t] = e;
  scale_col[t] = ldexp(1.0, e - 6);
}

// block row i of V (rows k in [128 i, 128 i + 128), T[t + k ldt]) -> digit tiles.  grid (mpad / 64, 4), 128 threads
template <int S>
__global__ void __launch_bounds__(128) split_v_kernel(const double* __restrict__ T, long ldt, int i,
                                                      const int* __restrict__ ecol, int8_t* __restrict__ Vs, int KB,
                                                      int* __restrict__ error) {
  const int tc = blockIdx.x, kq = blockIdx.y;
  const int nn = threadIdx.x & 63, kh = threadIdx.x >> 6;
  const long t = (long)tc * BN + nn;
  const int e = ecol[t];
  const double* p = T + t + ((long)i * BM + kq * BK + kh * 16) * ldt;
  long long Y[16];
  bool ovf = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) Y[j] = biased<S>(to_fixed<S>(p[(long)j * ldt], e, &ovf));
  if (ovf) atomicOr(error, 2);
  int8_t* tile = Vs + ((long)tc * KB + (4 * i + kq)) * (long)(S * B_TILE) + tile_off(nn, kh * 16);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (uint32_t)(digit<S>(Y[q * 4 + b], s) & 0xFF) << (8 * b);
      w[q] = v;
    }
    *reinterpret_cast<uint4*>(tile + s * B_TILE) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// tcgen05 / TMEM wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with A taken from tensor memory (128 lanes x 8 columns hold the 128 x 32 int8 tile): only B is read from shared
// memory by the MMA
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared memory (canonical K-major tile, 128 rows x 32 bytes) -> tensor memory (128 lanes x 256 bits); executes in issue
// order with the tcgen05.mma of the same thread
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// mbarrier arrives once all tcgen05 operations issued so far by this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// this thread's TMEM lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at
// [16,30), SBO >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 at [4,6)), A = B = signed 8 bit (1 at [7,10) and
// [10,13)), both K-major (0 at 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// try_wait with a watchdog: a dependency that never arrives raises the error flag and traps instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_guarded(uint32_t bar, uint32_t parity, int* error, int code) {
  for (long spin = 0; !mbar_try(bar, parity); ++spin) {
    if (spin > (1L << 24)) {  // try_wait itself blocks for a bounded time: 2^24 rounds are many seconds
      atomicOr(error, code);
      __threadfence_system();
      __trap();
    }
  }
}

struct UpdateParams {
  const int8_t* Ls;         // digit tiles of L
  const int8_t* Vs;         // digit tiles of V
  const double* scale_row;  // 2^(eL - 6) per row of L
  const double* scale_col;  // 2^(eV - 6) per test point
  double* T;                // K_star^T, element (test point t, row k) at T[t + k ldt]
  long ldt;
  int i;   // block row
  int KB;  // k blocks per row of tiles = n_pad / 32
  int* error;
  int dbg;            // tools/oz_test only: 1 = no MMA (feed rate), 2 = no TMA (MMA rate), 8 = products grouped by order
  long long* trace;   // tools/oz_test only: clock64 of CTA 0 at [kt][0] producer issues, [1] stage landed, [2] MMAs issued
};

// one lane of a converged warp (the issuing lane of the TMA / tcgen05 instructions); elect.sync keeps the branch
// warp-uniform for the compiler, so descriptors stay in uniform registers (an `if (lane == 0)` region makes every
// UTCIMMA an ELECT / branch loop of its own)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// R_i[128 x 64] -= L[i, <i] V[<i, tile]   one CTA per 64-test-point tile
// TS: the digit tiles of L are first copied shared -> tensor memory (tcgen05.cp, S * 8 columns behind the accumulators)
// and the products take A from there, so that each MMA reads only its 2 KB B tile from shared memory (the SS form
// re-reads the 4 KB A tile for every one of the S (S + 1) / 2 products and is bound by that, profiles/README.md).
template <int S, bool TS>
__global__ void __launch_bounds__(THREADS, 1) update_kernel(const UpdateParams p) {
  static_assert(!TS || S * (BN + 8) <= TMEM_COLS, "no room for the A tiles in tensor memory");
  using C = Cfg<S>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stg_all = reinterpret_cast<double*>(smem_raw + C::STAGES * C::STAGE_BYTES);  // drain staging, 4 x [32][9]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::STAGES * C::STAGE_BYTES + DRAIN_STAGING_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tc = blockIdx.x;
  const int KT = 4 * p.i;                      // k-steps
  constexpr int KT_CHUNK = C::KC / BK;         // k-steps per drain
  const int nchunks = (KT + KT_CHUNK - 1) / KT_CHUNK;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer (whole warp loops, one elected lane issues) =====
    const int8_t* a_src = p.Ls + (long)p.i * p.KB * (long)(S * A_TILE);
    const int8_t* b_src = p.Vs + (long)tc * p.KB * (long)(S * B_TILE);
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % C::STAGES;
      if (kt >= C::STAGES) mbar_wait_guarded(smem_u32(empty + s), ((kt / C::STAGES) - 1) & 1, p.error, 16);
      if (elect_one()) {
        const uint32_t bar = smem_u32(full + s);
        const uint32_t dst = smem_u32(smem_raw + s * C::STAGE_BYTES);
        if (p.trace && tc == 0 && kt < 512) p.trace[kt * 4 + 0] = clock64();
        if (p.dbg & 2) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
          bulk_g2s(dst, a_src + (long)kt * (S * A_TILE), S * A_TILE, bar);
          bulk_g2s(dst + S * A_TILE, b_src + (long)kt * (S * B_TILE), S * B_TILE, bar);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one elected lane waits, issues and commits (the other lanes idle until the final barrier) =====
    // The wait for the NEXT stage sits in the middle of this stage's products: the tensor pipe still has queued MMAs
    // while the issuing lane pays the try_wait / fence latency, so the pipe does not drain between k-steps.
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_i8(BM, BN);
      constexpr int NPAIRS = S * (S + 1) / 2, HALF = NPAIRS / 2;
      int kt = 0;
      for (int c = 0; c < nchunks; ++c) {
        if (c > 0) {
          mbar_wait_guarded(smem_u32(tmem_empty), (c - 1) & 1, p.error, 32);
          tc_fence_after();
        }
        const int kt_end = min(KT, (c + 1) * KT_CHUNK);
        mbar_wait_guarded(smem_u32(full + kt % C::STAGES), (kt / C::STAGES) & 1, p.error, 64);
        tc_fence_after();
        for (bool first = true; kt < kt_end; ++kt, first = false) {
          const int s = kt % C::STAGES;
          if (p.trace && tc == 0 && kt < 512) p.trace[kt * 4 + 1] = clock64();
          if (p.dbg & 1) {
            if (kt + 1 < kt_end) mbar_wait_guarded(smem_u32(full + (kt + 1) % C::STAGES), ((kt + 1) / C::STAGES) & 1, p.error, 64);
            mbar_arrive(smem_u32(empty + s));
          } else {
            const uint32_t a0 = smem_u32(smem_raw + s * C::STAGE_BYTES);
            const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + S * A_TILE, 128, 256);
            const uint32_t a_tmem = tmem_base + (uint32_t)(S * BN);
            if (TS) {
#pragma unroll
              for (int a = 0; a < S; ++a) tmem_cp_128x256b(a_tmem + (uint32_t)(a * 8), ad0 + (uint64_t)(a * (A_TILE >> 4)));
            }
            if (p.dbg & 8) {  // experiment: products grouped by accumulator (order-major)
#pragma unroll
              for (int o = 0; o < S; ++o)
#pragma unroll
                for (int a = 0; a <= o; ++a) {
                  if (TS)
                    mma_i8_ts(tmem_base + (uint32_t)(o * BN), a_tmem + (uint32_t)(a * 8),
                              bd0 + (uint64_t)((o - a) * (B_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
                  else
                    mma_i8(tmem_base + (uint32_t)(o * BN), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                           bd0 + (uint64_t)((o - a) * (B_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
                }
              if (kt + 1 < kt_end) {
                mbar_wait_guarded(smem_u32(full + (kt + 1) % C::STAGES), ((kt + 1) / C::STAGES) & 1, p.error, 64);
                tc_fence_after();
              }
              tc_commit(smem_u32(empty + s));
              if (kt + 1 == kt_end) tc_commit(smem_u32(tmem_full));
              continue;
            }
            int pair = 0;
#pragma unroll
            for (int a = 0; a < S; ++a) {
#pragma unroll
              for (int b = 0; b < S - a; ++b, ++pair) {
                if (pair == HALF && kt + 1 < kt_end) {
                  mbar_wait_guarded(smem_u32(full + (kt + 1) % C::STAGES), ((kt + 1) / C::STAGES) & 1, p.error, 64);
                  tc_fence_after();
                }
                // the first product into accumulator a + b of this chunk is (0, a + b) at the chunk's first k-step
                if (TS)
                  mma_i8_ts(tmem_base + (uint32_t)((a + b) * BN), a_tmem + (uint32_t)(a * 8),
                            bd0 + (uint64_t)(b * (B_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
                else
                  mma_i8(tmem_base + (uint32_t)((a + b) * BN), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                         bd0 + (uint64_t)(b * (B_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
              }
            }
            tc_commit(smem_u32(empty + s));  // stage s may be refilled once these MMAs have read it
          }
          if (kt + 1 == kt_end) tc_commit(smem_u32(tmem_full));  // accumulators of chunk c are complete
          if (p.trace && tc == 0 && kt < 512) p.trace[kt * 4 + 2] = clock64();
        }
      }
    }
    __syncwarp();
  } else {
    // ===== drain: TMEM -> registers -> (transpose through shared memory) -> R in place, coalesced =====
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;       // row of the tile = TMEM lane
    const double sr = p.scale_row[(long)p.i * BM + row];
    const double* sc = p.scale_col + (long)tc * BN;
    double* stg = stg_all + q * (32 * 9);
    // rows of R this lane touches when the warp walks the tile 4 rows x 8 columns at a time
    double* cbase = p.T + (long)tc * BN + ((long)p.i * BM + q * 32 + (lane >> 3)) * p.ldt + (lane & 7);
    {  // pull this tile of R into L2 while the first chunk is being multiplied
      const double* prow = p.T + (long)tc * BN + ((long)p.i * BM + row) * p.ldt;
#pragma unroll
      for (int j = 0; j < 4; ++j) prefetch_l2(prow + j * 16);
    }
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait_guarded(smem_u32(tmem_full), c & 1, p.error, 128);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < BN / 8; ++g) {
        int32_t acc[S][8];
#pragma unroll
        for (int o = 0; o < S; ++o) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN + g * 8), acc[o]);
        double cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = cbase[(long)(it * 4) * p.ldt + g * 8];  // in flight during the combine
        double scv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) scv[j] = __ldg(sc + g * 8 + j);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double h = (double)acc[S - 1][j];
#pragma unroll
          for (int o = S - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          stg[lane * 9 + j] = (sr * scv[j]) * h;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] -= stg[(it * 4 + (lane >> 3)) * 9 + (lane & 7)];
#pragma unroll
        for (int it = 0; it < 8; ++it) cbase[(long)(it * 4) * p.ldt + g * 8] = cur[it];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Wide variant: one CTA per 128 L-rows x 128 test points, the orders in TWO passes over the same k-range.
//
// A 128 x 64 x 32 MMA with both operands in shared memory takes 32 + 16 = 48 clk (4 KB of A + 2 KB of B through the
// 128 B/clk operand port; its math takes 32), a 128 x 128 x 32 MMA takes 64 clk = its math floor (tools/oz_test rate).
// 128 columns leave room for only 4 accumulators in the 512 TMEM columns, so the S orders are split:
//   pass 0: orders [0, S-4)   needs digit planes 0 .. S-5 of both operands   (S = 7: 6 products per k-step, 24 KB)
//   pass 1: orders [S-4, S)   needs all S planes                             (S = 7: 22 products per k-step, 56 KB)
// Each pass runs over a drain interval of k and is drained (R -= ...) before the other starts.  The shared-memory ring
// is re-partitioned per pass (small stages need a deeper ring to cover the feed latency); `pass_done` tells the
// producer that every MMA of the previous pass has read its stages.
// V digit tiles for this kernel are 128 points wide: Vs[128-point tile][k block][digit][128 x 32].
// ---------------------------------------------------------------------------------------------------------------
constexpr int BN2 = 128;
constexpr int B2_TILE = BN2 * BK;  // 4096 B

template <int S>
struct Cfg2 {
  static_assert(S >= 5 && S <= 8, "5..8 digits");
  static constexpr int NLO = S - 4;                         // orders of pass 0
  static constexpr int PAIRS0 = NLO * (NLO + 1) / 2;
  static constexpr int PAIRS1 = S * (S + 1) / 2 - PAIRS0;
  static constexpr int STAGE0 = NLO * (A_TILE + B2_TILE);   // pass 0 loads planes 0 .. NLO-1
  static constexpr int STAGE1 = S * (A_TILE + B2_TILE);
  static constexpr int RING = (215 * 1024) / STAGE1 * STAGE1;
  static constexpr int STAGES1 = RING / STAGE1;
  static constexpr int STAGES0 = (RING / STAGE0) > 8 ? 8 : (RING / STAGE0);
  static constexpr int SMEM_BYTES = RING + DRAIN_STAGING_BYTES + 512;
  static constexpr int KC = (S <= 7) ? 16384 : 8192;
};

// block row i of V -> 128-point digit tiles.  grid (mpad / 128, 4), 256 threads
template <int S>
__global__ void __launch_bounds__(256) split_v128_kernel(const double* __restrict__ T, long ldt, int i,
                                                         const int* __restrict__ ecol, int8_t* __restrict__ Vs, int KB,
                                                         int* __restrict__ error) {
  const int tc = blockIdx.x, kq = blockIdx.y;
  const int nn = threadIdx.x & 127, kh = threadIdx.x >> 7;
  const long t = (long)tc * BN2 + nn;
  const int e = ecol[t];
  const double* p = T + t + ((long)i * BM + kq * BK + kh * 16) * ldt;
  long long Y[16];
  bool ovf = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) Y[j] = biased<S>(to_fixed<S>(p[(long)j * ldt], e, &ovf));
  if (ovf) atomicOr(error, 2);
  int8_t* tile = Vs + ((long)tc * KB + (4 * i + kq)) * (long)(S * B2_TILE) + tile_off(nn, kh * 16);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (uint32_t)(digit<S>(Y[q * 4 + b], s) & 0xFF) << (8 * b);
      w[q] = v;
    }
    *reinterpret_cast<uint4*>(tile + s * B2_TILE) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int S>
__global__ void __launch_bounds__(THREADS, 1) update128_kernel(const UpdateParams p) {
  using C = Cfg2<S>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stg_all = reinterpret_cast<double*>(smem_raw + C::RING);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::RING + DRAIN_STAGING_BYTES);  // 8 + 8: pass 0 / pass 1
  uint64_t* empty = full + 16;
  uint64_t* tmem_full = empty + 16;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* pass_done = tmem_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pass_done + 1);
  static_assert(40 * 8 + 8 <= 512, "barrier block");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tc = blockIdx.x;
  const int KT = 4 * p.i;
  constexpr int KT_CHUNK = C::KC / BK;
  const int nchunks = (KT + KT_CHUNK - 1) / KT_CHUNK;
  // a "round" = (chunk, pass); round r = 2 * chunk + pass

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), 4);
    mbar_init(smem_u32(pass_done), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    const int8_t* a_src = p.Ls + (long)p.i * p.KB * (long)(S * A_TILE);
    const int8_t* b_src = p.Vs + (long)tc * p.KB * (long)(S * B2_TILE);
    int cnt[2] = {0, 0};  // stages produced so far per pass: slot = n % ring size, use index (barrier phase) = n / ring size;
                          // strictly round-robin across the rounds of a pass, mirrored by the consumer
    for (int r = 0; r < 2 * nchunks; ++r) {
      const int pass = r & 1, c = r >> 1;
      const int kt0 = c * KT_CHUNK, kt1 = min(KT, kt0 + KT_CHUNK);
      const int nst = pass ? C::STAGES1 : C::STAGES0, sbytes = pass ? C::STAGE1 : C::STAGE0;
      const int planes = pass ? S : C::NLO;
      // the ring is about to be re-partitioned: every MMA of the previous round must have read its stages
      if (r > 0) mbar_wait_guarded(smem_u32(pass_done), (r - 1) & 1, p.error, 8);
      for (int kt = kt0; kt < kt1; ++kt) {
        const int n = cnt[pass]++, s = n % nst;
        // reuse of slot s inside this round: wait for the release of its previous use (same round only: across rounds
        // pass_done already covers it, and the phase bookkeeping of `empty` continues per pass)
        if (kt - kt0 >= nst) mbar_wait_guarded(smem_u32(empty + pass * 8 + s), ((n / nst) - 1) & 1, p.error, 16);
        if (elect_one()) {
          const uint32_t bar = smem_u32(full + pass * 8 + s);
          const uint32_t dst = smem_u32(smem_raw + s * sbytes);
          mbar_arrive_expect_tx(bar, (uint32_t)sbytes);
          bulk_g2s(dst, a_src + (long)kt * (S * A_TILE), planes * A_TILE, bar);
          bulk_g2s(dst + planes * A_TILE, b_src + (long)kt * (S * B2_TILE), planes * B2_TILE, bar);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_i8(BM, BN2);
      int cnt[2] = {0, 0};
      for (int r = 0; r < 2 * nchunks; ++r) {
        const int pass = r & 1, c = r >> 1;
        const int kt0 = c * KT_CHUNK, kt1 = min(KT, kt0 + KT_CHUNK);
        const int nst = pass ? C::STAGES1 : C::STAGES0, sbytes = pass ? C::STAGE1 : C::STAGE0;
        if (r > 0) {  // accumulators of the previous round drained
          mbar_wait_guarded(smem_u32(tmem_empty), (r - 1) & 1, p.error, 32);
          tc_fence_after();
        }
        for (int kt = kt0; kt < kt1; ++kt) {
          const int n = cnt[pass]++, s = n % nst;
          mbar_wait_guarded(smem_u32(full + pass * 8 + s), (n / nst) & 1, p.error, 64);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem_raw + s * sbytes);
          const bool first = (kt == kt0);
          if (pass == 0) {
            const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + C::NLO * A_TILE, 128, 256);
#pragma unroll
            for (int a = 0; a < C::NLO; ++a)
#pragma unroll
              for (int b = 0; b < C::NLO - a; ++b)
                mma_i8(tmem_base + (uint32_t)((a + b) * BN2), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                       bd0 + (uint64_t)(b * (B2_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
          } else {
            const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + S * A_TILE, 128, 256);
#pragma unroll
            for (int a = 0; a < S; ++a)
#pragma unroll
              for (int b = 0; b < S - a; ++b)
                if (a + b >= C::NLO)  // first product into accumulator a + b - NLO of this round: a = 0
                  mma_i8(tmem_base + (uint32_t)((a + b - C::NLO) * BN2), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                         bd0 + (uint64_t)(b * (B2_TILE >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(empty + pass * 8 + s));
        }
        tc_commit(smem_u32(tmem_full));  // accumulators of this round complete
        tc_commit(smem_u32(pass_done));  // ... and its stages have been read
      }
    }
    __syncwarp();
  } else {
    // ===== drain =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const double sr = p.scale_row[(long)p.i * BM + row];
    const double* sc = p.scale_col + (long)tc * BN2;
    double* stg = stg_all + q * (32 * 9);
    double* cbase = p.T + (long)tc * BN2 + ((long)p.i * BM + q * 32 + (lane >> 3)) * p.ldt + (lane & 7);
    {
      const double* prow = p.T + (long)tc * BN2 + ((long)p.i * BM + row) * p.ldt;
#pragma unroll
      for (int j = 0; j < 8; ++j) prefetch_l2(prow + j * 16);
    }
    constexpr int NHI = 4;
    for (int r = 0; r < 2 * nchunks; ++r) {
      const int pass = r & 1;
      mbar_wait_guarded(smem_u32(tmem_full), r & 1, p.error, 128);
      tc_fence_after();
      // pass 0: sum_{o < NLO} acc_o 256^-o;  pass 1: 256^-NLO sum_{j < 4} acc_{NLO + j} 256^-j
      const double srp = pass ? ldexp(sr, -8 * C::NLO) : sr;
#pragma unroll 1
      for (int g = 0; g < BN2 / 8; ++g) {
        int32_t acc[NHI][8];
        if (pass == 0) {
#pragma unroll
          for (int o = 0; o < C::NLO; ++o)
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN2 + g * 8), acc[o]);
        } else {
#pragma unroll
          for (int o = 0; o < NHI; ++o)
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN2 + g * 8), acc[o]);
        }
        double cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = cbase[(long)(it * 4) * p.ldt + g * 8];
        double scv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) scv[j] = __ldg(sc + g * 8 + j);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double h;
          if (pass == 0) {
            h = (double)acc[C::NLO - 1][j];
#pragma unroll
            for (int o = C::NLO - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          } else {
            h = (double)acc[NHI - 1][j];
#pragma unroll
            for (int o = NHI - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          }
          stg[lane * 9 + j] = (srp * scv[j]) * h;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] -= stg[(it * 4 + (lane >> 3)) * 9 + (lane & 7)];
#pragma unroll
        for (int it = 0; it < 8; ++it) cbase[(long)(it * 4) * p.ldt + g * 8] = cur[it];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Stacked-plane variant (round 2): the same 128 x 64 tile and the same S accumulators, but the products of one digit
// plane of L with SEVERAL consecutive digit planes of V are ONE tcgen05.mma.
//
// The accumulators of orders o = a + b sit side by side in tensor memory (order o at columns 64 o), and the digit
// planes of a V tile sit one after the other in shared memory (64 rows x 32 B each, 8-row groups 256 B apart), so
//     [acc_a | acc_(a+1) | ... | acc_(a+nb-1)]  +=  A_a  x  [B_0 ; B_1 ; ... ; B_(nb-1)]^T
// is a plain 128 x (64 nb) x 32 MMA: same D / B base addresses as the single products, N = 64 nb in the instruction
// descriptor.  With both operands in shared memory a 128 x N x 32 INT8 MMA takes max(N / 2, 32 + N / 4) clk
// (tools/oz_test rate): the 4 KB A tile is fetched once per MMA through the 128 B/clk operand port, which is what held
// the 28 single products (N = 64: 48 clk against a 32 clk math floor) at 54 % of the pipe.  Stacked, S = 7 needs 10
// MMAs per k-step (N = 256 / 192 / 128 / 64) and 912 clk instead of 1344, 98 % of the 896 clk math floor; the A tiles
// cross the operand port 10 instead of 28 times.  Results are bit-identical to update_kernel (same integer sums).
//
// CL = 2: a cluster of two CTAs works on the SAME 64-point tile of V for two consecutive block rows of L
// (rows 2 j and 2 j + 1 against k < 256 j).  Each CTA's producer fetches its own L digits and HALF of the V digits of
// the k-step, multicast into both CTAs' shared memory: every V byte leaves L2 / HBM once per 256 rows instead of once
// per 128.  Each CTA issues its own cta_group::1 MMAs into its own tensor memory; a ring slot is refilled once BOTH
// CTAs' MMAs have read it (tcgen05.commit multicast onto both `empty` barriers, count 2).  The host then runs one small
// FP64 step between the two diagonal solves of the pair (R[2j+1] -= L[2j+1, 2j] V[2j]).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the same shared-memory offset in every CTA of `mask`; each destination's mbarrier (same offset)
// receives the byte count
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// the barrier at this offset in every CTA of `mask` arrives once the MMAs issued so far by this thread have completed
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// planes of V multiplied with plane a of L: b = 0 .. S-1-a, in one MMA if <= 4 planes (N <= 256), else in two halves
__host__ __device__ constexpr int stack_first(int S, int a) { return (S - a) <= 4 ? (S - a) : (S - a + 1) / 2; }
__host__ __device__ constexpr int stack_mmas(int S) {
  int n = 0;
  for (int a = 0; a < S; ++a) n += (S - a) <= 4 ? 1 : 2;
  return n;
}

// CL == 1: p.i = block row i (k < 128 i), grid = mpad / 64.
// CL == 2: p.i = pair index j (block rows 2 j, 2 j + 1, k < 256 j), grid = 2 * (mpad / 64), launched with cluster
// dimension (2, 1, 1).
template <int S, int CL>
__global__ void __launch_bounds__(THREADS, 1) update_stack_kernel(const UpdateParams p) {
  static_assert(CL == 1 || CL == 2, "cluster of 1 or 2 CTAs");
  static_assert(S * BN <= TMEM_COLS, "accumulators must fit the tensor memory");
  static_assert((S * B_TILE / CL) % 16 == 0, "bulk copies are multiples of 16 bytes");
  using C = Cfg<S>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stg_all = reinterpret_cast<double*>(smem_raw + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::STAGES * C::STAGE_BYTES + DRAIN_STAGING_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tmem_full = empty + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int tc = blockIdx.x / CL;
  const int brow = CL * p.i + (int)rank;       // this CTA's block row of L
  const int KT = 4 * CL * p.i;                 // k-steps (the same for every CTA of the cluster)
  constexpr int KT_CHUNK = C::KC / BK;
  const int nchunks = (KT + KT_CHUNK - 1) / KT_CHUNK;
  constexpr uint16_t ALL = (uint16_t)((1u << CL) - 1u);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), CL);
    }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // barrier inits of every CTA are visible before any multicast copy / commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer: own L digits; 1 / CL of the V digits of the k-step, multicast to the cluster =====
    const int8_t* a_src = p.Ls + (long)brow * p.KB * (long)(S * A_TILE);
    constexpr int VPART = S * B_TILE / CL;
    const int8_t* b_src = p.Vs + (long)tc * p.KB * (long)(S * B_TILE) + (long)rank * VPART;
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % C::STAGES;
      if (kt >= C::STAGES) mbar_wait_guarded(smem_u32(empty + s), ((kt / C::STAGES) - 1) & 1, p.error, 16);
      if (elect_one()) {
        const uint32_t bar = smem_u32(full + s);
        const uint32_t dst = smem_u32(smem_raw + s * C::STAGE_BYTES);
        if (p.dbg & 2) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
          bulk_g2s(dst, a_src + (long)kt * (S * A_TILE), S * A_TILE, bar);
          if (CL == 1)
            bulk_g2s(dst + S * A_TILE, b_src + (long)kt * (S * B_TILE), S * B_TILE, bar);
          else
            bulk_g2s_multicast(dst + S * A_TILE + rank * VPART, b_src + (long)kt * (S * B_TILE), VPART, bar, ALL);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      constexpr int NMMA = stack_mmas(S), WAITPOS = NMMA / 2 - 1;
      int kt = 0;
      for (int c = 0; c < nchunks; ++c) {
        if (c > 0) {
          mbar_wait_guarded(smem_u32(tmem_empty), (c - 1) & 1, p.error, 32);
          tc_fence_after();
        }
        const int kt_end = min(KT, (c + 1) * KT_CHUNK);
        mbar_wait_guarded(smem_u32(full + kt % C::STAGES), (kt / C::STAGES) & 1, p.error, 64);
        tc_fence_after();
        for (bool first = true; kt < kt_end; ++kt, first = false) {
          const int s = kt % C::STAGES;
          if (p.dbg & 1) {
            if (kt + 1 < kt_end) mbar_wait_guarded(smem_u32(full + (kt + 1) % C::STAGES), ((kt + 1) / C::STAGES) & 1, p.error, 64);
            if (CL == 1) mbar_arrive(smem_u32(empty + s));
            else tc_commit_multicast(smem_u32(empty + s), ALL);
          } else {
            if ((p.dbg & 64) && !first) {  // experiment: the wait for this stage at the top instead of mid-step
              mbar_wait_guarded(smem_u32(full + s), (kt / C::STAGES) & 1, p.error, 64);
              tc_fence_after();
            }
            const uint32_t a0 = smem_u32(smem_raw + s * C::STAGE_BYTES);
            const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + S * A_TILE, 128, 256);
            const bool midwait = !(p.dbg & 64) && kt + 1 < kt_end;
            // one MMA: plane a of L against planes b0 .. b0 + nb - 1 of V into the accumulators of orders a + b0 ...
            auto piece = [&](int a, int b0, int nb, uint32_t acc) {
              mma_i8(tmem_base + (uint32_t)((a + b0) * BN), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                     bd0 + (uint64_t)(b0 * (B_TILE >> 4)), instr_desc_i8(BM, BN * nb), acc);
            };
            auto wait_next = [&]() {
              // the wait for the NEXT stage sits in the middle of this stage's MMAs: the tensor pipe has queued work
              // while the issuing lane pays the try_wait / fence latency
              mbar_wait_guarded(smem_u32(full + (kt + 1) % C::STAGES), ((kt + 1) / C::STAGES) & 1, p.error, 64);
              tc_fence_after();
            };
            if (first || (p.dbg & 16)) {
              // plane 0 first: accumulator a + b is first written in this chunk by plane a = 0 at the chunk's first k-step
              int issued = 0;
#pragma unroll
              for (int a = 0; a < S; ++a) {
                const int n1 = stack_first(S, a), n2 = (S - a) - n1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int b0 = h ? n1 : 0, nb = h ? n2 : n1;
                  if (nb == 0) continue;
                  if (issued == WAITPOS && midwait) wait_next();
                  piece(a, b0, nb, (!first || a > 0) ? 1u : 0u);
                  ++issued;
                }
              }
            } else {
              // every other k-step: the same MMAs with the small ones first and the widest last, so that the tensor pipe
              // holds the most queued work while the issuing lane crosses the k-step boundary (commit, next descriptors):
              // 994 instead of 1087 clk per k-step (tools/oz_test ... dbg 16 restores plane-0-first everywhere)
              int issued = 0;
#pragma unroll
              for (int a = S - 1; a >= 0; --a) {
                const int n1 = stack_first(S, a), n2 = (S - a) - n1;
#pragma unroll
                for (int h = 1; h >= 0; --h) {
                  const int b0 = h ? n1 : 0, nb = h ? n2 : n1;
                  if (nb == 0) continue;
                  if (issued == NMMA - 1 - WAITPOS && midwait) wait_next();
                  piece(a, b0, nb, 1u);
                  ++issued;
                }
              }
            }
            if (CL == 1) tc_commit(smem_u32(empty + s));
            else tc_commit_multicast(smem_u32(empty + s), ALL);  // the slot is free once every CTA's MMAs have read it
          }
          if (kt + 1 == kt_end) tc_commit(smem_u32(tmem_full));
        }
      }
    }
    __syncwarp();
  } else {
    // ===== drain (as update_kernel): TMEM -> registers -> transpose through shared memory -> R in place =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const double sr = p.scale_row[(long)brow * BM + row];
    const double* sc = p.scale_col + (long)tc * BN;
    double* stg = stg_all + q * (32 * 9);
    double* cbase = p.T + (long)tc * BN + ((long)brow * BM + q * 32 + (lane >> 3)) * p.ldt + (lane & 7);
    {
      const double* prow = p.T + (long)tc * BN + ((long)brow * BM + row) * p.ldt;
#pragma unroll
      for (int j = 0; j < 4; ++j) prefetch_l2(prow + j * 16);
    }
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait_guarded(smem_u32(tmem_full), c & 1, p.error, 128);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < BN / 8; ++g) {
        int32_t acc[S][8];
#pragma unroll
        for (int o = 0; o < S; ++o) tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN + g * 8), acc[o]);
        double cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = cbase[(long)(it * 4) * p.ldt + g * 8];
        double scv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) scv[j] = __ldg(sc + g * 8 + j);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double h = (double)acc[S - 1][j];
#pragma unroll
          for (int o = S - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          stg[lane * 9 + j] = (sr * scv[j]) * h;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] -= stg[(it * 4 + (lane >> 3)) * 9 + (lane & 7)];
#pragma unroll
        for (int it = 0; it < 8; ++it) cbase[(long)(it * 4) * p.ldt + g * 8] = cur[it];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into its shared memory / barriers
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// tensor-pipe rate probe (bench.py's measured INT8 peak, tools/oz_test rate): one CTA per SM issues `count` int8 MMAs of
// shape 128 x N x 32 from (uninitialised) shared memory, rotating over `rot` accumulators and `nsrc` operand tiles;
// out[0] = clocks CTA 0 took
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int count, int rot, int nsrc, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = instr_desc_i8(128, N);
    const uint32_t s0 = smem_u32(sm);
    const long long t0 = clock64();
    // 8 products per trip, descriptors and accumulator addresses precomputed: the loop must not be issue-bound
    uint64_t ad[8], bd[8];
    uint32_t dd[8], at[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      ad[u] = smem_desc(s0 + (u % nsrc) * 4096, 128, 256);
      bd[u] = smem_desc(s0 + 65536 + (u % nsrc) * (N * 32), 128, 256);
      dd[u] = tb + (u % rot) * N;
      at[u] = tb + 448 + (u % nsrc) * 8;
    }
    for (int it = 0; it < count; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (TS) mma_i8_ts(dd[u], at[u], bd[u], idesc, 1u);
        else mma_i8(dd[u], ad[u], bd[u], idesc, 1u);
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait_guarded(smem_u32(&bar), 0, (int*)out + 8, 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tb, 512);
  }
}


}  // namespace oz
}  // namespace gprc
