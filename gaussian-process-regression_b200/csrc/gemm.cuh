// gemm.cuh -- the FP64 tensor-core (DMMA) GEMM core of libgprc and the tile policies built on it.
//
// One CTA (8 DMMA warps as 2 x 4, plus one producer warp) owns one 128 x 128 output tile; each DMMA warp owns 64 x 32
// of it as 8 x 4 m8n8k4 accumulator fragments (128 registers).  Operand k-slices (BK = 16) are staged in a 4-deep ring
// of shared-memory buffers that the producer warp fills through the TMA engine with 1-D bulk copies (cp.async.bulk ->
// SASS UBLKCP) completing on per-stage "full" mbarriers; the DMMA warps hand a stage back through an "empty" mbarrier.
// Rows are padded by 4 doubles so that every fragment load (LDS.64) is bank-conflict free.
// FP64 DMMA runs at 64 FMA/clk/SM, i.e. one DMMA.8x8x4 per 16 clk per SM sub-partition, so the kernel is bound by
// the tensor pipe: per k-slice a warp issues 12 LDS.64 for 32 DMMAs (19 % of the shared-memory bandwidth).
//
// Everything the library does at O(n^3) goes through this mainloop with a different policy:
//   SyrkPolicy     C -= P P^T on lower tiles            Cholesky left-looking panel update and trailing update
//   TrsmPolicy     C  = C Linv^T (in place)             Cholesky panel solve with the inverted diagonal block
//   Trtri1Policy   T  = L21 W11 (stored transposed)     level-wise recursive inversion of L, first product
//   Trtri2Policy   W21 = -W22 T                         ... second product
//   TrmmNormPolicy colsum((W Ks)^2) per row block       predictive variance v = L^-1 K_star, R/GPRclass.R:162-164
//   DgemmPolicy    C = beta C + alpha A op(B)           tests and the roofline microbenchmark
#pragma once
#include "common.cuh"

namespace gprc {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int GEMM_CONSUMERS = 256;                  // 8 DMMA warps
constexpr int GEMM_THREADS = GEMM_CONSUMERS + 32;    // + 1 producer warp that only drives the TMA engine
constexpr int LDA_S = BM + 4;     // MN-major tile [BK][LDA_S]: (lane%4)*132 mod 16 = {0,4,8,12} -> conflict free
constexpr int LDB_MN_S = BN + 4;  // MN-major B tile [BK][LDB_MN_S]
constexpr int LDB_K_S = BK + 4;   // K-major B tile [BN][LDB_K_S]: (lane/4)*20 mod 16 = {0,4,8,12} per half warp
constexpr int A_STAGE_DOUBLES = BK * LDA_S;
constexpr int B_STAGE_DOUBLES = (BN * LDB_K_S > BK * LDB_MN_S) ? BN * LDB_K_S : BK * LDB_MN_S;
constexpr int STAGE_DOUBLES = A_STAGE_DOUBLES + B_STAGE_DOUBLES;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_DOUBLES * 8 + 128;

struct TileWork {
  const double* A;  // element (m, k) at A[m + k * lda], m in [0, 128)
  long lda;
  const double* B;  // B_KMAJOR ? element (k, n) at B[k + n * ldb] : element (k, n) at B[n + k * ldb]
  long ldb;
  int k_begin, k_end;  // multiples of BK
};

using Acc = double[8][4][2];

struct WarpCoord {
  int lane, warp_m, warp_n;
  __device__ __forceinline__ WarpCoord() {
    lane = threadIdx.x & 31;
    int w = (threadIdx.x >> 5) & 7;
    warp_m = w & 1;
    warp_n = w >> 1;
  }
  // tile-local coordinates of accumulator element acc[mb][nb][r]
  __device__ __forceinline__ int row(int mb) const { return warp_m * 64 + mb * 8 + (lane >> 2); }
  __device__ __forceinline__ int col(int nb, int r) const { return warp_n * 32 + nb * 8 + 2 * (lane & 3) + r; }
};

// barrier among the 8 consumer warps only (the producer warp has left the kernel by then)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One k-slice, issued by ONE elected lane of the producer warp: announce the bytes (the barrier's single arrival),
// then one bulk copy per contiguous run (1 KB runs for MN-major tiles, 128 B runs for the K-major B tile).
template <bool B_KMAJOR>
__device__ __forceinline__ void produce_stage(const TileWork& w, int k0, double* stage, uint32_t bar) {
  constexpr uint32_t kBytes = (BM * BK + BN * BK) * 8;
  mbar_arrive_expect_tx(bar, kBytes);
  const uint32_t sa = smem_u32(stage), sb = smem_u32(stage + A_STAGE_DOUBLES);
  const double* ap = w.A + (long)k0 * w.lda;
#pragma unroll 4
  for (int r = 0; r < BK; ++r) bulk_g2s(sa + r * (LDA_S * 8), ap + (long)r * w.lda, BM * 8, bar);
  if (!B_KMAJOR) {
    const double* bp = w.B + (long)k0 * w.ldb;
#pragma unroll 4
    for (int r = 0; r < BK; ++r) bulk_g2s(sb + r * (LDB_MN_S * 8), bp + (long)r * w.ldb, BN * 8, bar);
  } else {
    const double* bp = w.B + k0;
#pragma unroll 8
    for (int r = 0; r < BN; ++r) bulk_g2s(sb + r * (LDB_K_S * 8), bp + (long)r * w.ldb, BK * 8, bar);
  }
}

template <bool B_KMAJOR>
__device__ __forceinline__ void compute_stage(const double* __restrict__ stage, const WarpCoord& wc, Acc& acc) {
  const double* As = stage;
  const double* Bs = stage + A_STAGE_DOUBLES;
  const int lk = wc.lane & 3, lr = wc.lane >> 2;
#pragma unroll
  for (int ks = 0; ks < BK / 4; ++ks) {
    double a[8], b[4];
    const double* ap = As + (ks * 4 + lk) * LDA_S + wc.warp_m * 64 + lr;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) a[mb] = ap[mb * 8];
    if (B_KMAJOR) {
      const double* bp = Bs + (wc.warp_n * 32 + lr) * LDB_K_S + ks * 4 + lk;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) b[nb] = bp[nb * 8 * LDB_K_S];
    } else {
      const double* bp = Bs + (ks * 4 + lk) * LDB_MN_S + wc.warp_n * 32 + lr;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) b[nb] = bp[nb * 8];
    }
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
  }
}

// Warp-specialised mainloop.  full[s]: TMA bytes landed (1 arrival + tx count); empty[s]: all 8 consumer warps have
// finished reading stage s.  No block-wide barrier inside the loop: the DMMA warps drift freely, so one warp's wait is
// covered by the other warp of its SM sub-partition.
template <class Policy>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const Policy p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_DOUBLES * 8);
  uint64_t* empty = full + STAGES;
  TileWork w;
  typename Policy::Tile t;
  if (!p.setup(w, t)) return;  // depends on blockIdx only: uniform for the CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KT = (w.k_end - w.k_begin) / BK;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), GEMM_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == GEMM_CONSUMERS / 32) {
    // ===== producer warp =====
    if (lane == 0) {
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % STAGES;
        if (kt >= STAGES) mbar_wait(smem_u32(empty + s), ((kt / STAGES) - 1) & 1);
        produce_stage<Policy::B_KMAJOR>(w, w.k_begin + kt * BK, smem + s * STAGE_DOUBLES, smem_u32(full + s));
      }
    }
    return;
  }
  // ===== consumer warps =====
  Acc acc;
#pragma unroll
  for (int mb = 0; mb < 8; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  p.prefetch(t);
  const WarpCoord wc;
  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % STAGES;
    mbar_wait(smem_u32(full + s), (kt / STAGES) & 1);
    compute_stage<Policy::B_KMAJOR>(smem + s * STAGE_DOUBLES, wc, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(empty + s));
  }
  p.epilogue(t, acc, smem);
}

// decode a linear index into the lower triangle (i >= j) enumerated row by row: t = i (i + 1) / 2 + j
__device__ __forceinline__ void tri_decode(long t, int& i, int& j) {
  long ii = (long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (ii * (ii + 1) / 2 > t) --ii;
  while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
  i = (int)ii;
  j = (int)(t - ii * (ii + 1) / 2);
}

// ---------------------------------------------------------------------------------------------------------------
// SYRK / panel update on lower tiles:  C(ti, tj) -= sum_{k in [k0, k1)} P(ti, k) P(tj, k)^T
//   mode 0 (column): tj = col_tile, ti = col_tile + blockIdx.x
//   mode 1 (trail):  (ti, tj) enumerate the lower triangle of tiles [first_tile, ntiles)
//   mode 2 (strips): the block columns tj in [first_tile, first_tile + ncols), each with its row tiles ti in [tj, nt)
//                    (the lookahead update of a whole outer panel in one launch)
// ---------------------------------------------------------------------------------------------------------------
struct SyrkPolicy {
  static constexpr bool B_KMAJOR = false;
  double* A;
  long ld;
  int mode, first_tile, k0, k1;
  int ncols = 1, nt = 0;  // mode 2 only
  struct Tile {
    double* C;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    int ti, tj;
    if (mode == 0) {
      tj = first_tile;
      ti = first_tile + blockIdx.x;
    } else if (mode == 2) {
      int idx = blockIdx.x;
      tj = first_tile;
      for (int q = 0; q < ncols; ++q, ++tj) {
        const int cnt = nt - tj;
        if (idx < cnt) break;
        idx -= cnt;
      }
      if (tj >= first_tile + ncols) return false;
      ti = tj + idx;
    } else {
      tri_decode(blockIdx.x, ti, tj);
      ti += first_tile;
      tj += first_tile;
    }
    w.A = A + (long)ti * NB;
    w.lda = ld;
    w.B = A + (long)tj * NB;
    w.ldb = ld;
    w.k_begin = k0;
    w.k_end = k1;
    t.C = A + (long)ti * NB + (long)tj * NB * ld;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile& t) const {
    // pull the C tile (128 columns x 8 lines of 128 B) into L2 while the mainloop runs: 4 lines per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int line = threadIdx.x + q * GEMM_CONSUMERS;
      prefetch_l2(t.C + (long)(line >> 3) * ld + (line & 7) * 16);
    }
  }
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      double* cp0 = t.C + (long)wc.col(nb, 0) * ld;
      double* cp1 = cp0 + ld;
      double c0[8], c1[8];
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        c0[mb] = cp0[wc.row(mb)];
        c1[mb] = cp1[wc.row(mb)];
      }
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        cp0[wc.row(mb)] = c0[mb] - acc[mb][nb][0];
        cp1[wc.row(mb)] = c1[mb] - acc[mb][nb][1];
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Panel solve with the inverted diagonal block:  A(ti, j) = A(ti, j) * Linv_j^T, ti = j + 1 + blockIdx.x, in place
// (all k-slices of the tile are in shared memory / registers before the epilogue writes it back)
// ---------------------------------------------------------------------------------------------------------------
struct TrsmPolicy {
  static constexpr bool B_KMAJOR = false;
  double* A;
  long ld;
  const double* linv;  // 128 x 128 col-major: element (k, n) of the B operand is Linv[n, k] = linv[n + k * 128]
  int j;
  struct Tile {
    double* C;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    int ti = j + 1 + blockIdx.x;
    t.C = A + (long)ti * NB + (long)j * NB * ld;
    w.A = t.C;
    w.lda = ld;
    w.B = linv;
    w.ldb = NB;
    w.k_begin = 0;
    w.k_end = NB;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double* cp = t.C + (long)wc.col(nb, r) * ld;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) cp[wc.row(mb)] = acc[mb][nb][r];
      }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Level-wise recursive inversion W = L^-1.  At a level with left-part size s (in tiles), group g covers tiles
// [2 s g, 2 s g + s) (left, W11 known) and [2 s g + s, min(2 s g + 2 s, nt)) (right, W22 known):
//   phase 1:  T   = L21 * W11      (k >= column tile: W11 is lower triangular)   stored TRANSPOSED in S
//   phase 2:  W21 = - W22 * T      (k <= row tile: W22 is lower triangular)
// W11 enters phase 1 as the "B" operand, element (k, c) = W11[k, c]; read from W itself that is a K-major operand
// (128-byte runs, which the TMA engine serves at only ~1 request / 54 clk).  So the inversion also maintains
// Wt = W^T (upper triangular): W11[k, c] = Wt[c + k ld] is MN-major and streams in 1 KB runs; phase 2 writes every
// new tile to both W and Wt.
// grid = (s, s, groups): blockIdx.x = column tile tj, blockIdx.y = row tile ti (within the right part)
// ---------------------------------------------------------------------------------------------------------------
struct Trtri1Policy {
  static constexpr bool B_KMAJOR = false;
  const double* L;
  const double* Wt;
  double* S;
  long ld;
  int s, nt;
  struct Tile {
    double* C;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    const int g = blockIdx.z, tj = blockIdx.x, ti = blockIdx.y;
    const int gl = 2 * s * g, gr = gl + s;
    const int r = min(s, nt - gr);
    if (ti >= r) return false;
    w.A = L + (long)(gr + ti) * NB + (long)gl * NB * ld;
    w.lda = ld;
    w.B = Wt + (long)(gl + tj) * NB + (long)gl * NB * ld;  // element (k, n) = W11[k, n] = Wt[n + k ld]
    w.ldb = ld;
    w.k_begin = tj * NB;
    w.k_end = s * NB;
    t.C = S + (long)(gl + tj) * NB + (long)(gr + ti) * NB * ld;  // transposed position
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
      double* cp = t.C + (long)wc.row(mb) * ld;  // row of T -> column of S
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        *reinterpret_cast<double2*>(cp + wc.col(nb, 0)) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
    }
  }
};

struct Trtri2Policy {
  static constexpr bool B_KMAJOR = false;
  double* W;
  double* Wt;
  const double* S;
  long ld;
  int s, nt;
  struct Tile {
    double* C;
    double* Ct;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    const int g = blockIdx.z, tj = blockIdx.x, ti = blockIdx.y;
    const int gl = 2 * s * g, gr = gl + s;
    const int r = min(s, nt - gr);
    if (ti >= r) return false;
    w.A = W + (long)(gr + ti) * NB + (long)gr * NB * ld;
    w.lda = ld;
    w.B = S + (long)(gl + tj) * NB + (long)gr * NB * ld;  // element (k, n) = T[k, n] = S[n + k * ld]
    w.ldb = ld;
    w.k_begin = 0;
    w.k_end = (ti + 1) * NB;
    t.C = W + (long)(gr + ti) * NB + (long)(gl + tj) * NB * ld;
    t.Ct = Wt + (long)(gl + tj) * NB + (long)(gr + ti) * NB * ld;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double* cp = t.C + (long)wc.col(nb, r) * ld;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) cp[wc.row(mb)] = -acc[mb][nb][r];
      }
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
      double* cp = t.Ct + (long)wc.row(mb) * ld;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        *reinterpret_cast<double2*>(cp + wc.col(nb, 0)) = make_double2(-acc[mb][nb][0], -acc[mb][nb][1]);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Predictive variance:  V = W * Ks  (W = L^-1 lower triangular n x n; Ks given TRANSPOSED, KsT mc x n with the test
// point contiguous, so that both operands stream in 1 KB runs), never stored: the epilogue reduces the squares of each
// tile column and writes partial[ti][column]  (R/GPRclass.R:162-164: v <- solve(L, K_star); colSums(v * v)).
// Tiles are issued heaviest first (largest row tile = longest k range) in groups of GROUP row tiles so that one wave
// of CTAs shares both W row panels and Ks column panels in L2.
// If VoutT != nullptr the tile of V is stored as well, transposed (full predictive covariance, R/GPRclass.R:167).
// ---------------------------------------------------------------------------------------------------------------
struct TrmmNormPolicy {
  static constexpr bool B_KMAJOR = false;
  static constexpr int GROUP = 8;
  const double* W;
  long ldw;
  const double* KsT;  // element (k, t) at KsT[t + k * ldk]
  long ldk;
  double* partial;  // [nt][ldp]
  long ldp;
  double* VoutT;
  long ldv;
  int nt, ntc;  // row tiles of W, column tiles of Ks
  struct Tile {
    int ti, tc;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    const long b = blockIdx.x;
    const long per_group = (long)GROUP * ntc;
    const int g = (int)(b / per_group);
    const long r = b % per_group;
    const int hi = nt - 1 - g * GROUP;  // heaviest row tile of this group
    const int gsize = min(GROUP, hi + 1);
    const int tc = (int)(r / GROUP), ii = (int)(r % GROUP);
    if (ii >= gsize) return false;
    t.ti = hi - ii;
    t.tc = tc;
    w.A = W + (long)t.ti * NB;
    w.lda = ldw;
    w.B = KsT + (long)tc * NB;
    w.ldb = ldk;
    w.k_begin = 0;
    w.k_end = (t.ti + 1) * NB;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double* smem) const {
    const WarpCoord wc;
    if (VoutT) {  // V^T[t + k * ldv]: test point t contiguous, so that Sigma = K** - V^T V is an "A B^T" product
      double* vp = VoutT + (long)t.tc * NB + (long)t.ti * NB * ldv;
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        double* cp = vp + (long)wc.row(mb) * ldv;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          *reinterpret_cast<double2*>(cp + wc.col(nb, 0)) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
      }
    }
    consumer_sync();  // every consumer warp is past its last stage: shared memory is reused for the reduction
    double* red = smem;  // [2][128]
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double s = 0.0;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) s = fma(acc[mb][nb][r], acc[mb][nb][r], s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((wc.lane >> 2) == 0) red[wc.warp_m * 128 + wc.col(nb, r)] = s;
      }
    consumer_sync();
    if (threadIdx.x < 128)
      partial[(long)t.ti * ldp + (long)t.tc * NB + threadIdx.x] = red[threadIdx.x] + red[128 + threadIdx.x];
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Predictive variance without forming L^-1: blocked forward substitution on the transposed right-hand side
// T = K_star^T (mc x n, test point contiguous).  For block row i = 0, 1, ...:
//   TrsmLeftUpdatePolicy   T[:, i] -= V[:, <i] L[i, <i]^T      i.e.  R_i = Ks_i - L[i, <i] V[<i]      (K = 128 i)
//   TrsmLeftDiagPolicy     V_i = Linv_i R_i  in place, and partial[i][t] = sum over the block's rows of V^2
// One CTA per tile of 128 test points: with chunks of 148 * 128 test points every launch is exactly one wave.
// Costs n^2 m flops like the W-based pass but skips the n^3/3 inversion (the Amdahl term when a factor is shared by
// many GPUs that each predict a shard).
// ---------------------------------------------------------------------------------------------------------------
struct TrsmLeftUpdatePolicy {
  static constexpr bool B_KMAJOR = false;
  const double* L;
  long ldl;
  double* T;
  long ldt;
  int i;
  int k0 = 0;  // first block column of the product: R_i -= L[i, k0 .. i) V[k0 .. i)  (k0 > 0: the in-pair step of the
               // INT8 pass, whose tensor-core update stopped at the pair boundary)
  struct Tile {
    double* C;  // element (m, n) of the tile at C[n + m * ldt]
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    const int tc = blockIdx.x;
    w.A = L + (long)i * NB;
    w.lda = ldl;
    w.B = T + (long)tc * NB;  // element (k, n) = V[k, t] = T[t + k * ldt]
    w.ldb = ldt;
    w.k_begin = k0 * NB;
    w.k_end = i * NB;
    t.C = T + (long)tc * NB + (long)i * NB * ldt;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile& t) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int line = threadIdx.x + q * GEMM_CONSUMERS;
      prefetch_l2(t.C + (long)(line >> 3) * ldt + (line & 7) * 16);
    }
  }
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
      double* cp = t.C + (long)wc.row(mb) * ldt;
      double2 v[4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) v[nb] = *reinterpret_cast<const double2*>(cp + wc.col(nb, 0));
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        v[nb].x -= acc[mb][nb][0];
        v[nb].y -= acc[mb][nb][1];
        *reinterpret_cast<double2*>(cp + wc.col(nb, 0)) = v[nb];
      }
    }
  }
};

struct TrsmLeftDiagPolicy {
  static constexpr bool B_KMAJOR = false;
  const double* linv;  // inverted diagonal block i: element (m, k) at linv[m + k * 128]
  double* T;
  long ldt;
  int i;
  double* partial;  // [n_tiles][ldp]
  long ldp;
  struct Tile {
    double* C;
    int tc;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    t.tc = blockIdx.x;
    t.C = T + (long)t.tc * NB + (long)i * NB * ldt;
    w.A = linv;
    w.lda = NB;
    w.B = t.C;  // element (k, n) = R_i[k, t]
    w.ldb = ldt;
    w.k_begin = 0;
    w.k_end = NB;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double* smem) const {
    const WarpCoord wc;
    // all 8 k-slices of the tile have been consumed: overwrite R_i by V_i (in place, transposed storage)
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
      double* cp = t.C + (long)wc.row(mb) * ldt;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        *reinterpret_cast<double2*>(cp + wc.col(nb, 0)) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
    }
    consumer_sync();
    double* red = smem;  // [2][128]
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double s = 0.0;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) s = fma(acc[mb][nb][r], acc[mb][nb][r], s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((wc.lane >> 2) == 0) red[wc.warp_m * 128 + wc.col(nb, r)] = s;
      }
    consumer_sync();
    if (threadIdx.x < 128)
      partial[(long)i * ldp + (long)t.tc * NB + threadIdx.x] = red[threadIdx.x] + red[128 + threadIdx.x];
  }
};

// ---------------------------------------------------------------------------------------------------------------
// The same substitution as ONE persistent kernel.  Work item = (block row i, tile t): update then diagonal solve,
// handed out row by row by an atomic counter.  A block row of a tile only waits for the previous block row of the
// SAME tile (tile columns never depend on each other): a per-tile progress counter in global
// memory (release / acquire) replaces the grid-wide barrier that a kernel boundary is, so the items of neighbouring
// block rows overlap and no SM idles when the number of tiles is not a multiple of 148.  Items are only ever assigned
// to running CTAs and every prerequisite has a smaller index, so the scheme needs no co-residency guarantee.
// Data written by other CTAs with ordinary stores is read here through the TMA engine: the acquiring CTA issues
// fence.proxy.async before its bulk copies.
// ---------------------------------------------------------------------------------------------------------------
struct TrsmPersistParams {
  const double* L;
  long ldl;
  const double* dinv;
  double* T;
  long ldt;
  double* partial;
  long ldp;
  int nt, ntc;
  unsigned long long* next;  // work counter (zeroed by the host)
  int* progress;             // [ntc] items completed per tile (zeroed by the host)
  int* error;                // watchdog: set if a dependency never arrives
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(GEMM_THREADS, 1) trsm_persistent_kernel(const TrsmPersistParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_DOUBLES * 8);
  uint64_t* empty = full + STAGES;
  __shared__ long long s_item;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), GEMM_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  // item = (block row i, tile t), row-major; it runs the update and then the diagonal solve of that block row for
  // that tile.  Its only prerequisite, item (i - 1, t), was handed out ntc items earlier, so with ntc >= ~2 x 148 a
  // fetched item is practically always ready and nobody spins.
  const long long total = (long long)p.ntc * p.nt;
  long long base = 0;  // k-slices this CTA has pushed through the ring so far (identical for producer and consumers)
  const WarpCoord wc;
  for (;;) {
    if (threadIdx.x == 0) {
      long long item = (long long)atomicAdd(p.next, 1ULL);
      if (item < total) {
        const int t = (int)(item % p.ntc), i = (int)(item / p.ntc);
        long long spins = 0;
        while (ld_acquire_gpu(p.progress + t) < i) {
          __nanosleep(200);
          if (++spins > (1LL << 26)) {  // ~10 s: something is wrong; fail loudly instead of hanging the device
            atomicExch(p.error, 1);
            item = total;
            break;
          }
        }
        if (item < total && *reinterpret_cast<volatile int*>(p.error)) item = total;
      }
      s_item = item;
    }
    __syncthreads();
    const long long item = s_item;
    if (item >= total) break;
    const int t = (int)(item % p.ntc), i = (int)(item / p.ntc);
    double* Ctile = p.T + (long)t * NB + (long)i * NB * p.ldt;  // element (m, n) of the tile at C[n + m * ldt]
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      // rows of T written with ordinary stores (by other CTAs before the acquire above, by this CTA in phase 0) are
      // about to be read by the TMA engine
      fence_proxy_async();
      TileWork w;
      if (phase == 0) {  // R_i = Ks_i - L[i, <i] V[<i]
        w.A = p.L + (long)i * NB;
        w.lda = p.ldl;
        w.B = p.T + (long)t * NB;
        w.ldb = p.ldt;
        w.k_begin = 0;
        w.k_end = i * NB;
      } else {  // V_i = Linv_i R_i
        w.A = p.dinv + (long)i * NB * NB;
        w.lda = NB;
        w.B = Ctile;
        w.ldb = p.ldt;
        w.k_begin = 0;
        w.k_end = NB;
      }
      const int KT = (w.k_end - w.k_begin) / BK;
      if (warp == GEMM_CONSUMERS / 32) {
        if (lane == 0) {
          for (int kt = 0; kt < KT; ++kt) {
            const long long gk = base + kt;
            const int s = (int)(gk % STAGES);
            if (gk >= STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(((gk / STAGES) - 1) & 1));
            produce_stage<false>(w, w.k_begin + kt * BK, smem + s * STAGE_DOUBLES, smem_u32(full + s));
          }
        }
      } else if (KT > 0) {
        Acc acc;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        if (phase == 0) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            int line = threadIdx.x + qq * GEMM_CONSUMERS;
            prefetch_l2(Ctile + (long)(line >> 3) * p.ldt + (line & 7) * 16);
          }
        }
        for (int kt = 0; kt < KT; ++kt) {
          const long long gk = base + kt;
          const int s = (int)(gk % STAGES);
          mbar_wait(smem_u32(full + s), (uint32_t)((gk / STAGES) & 1));
          compute_stage<false>(smem + s * STAGE_DOUBLES, wc, acc);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(empty + s));
        }
        if (phase == 0) {
          TrsmLeftUpdatePolicy pol{p.L, p.ldl, p.T, p.ldt, i};
          TrsmLeftUpdatePolicy::Tile tile{Ctile};
          pol.epilogue(tile, acc, smem);
        } else {
          TrsmLeftDiagPolicy pol{p.dinv + (long)i * NB * NB, p.T, p.ldt, i, p.partial, p.ldp};
          TrsmLeftDiagPolicy::Tile tile{Ctile, t};
          pol.epilogue(tile, acc, smem);
        }
        __threadfence();
      }
      __syncthreads();
      base += KT;
    }
    if (threadIdx.x == 0) st_release_gpu(p.progress + t, i + 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Plain DGEMM (tests, roofline microbenchmark):  C = beta C + alpha A op(B)
// ---------------------------------------------------------------------------------------------------------------
template <bool KMAJOR>
struct DgemmPolicy {
  static constexpr bool B_KMAJOR = KMAJOR;
  const double* A;
  long lda;
  const double* B;
  long ldb;
  double* C;
  long ldc;
  double alpha, beta;
  int K, tiles_m;
  struct Tile {
    double* C;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    // column-of-tiles major: consecutive CTAs share the B panel
    const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
    w.A = A + (long)tm * NB;
    w.lda = lda;
    w.B = KMAJOR ? B + (long)tn * NB * ldb : B + (long)tn * NB;
    w.ldb = ldb;
    w.k_begin = 0;
    w.k_end = K;
    t.C = C + (long)tm * NB + (long)tn * NB * ldc;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile&) const {}
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double* cp = t.C + (long)wc.col(nb, r) * ldc;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb) {
          double v = alpha * acc[mb][nb][r];
          if (beta != 0.0) v += beta * cp[wc.row(mb)];
          cp[wc.row(mb)] = v;
        }
      }
  }
};

template <class Policy>
inline int launch_gemm(gprc_ctx* ctx, const Policy& p, dim3 grid, cudaStream_t stream = nullptr) {
  if (!stream) stream = ctx->stream;
  static bool configured[64] = {false};
  if (!configured[ctx->device & 63]) {
    GPRC_CUDA(cudaFuncSetAttribute(gemm_kernel<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
  gemm_kernel<Policy><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
  ctx->launches++;
  GPRC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gprc
