// potrf.cuh -- blocked Cholesky factorisation of K + sigma^2 I (north_star subsystem 2) and the level-wise
// inversion of its factor.  Replaces t(chol(K + new_noise * diag(n))) of R/GPRclass.R:142, R/fit.R:121 and
// R/GPCclass.R:80,102.
//
// Structure (nb = 128, outer panel = 4 blocks = 512 columns; see potrf_blocked for the two-stream lookahead):
//   for each outer panel J:
//     for each block column j in J:
//        left update   A[j:, j] -= A[j:, J0:j] A[j, J0:j]^T        DMMA GEMM, K <= 384           (SyrkPolicy mode 0)
//        potrf_diag    factor the 128 x 128 diagonal block in shared memory and invert it in place (8-column inner
//                      blocks: DMMA updates from shared memory, warp-shuffle 8 x 8 pivot blocks); reports the first
//                      non-positive pivot in `info`
//        panel solve   A[j+1:, j] = A[j+1:, j] Linv_j^T              DMMA GEMM, K = 128            (TrsmPolicy)
//     trailing update  A[Jend:, Jend:] -= A[Jend:, J] A[Jend:, J]^T  DMMA SYRK on lower tiles, K = 512 (mode 1)
// Every trailing element is read and written once per 512 columns (64 flop/byte), so the factorisation is bound by
// the FP64 tensor pipe; the only serial piece is potrf_diag (one CTA).
#pragma once
#include "gemm.cuh"

namespace gprc {

constexpr int PD_LDS = NB + 4;  // 132: DMMA fragment reads are bank-conflict free along rows and along columns
constexpr int PD_SMEM_BYTES = (NB * PD_LDS + NB) * 8;
constexpr int OUTER_BLOCKS = 4;   // outer panel = 512 columns (1024 for large n, see potrf_blocked)
constexpr int PB = 8;            // inner block of the diagonal-block factorisation = one DMMA tile
constexpr int PNB = NB / PB;     // 16

// Factor the diagonal block j of A (lower, in place; zeros written above the diagonal of the block) and write
// the inverse of the factor to linv (128 x 128 col-major, zeros above the diagonal).
// info: atomicMin of the 1-based global index of the first pivot that is not > 0 (LAPACK dpotrf convention).
//
// The block lives in shared memory (S[c * 132 + r]).  Factorisation is left-looking over 16 inner blocks of 8 columns:
//   (a) the 8-column panel is updated with everything to its left by m8n8k4 DMMAs straight from shared memory,
//   (b) warp 0 factors the 8 x 8 pivot block in registers, exchanging pivots and multipliers with warp shuffles,
//   (c) one thread per row solves its 8 entries of the panel against the pivot block.
// The inverse is then formed in place: 8 x 8 diagonal blocks by substitution, and W21 = -W22 (L21 W11) level by level
// (block sizes 8, 16, 32, 64) again with DMMAs; the product T = L21 W11 is parked in the mirrored upper block.
// The body is shared by the stand-alone kernel (256 threads, __syncthreads) and the persistent tile Cholesky below, where
// the 256 DMMA threads of a 288-thread CTA run it (NAMED: barrier 1 among those 256 threads only).
template <bool NAMED>
__device__ __forceinline__ void pd_sync() {
  if (NAMED) consumer_sync();
  else __syncthreads();
}
#define __syncthreads_pd() pd_sync<NAMED>()
template <bool NAMED>
__device__ __forceinline__ void potrf_diag_body(double* S /* shared: PD_SMEM_BYTES */, double* A, long ld, int j,
                                                double* linv, long* info, double* diag_out) {
  double* rdiag = S + NB * PD_LDS;                // 1 / L_ii
  double* Ajj = A + (long)j * NB * (ld + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lk = lane & 3, lr = lane >> 2;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    S[c * PD_LDS + r] = (r >= c) ? Ajj[r + (long)c * ld] : 0.0;
  }
  __syncthreads_pd();

  for (int jb = 0; jb < PNB; ++jb) {
    const int c0 = jb * PB;
    if (jb > 0) {
      // (a) S[rt*8.., c0..c0+8) -= L[rt*8.., 0:c0) L[c0..c0+8, 0:c0)^T  for the row tiles rt >= jb
      for (int rt = jb + warp; rt < PNB; rt += 8) {
        double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;  // two accumulator pairs: halves the dependent DMMA chain
        const double* ap = S + lk * PD_LDS + rt * PB + lr;
        const double* bp = S + lk * PD_LDS + c0 + lr;
        for (int k0 = 0; k0 + 8 <= c0; k0 += 8) {
          dmma884(e0, e1, ap[k0 * PD_LDS], bp[k0 * PD_LDS]);
          dmma884(o0, o1, ap[(k0 + 4) * PD_LDS], bp[(k0 + 4) * PD_LDS]);
        }
        double* cp = S + (c0 + 2 * lk) * PD_LDS + rt * PB + lr;
        cp[0] -= e0 + o0;
        cp[PD_LDS] -= e1 + o1;
      }
      __syncthreads_pd();
    }
    if (tid == 0) {
      // (b) 8 x 8 pivot block, serially in the registers of ONE thread: the eight pivots are a dependent chain whatever
      // the mapping, and a single thread pays neither shuffles nor a second long-latency operation per pivot
      // (1 / sqrt(d) directly; L_kk = d * (1 / sqrt(d))): ~1.5k instead of ~4k clk per block, 16 blocks per tile
      double a[PB][PB];
#pragma unroll
      for (int q = 0; q < PB; ++q)
#pragma unroll
        for (int rr = q; rr < PB; ++rr) a[rr][q] = S[(c0 + q) * PD_LDS + c0 + rr];
#pragma unroll
      for (int kk = 0; kk < PB; ++kk) {
        const double d = a[kk][kk];
        if (!(d > 0.0))
          atomicMin(reinterpret_cast<unsigned long long*>(info), (unsigned long long)((long)j * NB + c0 + kk + 1));
        const double rinv = rsqrt(d);
        a[kk][kk] = d * rinv;
        rdiag[c0 + kk] = rinv;
#pragma unroll
        for (int rr = kk + 1; rr < PB; ++rr) a[rr][kk] *= rinv;
#pragma unroll
        for (int cc = kk + 1; cc < PB; ++cc)
#pragma unroll
          for (int rr = cc; rr < PB; ++rr) a[rr][cc] = fma(-a[rr][kk], a[cc][kk], a[rr][cc]);
      }
#pragma unroll
      for (int q = 0; q < PB; ++q)
#pragma unroll
        for (int rr = 0; rr < PB; ++rr) S[(c0 + q) * PD_LDS + c0 + rr] = (rr >= q) ? a[rr][q] : 0.0;
    }
    __syncthreads_pd();
    // (c) rows below the pivot block: x Ld^T = a  (forward substitution along the row's 8 entries)
    if (tid < NB && tid >= c0 + PB) {
      double x[PB];
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        double v = S[(c0 + q) * PD_LDS + tid];
#pragma unroll
        for (int pp = 0; pp < q; ++pp) v = fma(-x[pp], S[(c0 + pp) * PD_LDS + c0 + q], v);
        x[q] = v * rdiag[c0 + q];
      }
#pragma unroll
      for (int q = 0; q < PB; ++q) S[(c0 + q) * PD_LDS + tid] = x[q];
    }
    __syncthreads_pd();
  }
  // write L back (explicit zeros above the diagonal: the strictly upper part was loaded as zero and never touched)
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    Ajj[r + (long)c * ld] = S[c * PD_LDS + r];
  }
  if (diag_out && tid < NB) diag_out[(long)j * NB + tid] = S[tid * PD_LDS + tid];

  // ---- in-place inversion ----
  // (d) 8 x 8 diagonal blocks: thread = (block b, column c) solves Ld w = e_c
  double wcol[PB];
  const int ib = tid >> 3, ic = tid & 7;
  if (tid < NB) {
    const double* Ld = S + (ib * PB) * PD_LDS + ib * PB;  // Ld[i][k] = Ld[k * PD_LDS + i]
#pragma unroll
    for (int i = 0; i < PB; ++i) {
      double acc = (i == ic) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) acc = fma(-Ld[k * PD_LDS + i], wcol[k], acc);
      wcol[i] = (i >= ic) ? acc * rdiag[ib * PB + i] : 0.0;
    }
  }
  __syncthreads_pd();
  if (tid < NB) {
#pragma unroll
    for (int i = 0; i < PB; ++i) S[(ib * PB + ic) * PD_LDS + ib * PB + i] = wcol[i];
  }
  __syncthreads_pd();
  // (e) levels: left part [gl, gl + s), right part [gr, gr + s) in units of 8
  for (int s = 1; s < PNB; s *= 2) {
    const int tiles = (PNB / (2 * s)) * s * s;
    // phase 1: T = L21 W11, stored at the mirrored (strictly upper) position S[(gr*8 + x) * LDS + gl*8 + y] = T[x][y]
    for (int q = warp; q < tiles; q += 8) {
      const int g = q / (s * s), rem = q - g * s * s, ti = rem / s, tj = rem - ti * s;
      const int gl = 2 * s * g, gr = gl + s;
      double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
      const double* ap = S + ((gl + tj) * PB + lk) * PD_LDS + (gr + ti) * PB + lr;   // L21[row][k], k from tile tj
      const double* bp = S + ((gl + tj) * PB + lr) * PD_LDS + (gl + tj) * PB + lk;   // W11[k][col]
      for (int kb = tj; kb < s; ++kb) {
        const int ko = (kb - tj) * PB;
        dmma884(e0, e1, ap[ko * PD_LDS], bp[ko]);
        dmma884(o0, o1, ap[(ko + 4) * PD_LDS], bp[ko + 4]);
      }
      double* cp = S + ((gr + ti) * PB + lr) * PD_LDS + (gl + tj) * PB + 2 * lk;
      cp[0] = e0 + o0;
      cp[1] = e1 + o1;
    }
    __syncthreads_pd();
    // phase 2: W21 = -W22 T, overwriting L21
    for (int q = warp; q < tiles; q += 8) {
      const int g = q / (s * s), rem = q - g * s * s, ti = rem / s, tj = rem - ti * s;
      const int gl = 2 * s * g, gr = gl + s;
      double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
      const double* ap = S + (gr * PB + lk) * PD_LDS + (gr + ti) * PB + lr;   // W22[row][k]
      const double* bp = S + (gr * PB + lk) * PD_LDS + (gl + tj) * PB + lr;   // T[k][col]
      for (int kb = 0; kb <= ti; ++kb) {
        const int ko = kb * PB;
        dmma884(e0, e1, ap[ko * PD_LDS], bp[ko * PD_LDS]);
        dmma884(o0, o1, ap[(ko + 4) * PD_LDS], bp[(ko + 4) * PD_LDS]);
      }
      double* cp = S + ((gl + tj) * PB + 2 * lk) * PD_LDS + (gr + ti) * PB + lr;
      cp[0] = -(e0 + o0);
      cp[PD_LDS] = -(e1 + o1);
    }
    __syncthreads_pd();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    linv[e] = (r >= c) ? S[c * PD_LDS + r] : 0.0;  // the upper part holds scratch of the level products
  }
}
#undef __syncthreads_pd

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* A, long ld, int j, double* linv, long* info,
                                                         double* diag_out /* nullable: L_ii for this block */) {
  extern __shared__ __align__(16) unsigned char pd_raw[];
  potrf_diag_body<false>(reinterpret_cast<double*>(pd_raw), A, ld, j, linv, info, diag_out);
}

// ---------------------------------------------------------------------------------------------------------------
// Round 2: the whole factorisation as ONE persistent kernel for small and medium n (n <= 16 384 by default).
//
// The launch sequence below (potrf_blocked) spends three dependent launches per 128-column block -- left update,
// potrf_diag, panel solve -- and below n ~ 16k that chain, not the DMMA pipe, sets the time (n = 2048: 2.1 ms for
// 2.9 GFLOP).  Here the factor is computed tile by tile, left-looking: task (i, j), i >= j, produces tile L_ij in one go,
//     C   = A_ij - sum_{k<j} L_ik L_jk^T        DMMA GEMM with K = 128 j (SyrkPolicy's epilogue)
//     i > j:  L_ij = C Linv_j^T                  DMMA GEMM with K = 128  (TrsmPolicy's epilogue)
//     i = j:  L_jj = chol(C), Linv_j             potrf_diag_body in the shared memory of the (idle) operand ring
// Tasks are handed out column by column by an atomic counter to a resident grid; a per-ROW progress counter
// (progress[i] = number of finished tiles of row i, release / acquire) is the only synchronisation: task (i, j) needs
// rows i and j up to column j, and Linv_j.  The producer lane waits per 128-wide k-block, so the GEMM of a tile streams
// in as its operands complete, and the only thing on the critical path of column j + 1 is: diagonal block j -> panel
// solve of tile (j+1, j) -> the last k-block of the diagonal tile (j+1, j+1).  Every prerequisite of a task has a
// smaller task index and tasks are only handed to running CTAs, so no co-residency guarantee is needed.
// ---------------------------------------------------------------------------------------------------------------
struct CholPersistParams {
  double* A;
  long ld;
  int nt;
  double* dinv;
  long* info;
  double* diag;              // nullable
  unsigned long long* next;  // task counter (zeroed by the host)
  int* progress;             // [nt] finished tiles per row (zeroed by the host)
  int* error;                // watchdog
};

__device__ __forceinline__ bool chol_wait_progress(const int* prog, int want, int* error) {
  long long spins = 0;
  while (ld_acquire_gpu(prog) < want) {
    if (++spins > (1LL << 26) || *reinterpret_cast<volatile int*>(error)) {  // fail loudly instead of hanging the device
      atomicExch(error, 1);
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) chol_persistent_kernel(const CholPersistParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_DOUBLES * 8);
  uint64_t* empty = full + STAGES;
  static_assert(PD_SMEM_BYTES <= STAGES * STAGE_DOUBLES * 8, "potrf_diag_body works inside the operand ring");
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
  const long long total = (long long)p.nt * (p.nt + 1) / 2;
  long long base = 0;  // k-slices pushed through the ring so far (identical for producer and consumers)
  const WarpCoord wc;
  for (;;) {
    if (threadIdx.x == 0) {
      long long item = (long long)atomicAdd(p.next, 1ULL);
      if (*reinterpret_cast<volatile int*>(p.error)) item = total;
      s_item = item;
    }
    __syncthreads();
    const long long item = s_item;
    if (item >= total) break;
    // column-major enumeration of the lower triangle: column j holds nt - j tasks
    int j = 0;
    long long rest = item;
    while (rest >= p.nt - j) {
      rest -= p.nt - j;
      ++j;
    }
    const int i = j + (int)rest;
    double* Ctile = p.A + (long)i * NB + (long)j * NB * p.ld;
    // ---- phase 0: C -= sum_{k<j} L_ik L_jk^T, streamed in as the tiles of rows i and j complete ----
    {
      TileWork w;
      w.A = p.A + (long)i * NB;
      w.lda = p.ld;
      w.B = p.A + (long)j * NB;
      w.ldb = p.ld;
      w.k_begin = 0;
      w.k_end = j * NB;
      const int KT = j * (NB / BK);
      if (warp == GEMM_CONSUMERS / 32) {
        if (lane == 0) {
          bool ok = true;
          for (int kt = 0; kt < KT && ok; ++kt) {
            if (kt % (NB / BK) == 0) {  // a new k-block: tiles (i, kb) and (j, kb) must be final
              const int kb = kt / (NB / BK);
              ok = chol_wait_progress(p.progress + i, kb + 1, p.error) && chol_wait_progress(p.progress + j, kb + 1, p.error);
              fence_proxy_async();  // they were written with ordinary stores and are read by the TMA engine
            }
            const long long gk = base + kt;
            const int s = (int)(gk % STAGES);
            if (gk >= STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(((gk / STAGES) - 1) & 1));
            if (ok) produce_stage<false>(w, kt * BK, smem + s * STAGE_DOUBLES, smem_u32(full + s));
            else mbar_arrive(smem_u32(full + s));  // watchdog fired: release the consumers (results are garbage, error set)
          }
        }
      } else if (KT > 0) {
        Acc acc;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          int line = threadIdx.x + qq * GEMM_CONSUMERS;
          prefetch_l2(Ctile + (long)(line >> 3) * p.ld + (line & 7) * 16);
        }
        for (int kt = 0; kt < KT; ++kt) {
          const long long gk = base + kt;
          const int s = (int)(gk % STAGES);
          mbar_wait(smem_u32(full + s), (uint32_t)((gk / STAGES) & 1));
          compute_stage<false>(smem + s * STAGE_DOUBLES, wc, acc);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(empty + s));
        }
        SyrkPolicy pol{p.A, p.ld, 0, 0, 0, 0};
        SyrkPolicy::Tile tile{Ctile};
        pol.epilogue(tile, acc, smem);
        __threadfence();
        fence_proxy_async();  // phase 1 reads this tile through the TMA engine
      }
      __syncthreads();
      base += KT;
    }
    // ---- phase 1 ----
    if (i == j) {
      // diagonal tile: factor and invert in shared memory (the operand ring is idle: every stage has been consumed)
      if (warp < GEMM_CONSUMERS / 32)
        potrf_diag_body<true>(smem, p.A, p.ld, j, p.dinv + (long)j * NB * NB, p.info, p.diag);
      __threadfence();
      fence_proxy_async();  // the ring was written with ordinary stores and will be written by the TMA engine again
      __syncthreads();
    } else {
      TileWork w;
      w.A = Ctile;
      w.lda = p.ld;
      w.B = p.dinv + (long)j * NB * NB;  // element (k, n) of the B operand is Linv_j[n, k]
      w.ldb = NB;
      w.k_begin = 0;
      w.k_end = NB;
      constexpr int KT = NB / BK;
      if (warp == GEMM_CONSUMERS / 32) {
        if (lane == 0) {
          const bool ok = chol_wait_progress(p.progress + j, j + 1, p.error);  // L_jj factored, Linv_j written
          fence_proxy_async();  // Linv_j (another CTA) and C (this CTA, phase 0) were written with ordinary stores
          for (int kt = 0; kt < KT; ++kt) {
            const long long gk = base + kt;
            const int s = (int)(gk % STAGES);
            if (gk >= STAGES) mbar_wait(smem_u32(empty + s), (uint32_t)(((gk / STAGES) - 1) & 1));
            if (ok) produce_stage<false>(w, kt * BK, smem + s * STAGE_DOUBLES, smem_u32(full + s));
            else mbar_arrive(smem_u32(full + s));
          }
        }
      } else {
        Acc acc;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        for (int kt = 0; kt < KT; ++kt) {
          const long long gk = base + kt;
          const int s = (int)(gk % STAGES);
          mbar_wait(smem_u32(full + s), (uint32_t)((gk / STAGES) & 1));
          compute_stage<false>(smem + s * STAGE_DOUBLES, wc, acc);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(empty + s));
        }
        TrsmPolicy pol{p.A, p.ld, nullptr, j};
        TrsmPolicy::Tile tile{Ctile};
        pol.epilogue(tile, acc, smem);
        __threadfence();
      }
      __syncthreads();
      base += KT;
    }
    if (threadIdx.x == 0) st_release_gpu(p.progress + i, j + 1);
  }
}

// In-place Cholesky of the lower triangle of A (n, ld multiples of 128).  dinv receives the inverted diagonal blocks,
// ddiag (nullable) the diagonal of L.  d_info must be initialised to a LONG_MAX sentinel by the caller.
//
// Lookahead over two streams: the latency-bound factorisation of outer panel P + 1 (left updates, potrf_diag, panel
// solves) runs on a high-priority stream while the main stream applies outer panel P to everything right of panel
// P + 1; the narrow update of panel P + 1's own columns ("LA") goes first on the high-priority stream.
//   s1:  panel(0) | LA(0) panel(1) | LA(1) panel(2) | ...
//   s0:           | rest(0)        | rest(1)        | ...        rest(P) waits panel(P); LA(P) waits rest(P - 1)
inline int potrf_blocked(gprc_ctx* ctx, double* A, long n, long ld, double* dinv, long* d_info, double* ddiag) {
  static bool configured[64] = {false};
  if (!configured[ctx->device & 63]) {
    GPRC_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int nt = (int)(n / NB);
  if (nt >= 2 && nt <= ctx->opt_chol_tiles && ctx->d_sched) {
    // small / medium n: one persistent kernel, tile tasks with per-row progress counters (chol_persistent_kernel)
    static bool configured2[64] = {false};
    if (!configured2[ctx->device & 63]) {
      GPRC_CUDA(cudaFuncSetAttribute(chol_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
      configured2[ctx->device & 63] = true;
    }
    GPRC_CUDA(cudaMemsetAsync(ctx->d_sched, 0, 4096, ctx->stream));
    CholPersistParams q;
    q.A = A;
    q.ld = ld;
    q.nt = nt;
    q.dinv = dinv;
    q.info = d_info;
    q.diag = ddiag;
    q.next = reinterpret_cast<unsigned long long*>(ctx->d_sched);
    q.error = ctx->d_sched + 2;
    q.progress = ctx->d_sched + 4;
    const long tasks = (long)nt * (nt + 1) / 2;
    const int grid = (int)std::min<long>(tasks, ctx->sm_count);
    chol_persistent_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, ctx->stream>>>(q);
    ctx->launches++;
    GPRC_CUDA(cudaGetLastError());
    return 0;
  }
  // outer panel width: 512 columns; 1024 once the trailing update dominates (n >= 25 600), which halves the share of
  // the tile epilogues (C read-modify-write) in the trailing SYRK: 32.0 -> 33.4 TFLOP/s at n = 50k, but slower below
  const int OB = (nt >= 200) ? 2 * OUTER_BLOCKS : OUTER_BLOCKS;
  cudaStream_t s0 = ctx->stream, s1 = ctx->stream_hi;
  GPRC_CUDA(cudaEventRecord(ctx->ev_start, s0));
  GPRC_CUDA(cudaStreamWaitEvent(s1, ctx->ev_start, 0));

  auto panel = [&](int J0, int Jend) -> int {
    for (int j = J0; j < Jend; ++j) {
      if (j > J0) {
        SyrkPolicy p{A, ld, 0, j, J0 * NB, j * NB};
        GPRC_CHECK(launch_gemm(ctx, p, dim3(nt - j), s1));
      }
      potrf_diag_kernel<<<1, 256, PD_SMEM_BYTES, s1>>>(A, ld, j, dinv + (long)j * NB * NB, d_info, ddiag);
      ctx->launches++;
      GPRC_CUDA(cudaGetLastError());
      if (j + 1 < nt) {
        TrsmPolicy p{A, ld, dinv + (long)j * NB * NB, j};
        GPRC_CHECK(launch_gemm(ctx, p, dim3(nt - j - 1), s1));
      }
    }
    GPRC_CUDA(cudaEventRecord(ctx->ev_panel, s1));
    return 0;
  };

  GPRC_CHECK(panel(0, nt < OB ? nt : OB));
  for (int J0 = 0; J0 < nt; J0 += OB) {
    const int Jend = (J0 + OB < nt) ? J0 + OB : nt;
    if (Jend >= nt) break;
    const int Nend = (Jend + OB < nt) ? Jend + OB : nt;
    // LA(P): columns of the next panel, on the panel stream, after rest(P - 1) has finished with them
    if (J0 > 0) GPRC_CUDA(cudaStreamWaitEvent(s1, ctx->ev_rest, 0));
    {
      long tiles = 0;
      for (int tj = Jend; tj < Nend; ++tj) tiles += nt - tj;
      SyrkPolicy p{A, ld, 2, Jend, J0 * NB, Jend * NB, Nend - Jend, nt};
      GPRC_CHECK(launch_gemm(ctx, p, dim3((unsigned)tiles), s1));
    }
    // rest(P): everything right of the next panel, on the main stream, once panel P is factored
    GPRC_CUDA(cudaStreamWaitEvent(s0, ctx->ev_panel, 0));
    if (Nend < nt) {
      const long t = nt - Nend;
      SyrkPolicy p{A, ld, 1, Nend, J0 * NB, Jend * NB};
      GPRC_CHECK(launch_gemm(ctx, p, dim3((unsigned)(t * (t + 1) / 2)), s0));
    }
    GPRC_CUDA(cudaEventRecord(ctx->ev_rest, s0));
    GPRC_CHECK(panel(Jend, Nend));
  }
  GPRC_CUDA(cudaStreamWaitEvent(s0, ctx->ev_panel, 0));
  return 0;
}

// copy the inverted diagonal blocks into the diagonal tiles of W, and their transposes into those of Wt
__global__ void place_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ W,
                                         double* __restrict__ Wt, long ld) {
  const int j = blockIdx.x;
  const double* src = dinv + (long)j * NB * NB;
  double* dst = W + (long)j * NB * (ld + 1);
  double* dstt = Wt + (long)j * NB * (ld + 1);
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e & (NB - 1), c = e >> 7;
    dst[r + (long)c * ld] = src[e];
    dstt[r + (long)c * ld] = src[c + r * NB];
  }
}

// W = L^-1 (lower), level by level (see Trtri1Policy).  S: n x n scratch (only block positions strictly above the
// diagonal are written, so S may alias the buffer holding L).  Wt: n x n, receives W^T (needed only during the
// inversion; the caller may free it afterwards).
inline int trtri_levels(gprc_ctx* ctx, const double* L, long n, long ld, const double* dinv, double* W, double* Wt,
                        double* S) {
  const int nt = (int)(n / NB);
  place_diag_blocks_kernel<<<nt, 256, 0, ctx->stream>>>(dinv, W, Wt, ld);
  ctx->launches++;
  GPRC_CUDA(cudaGetLastError());
  for (int s = 1; s < nt; s *= 2) {
    const int groups = (nt + 2 * s - 1) / (2 * s);
    Trtri1Policy p1{L, Wt, S, ld, s, nt};
    GPRC_CHECK(launch_gemm(ctx, p1, dim3(s, s, groups)));
    Trtri2Policy p2{W, Wt, S, ld, s, nt};
    GPRC_CHECK(launch_gemm(ctx, p2, dim3(s, s, groups)));
  }
  return 0;
}

}  // namespace gprc
