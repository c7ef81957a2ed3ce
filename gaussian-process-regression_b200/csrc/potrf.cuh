// potrf.cuh -- blocked Cholesky factorisation of K + sigma^2 I (north_star subsystem 2) and the level-wise
// inversion of its factor.  Replaces t(chol(K + new_noise * diag(n))) of R/GPRclass.R:142, R/fit.R:121 and
// R/GPCclass.R:80,102.
//
// Structure (nb = 128, outer panel = 4 blocks = 512 columns):
//   for each outer panel J:
//     for each block column j in J:
//        left update   A[j:, j] -= A[j:, J0:j] A[j, J0:j]^T        DMMA GEMM, K <= 384           (SyrkPolicy mode 0)
//        potrf_diag    factor the 128 x 128 diagonal block in shared memory, invert it in place
//                      (warp-shuffle-free column sweep; reports the first non-positive pivot in `info`)
//        panel solve   A[j+1:, j] = A[j+1:, j] Linv_j^T              DMMA GEMM, K = 128            (TrsmPolicy)
//     trailing update  A[Jend:, Jend:] -= A[Jend:, J] A[Jend:, J]^T  DMMA SYRK on lower tiles, K = 512 (mode 1)
// Every trailing element is read and written once per 512 columns (64 flop/byte), so the factorisation is bound by
// the FP64 tensor pipe; the only serial piece is potrf_diag (one CTA).
#pragma once
#include "gemm.cuh"

namespace gprc {

constexpr int PD_LDS = NB + 1;
constexpr int PD_SMEM_BYTES = NB * PD_LDS * 8;
constexpr int OUTER_BLOCKS = 4;

// Factor the diagonal block j of A (lower, in place; zeros written above the diagonal of the block) and write
// the inverse of the factor to linv (128 x 128 col-major, zeros above the diagonal).
// info: atomicMin of the 1-based global index of the first pivot that is not > 0 (LAPACK dpotrf convention).
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* A, long ld, int j, double* linv, long* info,
                                                         double* diag_out /* nullable: L_ii for this block */) {
  extern __shared__ __align__(16) unsigned char pd_raw[];
  double* S = reinterpret_cast<double*>(pd_raw);  // S[c * PD_LDS + r]
  double* Ajj = A + (long)j * NB * (ld + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    S[c * PD_LDS + r] = (r >= c) ? Ajj[r + (long)c * ld] : 0.0;
  }
  const int i = tid & (NB - 1), half = tid >> 7;
  for (int c = 0; c < NB; ++c) {
    __syncthreads();
    const double d = S[c * PD_LDS + c];
    if (!(d > 0.0)) {
      if (tid == 0) atomicMin(reinterpret_cast<unsigned long long*>(info), (unsigned long long)((long)j * NB + c + 1));
    }
    const double r = sqrt(d);
    const double rinv = 1.0 / r;
    double li = 0.0;
    if (i > c) {
      li = S[c * PD_LDS + i] * rinv;
      for (int cc = c + 1 + half; cc <= i; cc += 2) {
        const double lc = S[c * PD_LDS + cc] * rinv;
        S[cc * PD_LDS + i] = fma(-li, lc, S[cc * PD_LDS + i]);
      }
    }
    __syncthreads();
    if (half == 0) {
      if (i > c) S[c * PD_LDS + i] = li;
      if (i == c) S[c * PD_LDS + c] = r;
    }
  }
  __syncthreads();
  // write L back (explicit zeros above the diagonal of the block)
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    Ajj[r + (long)c * ld] = S[c * PD_LDS + r];
  }
  if (diag_out && tid < NB) diag_out[(long)j * NB + tid] = S[tid * PD_LDS + tid];
  __syncthreads();
  // in-place inversion of the lower-triangular factor (LAPACK dtrti2, lower): columns from the last to the first;
  //   W[c][c] = 1 / L[c][c];  W[c+1:, c] = -W[c][c] * (W[c+1:, c+1:] * L[c+1:, c])
  for (int c = NB - 1; c >= 0; --c) {
    double y = 0.0;
    const double wcc = 1.0 / S[c * PD_LDS + c];
    if (half == 0 && i > c) {
      for (int k = c + 1; k <= i; ++k) y = fma(S[k * PD_LDS + i], S[c * PD_LDS + k], y);  // W[i][k] * L[k][c]
    }
    __syncthreads();
    if (half == 0) {
      if (i > c) S[c * PD_LDS + i] = -wcc * y;
      if (i == c) S[c * PD_LDS + c] = wcc;
    }
    __syncthreads();
  }
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e & (NB - 1), c = e >> 7;
    linv[e] = S[c * PD_LDS + r];  // linv[r + c * 128]; strictly upper entries are still the zeros loaded above
  }
}

// In-place Cholesky of the lower triangle of A (n, ld multiples of 128).  dinv receives the inverted diagonal blocks,
// ddiag (nullable) the diagonal of L.  d_info must be initialised to LONG_MAX-like sentinel by the caller.
inline int potrf_blocked(gprc_ctx* ctx, double* A, long n, long ld, double* dinv, long* d_info, double* ddiag) {
  static bool configured[64] = {false};
  if (!configured[ctx->device & 63]) {
    GPRC_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM_BYTES));
    configured[ctx->device & 63] = true;
  }
  const int nt = (int)(n / NB);
  for (int J0 = 0; J0 < nt; J0 += OUTER_BLOCKS) {
    const int Jend = (J0 + OUTER_BLOCKS < nt) ? J0 + OUTER_BLOCKS : nt;
    for (int j = J0; j < Jend; ++j) {
      if (j > J0) {
        SyrkPolicy p{A, ld, 0, j, J0 * NB, j * NB};
        GPRC_CHECK(launch_gemm(ctx, p, dim3(nt - j)));
      }
      potrf_diag_kernel<<<1, 256, PD_SMEM_BYTES, ctx->stream>>>(A, ld, j, dinv + (long)j * NB * NB, d_info, ddiag);
      ctx->launches++;
      GPRC_CUDA(cudaGetLastError());
      if (j + 1 < nt) {
        TrsmPolicy p{A, ld, dinv + (long)j * NB * NB, j};
        GPRC_CHECK(launch_gemm(ctx, p, dim3(nt - j - 1)));
      }
    }
    if (Jend < nt) {
      const long t = nt - Jend;
      SyrkPolicy p{A, ld, 1, Jend, J0 * NB, Jend * NB};
      GPRC_CHECK(launch_gemm(ctx, p, dim3((unsigned)(t * (t + 1) / 2))));
    }
  }
  return 0;
}

// copy the inverted diagonal blocks into the diagonal tiles of W (full tiles incl. the explicit zeros)
__global__ void place_diag_blocks_kernel(const double* __restrict__ dinv, double* __restrict__ W, long ld) {
  const int j = blockIdx.x;
  const double* src = dinv + (long)j * NB * NB;
  double* dst = W + (long)j * NB * (ld + 1);
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) dst[(e & (NB - 1)) + (long)(e >> 7) * ld] = src[e];
}

// W = L^-1 (lower), level by level (see Trtri1Policy).  S: n x n scratch (only block positions strictly above the
// diagonal are written, so S may alias the buffer holding L).
inline int trtri_levels(gprc_ctx* ctx, const double* L, long n, long ld, const double* dinv, double* W, double* S) {
  const int nt = (int)(n / NB);
  place_diag_blocks_kernel<<<nt, 256, 0, ctx->stream>>>(dinv, W, ld);
  ctx->launches++;
  GPRC_CUDA(cudaGetLastError());
  for (int s = 1; s < nt; s *= 2) {
    const int groups = (nt + 2 * s - 1) / (2 * s);
    Trtri1Policy p1{L, W, S, ld, s, nt};
    GPRC_CHECK(launch_gemm(ctx, p1, dim3(s, s, groups)));
    Trtri2Policy p2{W, S, ld, s, nt};
    GPRC_CHECK(launch_gemm(ctx, p2, dim3(s, s, groups)));
  }
  return 0;
}

}  // namespace gprc
