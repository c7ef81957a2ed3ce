// cov.cuh -- kernel-matrix construction (north_star subsystem 1).
//
// Restates covariance_matrix(A, B, k) (R/GPRclass.R:355-357) and the six `.matrix` kernels (R/GPRclass.R:381-403)
// as fused tile kernels: the reference gathers two d x (nA nB) operand matrices and calls k once; here each CTA
// stages 64 + 64 points in shared memory, forms the 64 x 64 block of squared distances / dot products in registers,
// applies the kernel's epilogue, adds the iid noise on the diagonal in-register and writes the block once.
// Optional fusions used by predict: a row scaling (GPC's sqrt(W) * K_star, R/GPCclass.R:114) and the weighted
// column sums that give the predictive mean t(K_star) %*% alpha (R/GPRclass.R:161) without re-reading K_star.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace gprc {

struct KSpecDev {
  int id;
  double c, sigma, p, l, gamma, alpha;
  const double* sigma_vec;  // device pointer or nullptr
  int sigma_len;
};

enum { FAM_DIST = 0, FAM_DOT = 1, FAM_CONST = 2 };
__host__ __device__ inline int kernel_family(int id) {
  return (id == GPRC_SQREXP || id == GPRC_GAMMAEXP || id == GPRC_RATQUAD) ? FAM_DIST
         : (id == GPRC_CONSTANT)                                           ? FAM_CONST
                                                                           : FAM_DOT;
}

// R's `^` (arithmetic.c R_POW): exponent exactly 2 is x*x, everything else libm pow
__device__ __forceinline__ double r_pow(double x, double y) { return (y == 2.0) ? x * x : pow(x, y); }

// epilogue on the squared distance r2 = sum_d (x_d - y_d)^2.  Operation order follows the R expressions.
__device__ __forceinline__ double kfun_dist(const KSpecDev& k, double r2) {
  switch (k.id) {
    case GPRC_SQREXP:  // exp(-colSums((x - y)^2)/(2 * l^2))                      R/GPRclass.R:394
      return exp(-r2 / (2.0 * (k.l * k.l)));
    case GPRC_GAMMAEXP:  // exp(-(sqrt(colSums((x - y)^2))/l)^gamma)              R/GPRclass.R:398
      return exp(-r_pow(sqrt(r2) / k.l, k.gamma));
    default:  // (1 + colSums((x - y)^2) / (2 * alpha * l^2))^(-alpha)             R/GPRclass.R:402
      return r_pow(1.0 + r2 / (2.0 * k.alpha * (k.l * k.l)), -k.alpha);
  }
}
// epilogue on the (sigma-weighted) dot product
__device__ __forceinline__ double kfun_dot(const KSpecDev& k, double dot) {
  if (k.id == GPRC_POLYNOMIAL) return r_pow(dot + k.sigma, k.p);  // (colSums(x * y) + sigma)^p   R/GPRclass.R:390
  return dot;                                                      // colSums(sigma * x * y)        R/GPRclass.R:386
}
__device__ __forceinline__ double linear_sigma(const KSpecDev& k, int dd) {
  if (k.id != GPRC_LINEAR) return 1.0;
  if (k.sigma_len <= 0) return k.sigma;
  return k.sigma_vec[dd % k.sigma_len];  // R recycles sigma down the rows of the d x N operand
}

struct CovParams {
  KSpecDev k;
  const double* A;  // d x nA, points contiguous
  const double* B;  // d x nB
  int d;
  long nA, nB;
  double* out;  // out[i + j * ldo] = k(A[, i], B[, j])
  long ldo;
  long rows_pad, cols_pad;  // extent written (multiples of 64); beyond (nA, nB): 0, or 1 on the diagonal if pad_identity
  int lower_only;           // skip 64 x 64 tiles strictly above the diagonal
  int symmetric;            // A == B: diag_add is added where i == j
  double diag_add;
  int pad_identity;
  const double* rowscale;  // nullable: stored value is rowscale[i] * k
  const double* weights;   // nullable: pmean[tile_i][j] = sum_{i in tile} k(i, j) * weights[i]   (unscaled k)
  double* pmean;
  long ldpm;
  // the same two fusions along the other axis (K_star^T builds: rows are test points, columns training points)
  const double* colscale;    // nullable: stored value is colscale[j] * k
  const double* colweights;  // nullable: pmean[tile_j][i] = sum_{j in tile} k(i, j) * colweights[j]
  long col_offset;           // column j of this call is column j + col_offset of the full matrix (panel builds)
};

constexpr int CT = 64;    // covariance tile
constexpr int CDCH = 8;   // dimensions staged per pass

template <int FAMILY>
__global__ void __launch_bounds__(256) cov_tile_kernel(const CovParams p) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long ti = blockIdx.x, tj = blockIdx.y;
  if (p.lower_only && tj * CT + p.col_offset > ti * CT + (CT - 1)) return;
  __shared__ double sh[2 * CDCH * (CT + 1)];  // one array: the row-sum reduction below reuses all of it
  double (*As)[CT + 1] = reinterpret_cast<double (*)[CT + 1]>(sh);
  double (*Bs)[CT + 1] = reinterpret_cast<double (*)[CT + 1]>(sh + CDCH * (CT + 1));
  const long i0 = ti * CT, j0 = tj * CT;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  if (FAMILY != FAM_CONST) {
    for (int d0 = 0; d0 < p.d; d0 += CDCH) {
      const int dc = min(CDCH, p.d - d0);
      __syncthreads();
      for (int e = threadIdx.x; e < CT * dc; e += 256) {
        const int pi = e / dc, dd = e - pi * dc;
        const long gi = i0 + pi, gj = j0 + pi;
        As[dd][pi] = (gi < p.nA) ? p.A[gi * p.d + d0 + dd] : 0.0;
        Bs[dd][pi] = (gj < p.nB) ? p.B[gj * p.d + d0 + dd] : 0.0;
      }
      __syncthreads();
      for (int dd = 0; dd < dc; ++dd) {
        double a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = As[dd][tx + 16 * q];
          b[q] = Bs[dd][ty * 4 + q];
        }
        if (FAMILY == FAM_DIST) {
#pragma unroll
          for (int qa = 0; qa < 4; ++qa)
#pragma unroll
            for (int qb = 0; qb < 4; ++qb) {
              // separate rounding of the square and of the running sum (no FMA contraction), the order of R's
              // (x - y)^2 then colSums.  Not bit-for-bit R for D > 1: colSums accumulates in long double (80 bit on x86)
              // and rounds once at the end, so r^2 can differ from R's by 1 ulp; the CPU restatement under tests/ sums in double like this
              const double df = a[qa] - b[qb];
              acc[qa][qb] = __dadd_rn(acc[qa][qb], __dmul_rn(df, df));
            }
        } else {
          const double sg = linear_sigma(p.k, d0 + dd);
#pragma unroll
          for (int qa = 0; qa < 4; ++qa) {
            const double sa = (p.k.id == GPRC_LINEAR) ? sg * a[qa] : a[qa];
#pragma unroll
            for (int qb = 0; qb < 4; ++qb) acc[qa][qb] = __dadd_rn(acc[qa][qb], __dmul_rn(sa, b[qb]));
          }
        }
      }
    }
  }

  double msum[4] = {0.0, 0.0, 0.0, 0.0};  // per column b: sum over this thread's rows of k * weight
  double rsum[4] = {0.0, 0.0, 0.0, 0.0};  // per row a: sum over this thread's columns of k * colweight
#pragma unroll
  for (int qb = 0; qb < 4; ++qb) {
    const long gj = j0 + ty * 4 + qb;
#pragma unroll
    for (int qa = 0; qa < 4; ++qa) {
      const long gi = i0 + tx + 16 * qa;
      double v;
      if (gi < p.nA && gj < p.nB) {
        v = (FAMILY == FAM_DIST) ? kfun_dist(p.k, acc[qa][qb]) : (FAMILY == FAM_DOT) ? kfun_dot(p.k, acc[qa][qb]) : p.k.c;
        if (p.symmetric && gi == gj + p.col_offset) v += p.diag_add;
        if (p.weights) msum[qb] = fma(v, p.weights[gi], msum[qb]);
        if (p.colweights) rsum[qa] = fma(v, p.colweights[gj], rsum[qa]);
        if (p.rowscale) v *= p.rowscale[gi];
        if (p.colscale) v *= p.colscale[gj];
      } else {
        v = (p.pad_identity && gi == gj + p.col_offset) ? 1.0 : 0.0;
      }
      if (gi < p.rows_pad && gj < p.cols_pad) p.out[gi + gj * p.ldo] = v;
    }
  }
  if (p.weights) {
#pragma unroll
    for (int qb = 0; qb < 4; ++qb) {
      double s = msum[qb];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      const long gj = j0 + ty * 4 + qb;
      if (tx == 0 && gj < p.cols_pad) p.pmean[ti * p.ldpm + gj] = s;
    }
  }
  if (p.colweights) {
    double* red = sh;  // 16 x 64 doubles <= 2 x 8 x 65
    __syncthreads();
#pragma unroll
    for (int qa = 0; qa < 4; ++qa) red[ty * CT + tx + 16 * qa] = rsum[qa];
    __syncthreads();
    if (threadIdx.x < CT) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += red[q * CT + threadIdx.x];
      const long gi = i0 + threadIdx.x;
      if (gi < p.rows_pad) p.pmean[tj * p.ldpm + gi] = s;
    }
  }
}

// k(A, B) applied column-wise (the `.matrix` contract, R/GPRclass.R:378-380): out[i] = k(A[, i], B[, i])
__global__ void cov_pointwise_kernel(const KSpecDev k, const double* __restrict__ A, const double* __restrict__ B, int d,
                                     long n, double* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int fam = kernel_family(k.id);
  double acc = 0.0;
  if (fam == FAM_DIST) {
    for (int dd = 0; dd < d; ++dd) {
      const double df = A[i * d + dd] - B[i * d + dd];
      acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    out[i] = kfun_dist(k, acc);
  } else if (fam == FAM_DOT) {
    for (int dd = 0; dd < d; ++dd) {
      const double a = (k.id == GPRC_LINEAR) ? linear_sigma(k, dd) * A[i * d + dd] : A[i * d + dd];
      acc = __dadd_rn(acc, __dmul_rn(a, B[i * d + dd]));
    }
    out[i] = kfun_dot(k, acc);
  } else {
    out[i] = k.c;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core build (north_star subsystem 1): the d-dimensional contraction G = A^T B of a 64 x 64 tile runs on
// the FP64 tensor cores (DMMA m8n8k4 straight from the point tiles, which the TMA engine drops into shared memory
// with ONE bulk copy each: a tile of 64 points x d is contiguous in the reference's d x n layout); the epilogue
// turns G into squared distances r2 = |a|^2 + |b|^2 - 2 G (clamped at 0, exactly 0 on the diagonal of K(X, X)) or
// uses it as the dot product, applies the kernel, adds the noise on the diagonal in-register and stores the tile.
// Eligible: sqrexp / rationalquadratic / polynomial / linear with d a multiple of 4 (4..32); see cov_uses_gram for
// when it is actually chosen.  gammaexp always keeps the direct differences (its sqrt amplifies the cancellation error
// of the expansion near coincident points).  The expansion perturbs r2 by ~1e-16 (|a|^2 + |b|^2), i.e. kernel
// entries by ~1e-15 relative -- three orders below what the Cholesky of K + noise I itself commits.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GT = 64;         // tensor-core build tile (same 64 x 64 granularity as the direct build)
constexpr int GRAM_THREADS = 128;  // 4 warps as 2 x 2, each a 32 x 32 block of 4 x 4 m8n8k4 accumulators (64 registers)
constexpr int GRAM_DMAX = 32;
inline size_t gram_smem_bytes(int d) { return (size_t)(2 * GT * d + 2 * GT + 2 * GT) * 8 + 64; }

template <int FAMILY>
__global__ void __launch_bounds__(GRAM_THREADS, 4) cov_gram_kernel(const CovParams p) {
  extern __shared__ __align__(128) unsigned char graw[];
  const int D = p.d, tid = threadIdx.x;
  double* As = reinterpret_cast<double*>(graw);  // [GT][D], point-major (K-major for the DMMA fragments)
  double* Bs = As + GT * D;
  double* na = Bs + GT * D;
  double* nbv = na + GT;
  double* red = nbv + GT;  // [2][GT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * GT);
  const long ti = blockIdx.x, tj = blockIdx.y;
  if (p.lower_only && tj * GT + p.col_offset > ti * GT + (GT - 1)) return;
  const long i0 = ti * GT, j0 = tj * GT;
  const int validA = (int)max(0L, min((long)GT, p.nA - i0)), validB = (int)max(0L, min((long)GT, p.nB - j0));
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytesA = (uint32_t)validA * D * 8, bytesB = (uint32_t)validB * D * 8;
    mbar_arrive_expect_tx(smem_u32(bar), bytesA + bytesB);
    if (bytesA) bulk_g2s(smem_u32(As), p.A + i0 * D, bytesA, smem_u32(bar));
    if (bytesB) bulk_g2s(smem_u32(Bs), p.B + j0 * D, bytesB, smem_u32(bar));
  }
  for (int e = validA * D + tid; e < GT * D; e += GRAM_THREADS) As[e] = 0.0;  // points beyond the matrix: zero rows
  for (int e = validB * D + tid; e < GT * D; e += GRAM_THREADS) Bs[e] = 0.0;
  mbar_wait(smem_u32(bar), 0);
  __syncthreads();
  if (FAMILY == FAM_DOT && p.k.id == GPRC_LINEAR) {
    for (int e = tid; e < GT * D; e += GRAM_THREADS) As[e] = linear_sigma(p.k, e % D) * As[e];  // (sigma * x) * y
    __syncthreads();
  }
  if (FAMILY == FAM_DIST) {
    const double* src = (tid < GT) ? As + tid * D : Bs + (tid - GT) * D;
    double s = 0.0;
    for (int dd = 0; dd < D; ++dd) s = fma(src[dd], src[dd], s);
    (tid < GT ? na : nbv)[tid & (GT - 1)] = s;
    __syncthreads();
  }
  const int lane = tid & 31, warp = tid >> 5, warp_m = warp & 1, warp_n = warp >> 1;
  const int lk = lane & 3, lr = lane >> 2;
  double acc[4][4][2];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  for (int k0 = 0; k0 < D; k0 += 4) {
    double a[4], b[4];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) a[mb] = As[(warp_m * 32 + mb * 8 + lr) * D + k0 + lk];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) b[nb] = Bs[(warp_n * 32 + nb * 8 + lr) * D + k0 + lk];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
  }
  double rsum[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int cl = warp_n * 32 + nb * 8 + 2 * lk + r;
      const long gj = j0 + cl;
      const double cw = (p.colweights && gj < p.nB) ? p.colweights[gj] : 0.0;
      const double cs = (p.colscale && gj < p.nB) ? p.colscale[gj] : 1.0;
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        const int rl = warp_m * 32 + mb * 8 + lr;
        const long gi = i0 + rl;
        double v;
        if (gi < p.nA && gj < p.nB) {
          if (FAMILY == FAM_DIST) {
            double r2 = (na[rl] + nbv[cl]) - 2.0 * acc[mb][nb][r];
            r2 = fmax(r2, 0.0);
            if (p.symmetric && gi == gj + p.col_offset) r2 = 0.0;
            v = kfun_dist(p.k, r2);
          } else {
            v = kfun_dot(p.k, acc[mb][nb][r]);
          }
          if (p.symmetric && gi == gj + p.col_offset) v += p.diag_add;
          rsum[mb] = fma(v, cw, rsum[mb]);
          v *= cs;
        } else {
          v = (p.pad_identity && gi == gj + p.col_offset) ? 1.0 : 0.0;
        }
        if (gi < p.rows_pad && gj < p.cols_pad) p.out[gi + gj * p.ldo] = v;
      }
    }
  if (p.colweights) {
    // pmean[tile_j][i] = sum over the tile's 64 columns: the lanes of a quad, then the two column warps
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      double s = rsum[mb];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (lk == 0) red[warp_n * GT + warp_m * 32 + mb * 8 + lr] = s;
    }
    __syncthreads();
    if (tid < GT) {
      const long gi = i0 + tid;
      if (gi < p.rows_pad) p.pmean[tj * p.ldpm + gi] = red[tid] + red[GT + tid];
    }
  }
}

// which build serves these parameters (callers need the tile height to size the mean partials)
// GPRC_OPT_GRAM_DMMA: 0 never, 2 wherever eligible, 1 (default) where it measured faster on B200: the build is bound
// by the FP64 exp / pow / division of the epilogue, not by the d-dimensional contraction, so the tensor-core path only
// pays for the dot-product kernels from d = 12 (n = 16 384, lower triangle, ms: sqrexp d=8 1.14 vs 0.99 direct,
// d=32 1.72 vs 1.57; polynomial d=12 1.61 vs 1.64, d=16 1.65 vs 1.71, d=32 1.76 vs 2.08).
inline bool cov_uses_gram(const gprc_ctx* ctx, const CovParams& p) {
  const int fam = kernel_family(p.k.id);
  const bool eligible = (fam == FAM_DOT || (fam == FAM_DIST && p.k.id != GPRC_GAMMAEXP)) && p.d >= 4 &&
                        p.d <= GRAM_DMAX && p.d % 4 == 0 && !p.rowscale && !p.weights && p.rows_pad % GT == 0 &&
                        p.cols_pad % GT == 0;
  if (!eligible || ctx->opt_gram_dmma == 0) return false;
  if (ctx->opt_gram_dmma >= 2) return true;
  return fam == FAM_DOT && p.d >= 12;
}
inline int cov_tile_rows(const gprc_ctx* ctx, const CovParams& p) { return cov_uses_gram(ctx, p) ? GT : CT; }

inline int launch_cov(gprc_ctx* ctx, const CovParams& p) {
  if (cov_uses_gram(ctx, p)) {
    dim3 g((unsigned)(p.rows_pad / GT), (unsigned)(p.cols_pad / GT));
    if (g.x == 0 || g.y == 0) return 0;
    if (g.y > 65535) return set_error(-1, __FILE__, __LINE__, "cov tile grid too wide: chunk the columns");
    static bool configured[64] = {false};
    if (!configured[ctx->device & 63]) {
      GPRC_CUDA(cudaFuncSetAttribute(cov_gram_kernel<FAM_DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)gram_smem_bytes(GRAM_DMAX)));
      GPRC_CUDA(cudaFuncSetAttribute(cov_gram_kernel<FAM_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)gram_smem_bytes(GRAM_DMAX)));
      configured[ctx->device & 63] = true;
    }
    const size_t smem = gram_smem_bytes(p.d);
    if (kernel_family(p.k.id) == FAM_DIST)
      cov_gram_kernel<FAM_DIST><<<g, GRAM_THREADS, smem, ctx->stream>>>(p);
    else
      cov_gram_kernel<FAM_DOT><<<g, GRAM_THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    GPRC_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid((unsigned)(p.rows_pad / CT), (unsigned)(p.cols_pad / CT));
  if (grid.x == 0 || grid.y == 0) return 0;
  if (grid.y > 65535) return set_error(-1, __FILE__, __LINE__, "cov tile grid too wide: chunk the columns");
  switch (kernel_family(p.k.id)) {
    case FAM_DIST: cov_tile_kernel<FAM_DIST><<<grid, 256, 0, ctx->stream>>>(p); break;
    case FAM_DOT: cov_tile_kernel<FAM_DOT><<<grid, 256, 0, ctx->stream>>>(p); break;
    default: cov_tile_kernel<FAM_CONST><<<grid, 256, 0, ctx->stream>>>(p); break;
  }
  ctx->launches++;
  GPRC_CUDA(cudaGetLastError());
  return 0;
}

// mean[t] = sum_r pmean[r][t];  var[t] = kss[t] - sum_r pvar[r][t]   (fixed order: bitwise reproducible)
__global__ void finalize_predict_kernel(const double* __restrict__ pmean, long ldpm, int rows_mean,
                                        const double* __restrict__ pvar, long ldpv, int rows_var,
                                        const double* __restrict__ kss, long m, double* __restrict__ mean,
                                        double* __restrict__ var) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= m) return;
  if (mean) {
    double s = 0.0;
    for (int r = 0; r < rows_mean; ++r) s += pmean[(long)r * ldpm + t];
    mean[t] = s;
  }
  if (var) {
    double s = 0.0;
    for (int r = 0; r < rows_var; ++r) s += pvar[(long)r * ldpv + t];
    var[t] = kss[t] - s;
  }
}

}  // namespace gprc
