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
              // separate rounding of the square and of the running sum, as R's (x - y)^2 then colSums
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

inline int launch_cov(gprc_ctx* ctx, const CovParams& p) {
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
