// gpc.cuh -- kernels of the GPC Laplace-approximation Newton loop (north_star subsystem 4), R/GPCclass.R:76-103.
//
// Per iteration the reference computes  P = sigmoid(f); W = (1 - P) P; B = I + (sqrt(W) %o% sqrt(W)) * K;
// L = chol(B)';  b = W f + (y + 1)/2 - P;  a = b - sqrt(W) * (L' \ (L \ (sqrt(W) * (K b))));  f = K a;
// objective = -sum(a f)/2 - sum(log(1 + exp(-y f))).
// Here the sigmoid/W evaluation is fused into the kernel that forms B (each 64 x 64 tile evaluates the 128 values of
// sqrt(W) it needs from f in shared memory; the first tile column also emits P-derived vectors b and sqrt(W)).
#pragma once
#include "common.cuh"

namespace gprc {

// private$.sigmoid, R/GPCclass.R:63
__device__ __forceinline__ double r_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// B[i + j ld] = [i == j] + (sw_i sw_j) K[i + j ld] on 64 x 64 tiles of the lower triangle, identity in the padding.
// Tiles with blockIdx.y == 0 also write sw = sqrt(W), bvec = W f + (y + 1)/2 - P and (optionally) gradl = (y+1)/2 - P.
__global__ void __launch_bounds__(256) gpc_build_B_kernel(const double* __restrict__ K, long ldk,
                                                          const double* __restrict__ f, const double* __restrict__ y,
                                                          long n, long n_pad, double* __restrict__ Bm, long ldb,
                                                          double* __restrict__ sw, double* __restrict__ bvec,
                                                          double* __restrict__ gradl) {
  const long ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  __shared__ double swi[64], swj[64];
  const int tid = threadIdx.x;
  if (tid < 128) {
    const bool is_row = tid < 64;
    const long g = (is_row ? ti : tj) * 64 + (tid & 63);
    double s = 0.0;
    if (g < n) {
      const double P = r_sigmoid(f[g]);
      const double W = (1.0 - P) * P;  // R/GPCclass.R:79
      s = sqrt(W);
      if (is_row && tj == 0) {
        sw[g] = s;
        const double gl = (y[g] + 1.0) / 2.0 - P;
        bvec[g] = W * f[g] + gl;  // R/GPCclass.R:81
        if (gradl) gradl[g] = gl;
      }
    } else if (is_row && tj == 0 && g < n_pad) {
      sw[g] = 0.0;
      bvec[g] = 0.0;
      if (gradl) gradl[g] = 0.0;
    }
    (is_row ? swi : swj)[tid & 63] = s;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int qb = 0; qb < 4; ++qb) {
    const long gj = tj * 64 + ty * 4 + qb;
#pragma unroll
    for (int qa = 0; qa < 4; ++qa) {
      const long gi = ti * 64 + tx + 16 * qa;
      double v;
      if (gi < n && gj < n) {
        v = (swi[tx + 16 * qa] * swj[ty * 4 + qb]) * K[gi + gj * ldk];
        if (gi == gj) v = 1.0 + v;
      } else {
        v = (gi == gj) ? 1.0 : 0.0;
      }
      Bm[gi + gj * ldb] = v;
    }
  }
}

// elementwise helpers of the Newton step
// c = sw * Kb                                                     (R/GPCclass.R:82, right-hand side)
__global__ void vec_mul_kernel(const double* __restrict__ a, const double* __restrict__ b, long n,
                               double* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}
// a = b - sw * t                                                  (R/GPCclass.R:84)
__global__ void gpc_a_kernel(const double* __restrict__ b, const double* __restrict__ sw, const double* __restrict__ t,
                             long n, double* __restrict__ a) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) a[i] = b[i] - sw[i] * t[i];
}
// objective = -sum(a * f)/2 - sum(log(1 + exp(-y * f)))           (R/GPCclass.R:86; no log1p/softplus: A.12)
__global__ void __launch_bounds__(1024) gpc_objective_kernel(const double* __restrict__ a, const double* __restrict__ f,
                                                             const double* __restrict__ y, long n,
                                                             double* __restrict__ out) {
  __shared__ double s1[1024], s2[1024];
  const int tid = threadIdx.x;
  const long chunk = (n + 1023) / 1024;
  const long lo = tid * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
  double af = 0.0, ll = 0.0;
  for (long i = lo; i < hi; ++i) {
    af += a[i] * f[i];
    ll += log(1.0 + exp(-y[i] * f[i]));
  }
  s1[tid] = af;
  s2[tid] = ll;
  __syncthreads();
  if (tid == 0) {
    double taf = 0.0, tll = 0.0;
    for (int t = 0; t < 1024; ++t) {
      taf += s1[t];
      tll += s2[t];
    }
    out[0] = -taf / 2.0 - tll;
  }
}

}  // namespace gprc
