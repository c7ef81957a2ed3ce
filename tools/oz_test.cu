// oz_test.cu -- standalone check of the INT8 (tcgen05) substitution update of csrc/ozaki.cuh against (a) a host
// emulation of exactly the same digit arithmetic and (b) the plain FP64 product; plus a timing mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/oz_test tools/oz_test.cu
//   tools/oz_test check <S> <n> <mc> <i> [lbo sbo]        tools/oz_test time <S> <n> <mc> <i>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <string>
#include <algorithm>

#include "../gaussian-process-regression_b200/csrc/ozaki.cuh"

namespace gprc {
thread_local std::string g_last_error;
}
using namespace gprc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e__ = (x);                                                             \
    if (e__ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

__global__ void clock_probe(long long* out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < 200000ull);
  out[0] = clock64() - c0;
  out[1] = (long long)(t1 - t0);
}

template <int S>
static int run(bool timing, long n, long mc, int i, uint32_t lbo, uint32_t sbo, int dbg) {
  const long n_pad = n, KB = n_pad / 32;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  std::vector<double> L((size_t)n_pad * n_pad, 0.0), T((size_t)mc * n_pad), kss(mc);
  for (long c = 0; c < n_pad; ++c)
    for (long r = c; r < n_pad; ++r) L[r + c * n_pad] = U(rng) * (r == c ? 1.0 : 0.37);
  for (long t = 0; t < mc; ++t) kss[t] = 0.25 + 3.0 * (double)(t % 7);
  for (long k = 0; k < n_pad; ++k)
    for (long t = 0; t < mc; ++t) T[t + k * mc] = U(rng) * sqrt(kss[t]);

  double *dL, *dT, *dkss, *dsr, *dsc;
  int *derow, *decol, *derr;
  int8_t *dLs, *dVs;
  CK(cudaMalloc(&dL, L.size() * 8));
  CK(cudaMalloc(&dT, T.size() * 8));
  CK(cudaMalloc(&dkss, mc * 8));
  CK(cudaMalloc(&dsr, n_pad * 8));
  CK(cudaMalloc(&dsc, mc * 8));
  CK(cudaMalloc(&derow, n_pad * 4));
  CK(cudaMalloc(&decol, mc * 4));
  CK(cudaMalloc(&derr, 4));
  CK(cudaMalloc(&dLs, (size_t)n_pad * n_pad * S));
  CK(cudaMalloc(&dVs, (size_t)mc * n_pad * S));
  CK(cudaMemset(dLs, 0, (size_t)n_pad * n_pad * S));
  CK(cudaMemset(dVs, 0, (size_t)mc * n_pad * S));
  CK(cudaMemset(derr, 0, 4));
  CK(cudaMemcpy(dL, L.data(), L.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dT, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dkss, kss.data(), mc * 8, cudaMemcpyHostToDevice));

  const int nt = (int)(n_pad / 128);
  oz::rowmax_kernel<<<nt, 256>>>(dL, n_pad, derow, dsr);
  oz::split_l_kernel<S><<<dim3(4 * nt, nt), 256>>>(dL, n_pad, derow, dLs, (int)KB, derr);
  oz::colscale_kernel<<<(unsigned)((mc + 255) / 256), 256>>>(dkss, mc, mc, decol, dsc);
  for (int b = 0; b < i; ++b)
    oz::split_v_kernel<S><<<dim3((unsigned)(mc / 64), 4), 128>>>(dT, mc, b, decol, dVs, (int)KB, derr);
  CK(cudaDeviceSynchronize());
  int herr = 0;
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  printf("split done, overflow flag = %d\n", herr);

  CK(cudaFuncSetAttribute(oz::update_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<S>::SMEM_BYTES));
  long long* dtrace = nullptr;
  CK(cudaMalloc(&dtrace, 512 * 4 * 8));
  CK(cudaMemset(dtrace, 0, 512 * 4 * 8));
  oz::UpdateParams p{dLs, dVs, dsr, dsc, dT, mc, i, (int)KB, derr, dbg, nullptr};
  printf("S=%d n=%ld mc=%ld i=%d stages=%d smem=%d lbo=%u sbo=%u\n", S, n, mc, i, oz::Cfg<S>::STAGES,
         oz::Cfg<S>::SMEM_BYTES, lbo, sbo);

  if (timing) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int w = 0; w < 100; ++w) oz::update_kernel<S><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(p);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    cudaEventRecord(a);
    for (int w = 0; w < reps; ++w) oz::update_kernel<S><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(p);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    if (dbg & 32) {
      p.trace = dtrace;
      oz::update_kernel<S><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(p);
      CK(cudaDeviceSynchronize());
      std::vector<long long> tr(512 * 4);
      CK(cudaMemcpy(tr.data(), dtrace, 512 * 4 * 8, cudaMemcpyDeviceToHost));
      const long long t0 = tr[0];
      printf("trace of CTA 0 (clk since first producer issue): kt  producer_issue  stage_landed  mmas_issued\n");
      for (int kt = 0; kt < 4 * i && kt < 512; ++kt)
        if (kt < 24 || (kt >= 200 && kt < 216) || (kt >= 250 && kt < 266))
          printf("  %3d %9lld %9lld %9lld\n", kt, tr[kt * 4] - t0, tr[kt * 4 + 1] - t0, tr[kt * 4 + 2] - t0);
    }
    const double K = 128.0 * i, flops = 2.0 * 128 * mc * K;
    long long* dclk;
    CK(cudaMalloc(&dclk, 16));
    clock_probe<<<1, 1>>>(dclk);
    long long hclk[2];
    CK(cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost));
    const double mhz = (double)hclk[0] / (double)hclk[1] * 1e3;
    printf("update_kernel<%d> dbg=%d: %.3f ms  -> %.1f TFLOP/s FP64-equivalent, %.2f POP/s int8 (%d products); SM clock %.0f MHz, "
           "%.0f clk per k-step\n", S, dbg, ms, flops / ms * 1e-9, flops * (S * (S + 1) / 2) / ms * 1e-12, S * (S + 1) / 2, mhz,
           ms * 1e-3 * mhz * 1e6 / (4.0 * i) / ((mc / 64 + 147) / 148));
    return 0;
  }

  oz::update_kernel<S><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("update kernel failed: %s\n", cudaGetErrorString(e));
    return 3;
  }
  std::vector<double> got((size_t)128 * mc);
  CK(cudaMemcpy2D(got.data(), mc * 8, dT + (size_t)i * 128 * mc, mc * 8, mc * 8, 128, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));

  // host emulation of the same digits
  const long K = 128L * i;
  std::vector<int> erow(n_pad), ecol(mc);
  for (long r = 0; r < n_pad; ++r) {
    double mx = 0;
    for (long k = 0; k < (r / 128) * 128; ++k) mx = std::max(mx, fabs(L[r + k * n_pad]));
    erow[r] = oz::exponent_of(mx);
  }
  for (long t = 0; t < mc; ++t) ecol[t] = oz::exponent_of(sqrt(kss[t]) * (1.0 + 1e-9));
  std::vector<int8_t> dv((size_t)S * K * mc);  // [s][k][t]
  bool ovf = false;
  for (long k = 0; k < K; ++k)
    for (long t = 0; t < mc; ++t) {
      long long Y = oz::biased<S>(oz::to_fixed<S>(T[t + k * mc], ecol[t], &ovf));
      for (int s = 0; s < S; ++s) dv[((size_t)s * K + k) * mc + t] = (int8_t)oz::digit<S>(Y, s);
    }
  double max_emul = 0, max_fp64 = 0, max_ref = 0;
  long bad = 0;
  std::vector<long long> acc((size_t)S * mc);
  std::vector<double> ref(mc);
  for (int r = 0; r < 128; ++r) {
    const long row = 128L * i + r;
    std::fill(acc.begin(), acc.end(), 0LL);
    std::fill(ref.begin(), ref.end(), 0.0);
    for (long k = 0; k < K; ++k) {
      const double lv = L[row + k * n_pad];
      long long Y = oz::biased<S>(oz::to_fixed<S>(lv, erow[row], &ovf));
      int da[8];
      for (int s = 0; s < S; ++s) da[s] = oz::digit<S>(Y, s);
      for (int a = 0; a < S; ++a)
        for (int b = 0; b < S - a; ++b) {
          const int8_t* dvb = &dv[((size_t)b * K + k) * mc];
          long long* ao = &acc[(size_t)(a + b) * mc];
          const int x = da[a];
          for (long t = 0; t < mc; ++t) ao[t] += x * (int)dvb[t];
        }
      const double* tv = &T[k * mc];
      for (long t = 0; t < mc; ++t) ref[t] += lv * tv[t];
    }
    const double sr = ldexp(1.0, erow[row] - 6);
    for (long t = 0; t < mc; ++t) {
      double h = (double)acc[(size_t)(S - 1) * mc + t];
      for (int o = S - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[(size_t)o * mc + t]);
      const double sc = ldexp(1.0, ecol[t] - 6);
      const double t0 = T[t + row * mc];
      const double want = fma(-(sr * sc), h, t0);
      const double g = got[(size_t)r * mc + t];
      const double d1 = fabs(g - want), d2 = fabs(g - (t0 - ref[t]));
      max_emul = std::max(max_emul, d1);
      max_fp64 = std::max(max_fp64, d2);
      max_ref = std::max(max_ref, fabs(ref[t]));
      if (d1 > 1e-12 * (1.0 + fabs(want))) {
        if (bad < 6) printf("  mismatch r=%d t=%ld got=%.17g want=%.17g fp64=%.17g\n", r, t, g, want, t0 - ref[t]);
        ++bad;
      }
    }
  }
  printf("RESULT S=%d i=%d lbo=%u sbo=%u: err flag %d, vs digit emulation max %.3e (%ld bad of %ld), vs fp64 max %.3e (|ref| max %.3e)%s\n",
         S, i, lbo, sbo, herr, max_emul, bad, 128 * mc, max_fp64, max_ref, bad == 0 ? "  OK" : "  FAIL");
  return bad == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    printf("usage: oz_test check|time S n mc i [lbo sbo]\n");
    return 64;
  }
  const bool timing = std::string(argv[1]) == "time";
  const int S = atoi(argv[2]);
  const long n = atol(argv[3]), mc = atol(argv[4]);
  const int i = atoi(argv[5]);
  const uint32_t lbo = argc > 7 ? (uint32_t)atoi(argv[6]) : 128u, sbo = argc > 7 ? (uint32_t)atoi(argv[7]) : 256u;
  const int dbg = argc > 8 ? atoi(argv[8]) : 0;
  if (n % 128 || mc % 64 || i < 1 || i >= n / 128) {
    printf("bad sizes\n");
    return 64;
  }
  switch (S) {
    case 1: return run<1>(timing, n, mc, i, lbo, sbo, dbg);
    case 2: return run<2>(timing, n, mc, i, lbo, sbo, dbg);
    case 6: return run<6>(timing, n, mc, i, lbo, sbo, dbg);
    case 7: return run<7>(timing, n, mc, i, lbo, sbo, dbg);
    case 8: return run<8>(timing, n, mc, i, lbo, sbo, dbg);
  }
  printf("S must be 1, 2, 6, 7 or 8\n");
  return 64;
}
