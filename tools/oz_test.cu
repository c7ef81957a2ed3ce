// oz_test.cu -- standalone check of the INT8 (tcgen05) substitution update of csrc/ozaki.cuh against (a) a host
// emulation of exactly the same digit arithmetic and (b) the plain FP64 product; plus a timing mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/oz_test tools/oz_test.cu
//   tools/oz_test check|time <S> <n> <mc> <i> [ts 0 [dbg]]   (ts = 1: A operand from tensor memory, 2: wide 128 x 128 kernel, 4: stacked-plane kernel, 5: stacked-plane kernel on clusters of 2 with V multicast, i = pair index)   tools/oz_test digits
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

template <int S, bool TS, bool WIDE = false, bool PAIR = false, bool STACK = false>
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
  const int rows_done = PAIR ? 2 * i : i;  // block rows of V that exist before the update under test
  for (int b = 0; b < rows_done; ++b) {
    if constexpr (WIDE) oz::split_v128_kernel<S><<<dim3((unsigned)(mc / 128), 4), 256>>>(dT, mc, b, decol, dVs, (int)KB, derr);
    else oz::split_v_kernel<S><<<dim3((unsigned)(mc / 64), 4), 128>>>(dT, mc, b, decol, dVs, (int)KB, derr);
  }
  CK(cudaDeviceSynchronize());
  int herr = 0;
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  printf("split done, overflow flag = %d\n", herr);

  auto launch = [&](const oz::UpdateParams& q) {
    if constexpr (PAIR) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(2 * (mc / 64)));
      cfg.blockDim = dim3(oz::THREADS);
      cfg.dynamicSmemBytes = oz::Cfg<S>::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, oz::update_stack_kernel<S, 2>, q);
    } else if constexpr (STACK) oz::update_stack_kernel<S, 1><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(q);
    else if constexpr (WIDE) oz::update128_kernel<S><<<(unsigned)(mc / 128), oz::THREADS, oz::Cfg2<S>::SMEM_BYTES>>>(q);
    else oz::update_kernel<S, TS><<<(unsigned)(mc / 64), oz::THREADS, oz::Cfg<S>::SMEM_BYTES>>>(q);
  };
  if constexpr (PAIR) {
    CK(cudaFuncSetAttribute(oz::update_stack_kernel<S, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<S>::SMEM_BYTES));
    printf("stacked kernel on clusters of 2: block rows %d and %d against k < %d, V digits multicast; %d MMAs per k-step\n",
           2 * i, 2 * i + 1, 256 * i, oz::stack_mmas(S));
  } else if constexpr (STACK) {
    CK(cudaFuncSetAttribute(oz::update_stack_kernel<S, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<S>::SMEM_BYTES));
    printf("stacked kernel: %d MMAs per k-step\n", oz::stack_mmas(S));
  } else if constexpr (WIDE) {
    CK(cudaFuncSetAttribute(oz::update128_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg2<S>::SMEM_BYTES));
    printf("wide kernel: 128 x 128 tiles, two order passes, ring %d B, stages %d / %d\n", oz::Cfg2<S>::RING,
           oz::Cfg2<S>::STAGES0, oz::Cfg2<S>::STAGES1);
  } else {
    CK(cudaFuncSetAttribute(oz::update_kernel<S, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<S>::SMEM_BYTES));
  }
  long long* dtrace = nullptr;
  CK(cudaMalloc(&dtrace, 512 * 4 * 8));
  CK(cudaMemset(dtrace, 0, 512 * 4 * 8));
  oz::UpdateParams p{dLs, dVs, dsr, dsc, dT, mc, i, (int)KB, derr, dbg, nullptr};
  printf("S=%d %s n=%ld mc=%ld i=%d stages=%d smem=%d\n", S, TS ? "TS (A in tensor memory)" : "SS", n, mc, i,
         oz::Cfg<S>::STAGES, oz::Cfg<S>::SMEM_BYTES);

  if (timing) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int w = 0; w < 100; ++w) launch(p);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    cudaEventRecord(a);
    for (int w = 0; w < reps; ++w) launch(p);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    if (dbg & 32) {
      p.trace = dtrace;
      launch(p);
      CK(cudaDeviceSynchronize());
      std::vector<long long> tr(512 * 4);
      CK(cudaMemcpy(tr.data(), dtrace, 512 * 4 * 8, cudaMemcpyDeviceToHost));
      const long long t0 = tr[0];
      printf("trace of CTA 0 (clk since first producer issue): kt  producer_issue  stage_landed  mmas_issued\n");
      for (int kt = 0; kt < 4 * i && kt < 512; ++kt)
        if (kt < 24 || (kt >= 200 && kt < 216) || (kt >= 250 && kt < 266))
          printf("  %3d %9lld %9lld %9lld\n", kt, tr[kt * 4] - t0, tr[kt * 4 + 1] - t0, tr[kt * 4 + 2] - t0);
    }
    const double K = (PAIR ? 256.0 : 128.0) * i, flops = 2.0 * (PAIR ? 256 : 128) * mc * K;
    long long* dclk;
    CK(cudaMalloc(&dclk, 16));
    clock_probe<<<1, 1>>>(dclk);
    long long hclk[2];
    CK(cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost));
    const double mhz = (double)hclk[0] / (double)hclk[1] * 1e3;
    const long ctas = PAIR ? 2 * (mc / 64) : (WIDE ? mc / 128 : mc / 64);
    printf("update_kernel<%d,%s%s> dbg=%d: %.3f ms  -> %.1f TFLOP/s FP64-equivalent, %.2f POP/s int8 (%d products); SM clock %.0f MHz, "
           "%.0f clk per k-step and CTA\n", S, TS ? "TS" : "SS", PAIR ? " stack pair" : (STACK ? " stack" : (WIDE ? " wide" : "")), dbg, ms,
           flops / ms * 1e-9, flops * (S * (S + 1) / 2) / ms * 1e-12, S * (S + 1) / 2, mhz,
           ms * 1e-3 * mhz * 1e6 / ((PAIR ? 8.0 : 4.0) * i) / ((ctas + 147) / 148));
    return 0;
  }

  launch(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("update kernel failed: %s\n", cudaGetErrorString(e));
    return 3;
  }
  const int nrows = PAIR ? 256 : 128;
  const long row0 = PAIR ? 256L * i : 128L * i;
  std::vector<double> got((size_t)nrows * mc);
  CK(cudaMemcpy2D(got.data(), mc * 8, dT + (size_t)row0 * mc, mc * 8, mc * 8, nrows, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));

  // host emulation of the same digits
  const long K = PAIR ? 256L * i : 128L * i;
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
  for (int r = 0; r < nrows; ++r) {
    const long row = row0 + r;
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
  printf("RESULT S=%d %s%s i=%d: err flag %d, vs digit emulation max %.3e (%ld bad of %ld), vs fp64 max %.3e (|ref| max %.3e)%s\n",
         S, TS ? "TS" : "SS", PAIR ? " stack pair" : (STACK ? " stack" : (WIDE ? " wide" : "")), i, herr, max_emul, bad, (long)nrows * mc, max_fp64, max_ref,
         bad == 0 ? "  OK" : "  FAIL");
  return bad == 0 ? 0 : 1;
}

template <int N, bool TS>
static void mma_rate(int rot, int nsrc) {
  long long* d;
  cudaMalloc(&d, 128);
  cudaMemset(d, 0, 128);
  const int smem = 65536 + 8 * N * 32 + 1024, count = 20000;
  cudaFuncSetAttribute(oz::mma_rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int w = 0; w < 3; ++w) oz::mma_rate_kernel<N, TS><<<148, 128, smem>>>(count, rot, nsrc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("mma rate: 128x%dx32 int8 %s, %d accumulators, %d operand tiles: %.1f clk per MMA (floor %d) %s\n", N, TS ? "TS" : "SS",
         rot, nsrc, (double)h / count, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

// host-only self-test of the digit arithmetic (no GPU): sum_t digit_t 256^(S-1-t) reproduces the fixed-point value
// exactly, digits stay in int8 range, the per-order int32 headroom of the drain interval holds in the worst case
template <int S>
static int digits_selftest() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  long bad = 0;
  for (int rep = 0; rep < 200000; ++rep) {
    const int e = (int)(rng() % 40) - 20;
    double x = ldexp(U(rng), e);
    if (rep % 1000 == 0) x = ldexp(rep % 2000 ? 1.0 : -1.0, e) * (1.0 - 1e-16);  // the edges of the range
    if (rep % 1000 == 1) x = 0.0;
    bool ovf = false;
    const long long X = oz::to_fixed<S>(x, e, &ovf);
    const long long Y = oz::biased<S>(X);
    long long rec = 0;
    for (int s = 0; s < S; ++s) {
      const int dgt = oz::digit<S>(Y, s);
      if (dgt < -128 || dgt > 127) ++bad;
      rec = rec * 256 + dgt;
    }
    if (ovf || rec != X) ++bad;
    if (fabs(ldexp((double)X, e - (8 * S - 2)) - x) > ldexp(1.0, e - (8 * S - 2))) ++bad;  // <= 1 unit of the last digit
  }
  bool ovf = false;
  oz::to_fixed<S>(2.0, 0, &ovf);  // |x| >= 1.9 * 2^e must raise the flag
  if (!ovf) ++bad;
  if ((long long)S * oz::Cfg<S>::KC * 16384LL >= (1LL << 31)) ++bad;  // all S pairs of one order at (-128)^2 for KC steps
  if (oz::Cfg<S>::SMEM_BYTES > 227 * 1024 || S * oz::BN > oz::TMEM_COLS) ++bad;
  // stacked planes (update_stack_kernel): the MMAs of one k-step must cover every digit pair (a, b), a + b < S, exactly
  // once, each with N = 64 nb <= 256 and its accumulator columns inside the S * 64 allocated ones
  {
    int cover[8][8] = {{0}}, count = 0;
    for (int a = 0; a < S; ++a) {
      const int n1 = oz::stack_first(S, a), n2 = (S - a) - n1;
      for (int h = 0; h < 2; ++h) {
        const int b0 = h ? n1 : 0, nb = h ? n2 : n1;
        if (nb == 0) continue;
        ++count;
        if (nb < 1 || 64 * nb > 256 || (a + b0 + nb) * oz::BN > S * oz::BN) ++bad;
        for (int b = b0; b < b0 + nb; ++b) ++cover[a][b];
      }
    }
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b)
        if (cover[a][b] != ((a < S && b < S && a + b < S) ? 1 : 0)) ++bad;
    if (count != oz::stack_mmas(S)) ++bad;
  }
  printf("digits S=%d: %ld failures (KC=%d, stages=%d, smem=%d B, %d stacked MMAs per k-step)\n", S, bad, oz::Cfg<S>::KC,
         oz::Cfg<S>::STAGES, oz::Cfg<S>::SMEM_BYTES, oz::stack_mmas(S));
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "digits")
    return digits_selftest<6>() | digits_selftest<7>() | digits_selftest<8>() | digits_selftest<1>() | digits_selftest<2>();
  if (argc >= 5 && std::string(argv[1]) == "digitsof") {  // digitsof S e x...: the digits the kernels would use (host)
    const int S = atoi(argv[2]), e = atoi(argv[3]);
    for (int a = 4; a < argc; ++a) {
      const double x = strtod(argv[a], nullptr);
      bool ovf = false;
      long long Y = 0;
      int d[8] = {0};
      switch (S) {
        case 6: Y = oz::biased<6>(oz::to_fixed<6>(x, e, &ovf)); for (int s = 0; s < 6; ++s) d[s] = oz::digit<6>(Y, s); break;
        case 7: Y = oz::biased<7>(oz::to_fixed<7>(x, e, &ovf)); for (int s = 0; s < 7; ++s) d[s] = oz::digit<7>(Y, s); break;
        case 8: Y = oz::biased<8>(oz::to_fixed<8>(x, e, &ovf)); for (int s = 0; s < 8; ++s) d[s] = oz::digit<8>(Y, s); break;
        default: printf("S must be 6, 7 or 8\n"); return 64;
      }
      printf("%d", ovf ? 1 : 0);
      for (int s = 0; s < S; ++s) printf(" %d", d[s]);
      printf("\n");
    }
    return 0;
  }
  if (argc >= 2 && std::string(argv[1]) == "rate") {
    mma_rate<32, false>(1, 1);
    mma_rate<64, false>(1, 1);
    mma_rate<64, false>(7, 7);
    mma_rate<64, true>(7, 7);
    mma_rate<80, false>(6, 6);
    mma_rate<96, false>(4, 7);
    mma_rate<128, false>(1, 1);
    mma_rate<128, false>(3, 7);
    mma_rate<128, true>(3, 7);
    mma_rate<192, false>(2, 7);
    mma_rate<256, false>(1, 7);
    mma_rate<256, true>(1, 7);
    return 0;
  }
  if (argc < 6) {
    printf("usage: oz_test check|time S n mc i [ts 0 [dbg]] | oz_test digits\n");
    return 64;
  }
  const bool timing = std::string(argv[1]) == "time";
  const int S = atoi(argv[2]);
  const long n = atol(argv[3]), mc = atol(argv[4]);
  const int i = atoi(argv[5]);
  const uint32_t lbo = argc > 7 ? (uint32_t)atoi(argv[6]) : 0u, sbo = 0u;  // argv[6] = 1 selects the TS kernel
  const int dbg = argc > 8 ? atoi(argv[8]) : 0;
  if (n % 128 || mc % 128 || i < 1 || i >= n / 128 || (argc > 7 && atoi(argv[6]) == 5 && 2 * i + 1 >= n / 128)) {
    printf("bad sizes\n");
    return 64;
  }
#define OZ_DISPATCH(SS)                                                                          \
  (lbo == 5   ? run<SS, false, false, true, true>(timing, n, mc, i, lbo, sbo, dbg)               \
   : lbo == 4 ? run<SS, false, false, false, true>(timing, n, mc, i, lbo, sbo, dbg)              \
   : lbo == 2 ? run<SS, false, true>(timing, n, mc, i, lbo, sbo, dbg)                            \
   : lbo == 1 ? run<SS, true>(timing, n, mc, i, lbo, sbo, dbg)                                   \
              : run<SS, false>(timing, n, mc, i, lbo, sbo, dbg))
  switch (S) {
    case 1: return (lbo == 1 ? run<1, true>(timing, n, mc, i, lbo, sbo, dbg) : run<1, false>(timing, n, mc, i, lbo, sbo, dbg));
    case 2: return (lbo == 4 ? run<2, false, false, false, true>(timing, n, mc, i, lbo, sbo, dbg)
                             : lbo == 1 ? run<2, true>(timing, n, mc, i, lbo, sbo, dbg) : run<2, false>(timing, n, mc, i, lbo, sbo, dbg));
    case 6: return OZ_DISPATCH(6);
    case 7: return OZ_DISPATCH(7);
    case 8: return (lbo == 5   ? run<8, false, false, true, true>(timing, n, mc, i, lbo, sbo, dbg)
                    : lbo == 4 ? run<8, false, false, false, true>(timing, n, mc, i, lbo, sbo, dbg)
                    : lbo == 2 ? run<8, false, true>(timing, n, mc, i, lbo, sbo, dbg) : run<8, false>(timing, n, mc, i, lbo, sbo, dbg));
  }
  printf("S must be 1, 2, 6, 7 or 8\n");
  return 64;
}
