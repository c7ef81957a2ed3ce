// Host build of quad.cuh for the CPU test tier (compared with scipy.integrate.quad = QUADPACK dqagie).
#include "quad.cuh"

extern "C" void gprc_quad_host(const double* mean, const double* sd, long m, double* out, double* abserr, int* ier,
                               int* neval, int* last) {
  for (long i = 0; i < m; ++i) {
    const gprc_quad::QuadResult r = gprc_quad::logistic_gaussian(mean[i], sd[i]);
    out[i] = r.result;
    if (abserr) abserr[i] = r.abserr;
    if (ier) ier[i] = r.ier;
    if (neval) neval[i] = r.neval;
    if (last) last[i] = r.last;
  }
}

// other integrands, only to exercise every branch of the port (bisection ordering, extrapolation, error flags)
struct TestIntegrand {
  int kind;
  double p;
  double operator()(double z) const {
    switch (kind) {
      case 0: return 1.0 / (1.0 + z * z);                                   // smooth, heavy tails
      case 1: return exp(-fabs(z)) / sqrt(fabs(z) + 1e-300);                // integrable singularity at 0
      case 2: return cos(p * z) * exp(-z * z);                              // oscillatory
      case 3: return exp(-fabs(z - p)) * log(fabs(z - p) + 1e-300);         // log singularity off centre
      case 4: return 1.0 / (1.0 + fabs(z));                                 // divergent
      default: return fabs(z) < p ? 1.0 : 0.0;                              // discontinuous
    }
  }
};

extern "C" void gprc_quad_host_test(int kind, double p, double epsabs, double epsrel, double* out) {
  TestIntegrand f{kind, p};
  const gprc_quad::QuadResult r = gprc_quad::qagi(f, epsabs, epsrel);
  out[0] = r.result;
  out[1] = r.abserr;
  out[2] = r.neval;
  out[3] = r.ier;
  out[4] = r.last;
}
