// optim.hpp -- the optimisers behind fit() inside the library: restatements of the two engines
// that R/fit.R:143-160 reaches through stats::optim / optimize,
//   Brent_fmin  (R src/appl/fmin.c: optimize(), optim(method = "Brent"))
//   vmmin       (R src/appl/optim.c: optim(method = "BFGS"))
// and of the reference's error-tolerant wrapper optim_until_error (R/fit.R:47-69).  Host code: the optimiser is scalar
// control flow; what it drives -- one Cholesky-based objective or gradient per evaluation -- runs on the device without
// the data ever leaving it (gprc_fit_family).  The arithmetic follows the same statement order as the Python mirror
// (_optim.py), which the CPU tier pins to R's documented example(optim) result, so both produce identical iterates.
//
// Provenance / licence: the ALGORITHMS are Brent's localmin (R. P. Brent, "Algorithms for Minimization without
// Derivatives", 1973, ch. 5 -- the Fortran `fmin` on netlib is public domain) and the variable-metric method of J. C. Nash,
// "Compact Numerical Methods for Computers", 2nd ed. 1990, Algorithm 21 (on which R's vmmin is based).  R's own C
// sources (src/appl/fmin.c, src/appl/optim.c) are GPL-2 | GPL-3; they were READ to reproduce R's constants, tolerances,
// step acceptance rule and counters exactly (so that trajectories match R's optim digit for digit) and are cited above
// for that reason -- this file is an independent restatement written for this library, not a copy of those files.
// A downstream that needs licence certainty for trajectory-identical behaviour should treat optim.hpp and _optim.py as
// derived from GPL-2+ material; nothing else in libgprc depends on them (fit() can be driven by any optimiser).
#pragma once
#include <cmath>
#include <functional>
#include <limits>
#include <utility>
#include <vector>

namespace gprc {
namespace optim {

struct ObjectiveError {};  // the objective / gradient "threw" (R: stopifnot, chol, solve)

using Fn1 = std::function<double(double)>;
using FnN = std::function<double(const double*)>;
using GrN = std::function<void(const double*, double*)>;

inline double brent_fmin(const Fn1& f, double ax, double bx, double tol) {
  const double c = (3.0 - std::sqrt(5.0)) * 0.5;
  const double eps = std::sqrt(std::numeric_limits<double>::epsilon());
  double a = ax, b = bx;
  double v = a + c * (b - a);
  double w = v, x = v;
  double d = 0.0, e = 0.0;
  double fx = f(x);
  double fv = fx, fw = fx;
  const double tol3 = tol / 3.0;
  for (;;) {
    const double xm = (a + b) * 0.5;
    const double tol1 = eps * std::fabs(x) + tol3;
    const double t2 = tol1 * 2.0;
    if (std::fabs(x - xm) <= t2 - (b - a) * 0.5) break;
    double p = 0.0, q = 0.0, r = 0.0;
    if (std::fabs(e) > tol1) {
      r = (x - w) * (fx - fv);
      q = (x - v) * (fx - fw);
      p = (x - v) * q - (x - w) * r;
      q = (q - r) * 2.0;
      if (q > 0.0) p = -p;
      else q = -q;
      r = e;
      e = d;
    }
    double u;
    if (std::fabs(p) >= std::fabs(q * 0.5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
      e = (x < xm) ? (b - x) : (a - x);
      d = c * e;
    } else {
      d = p / q;
      u = x + d;
      if (u - a < t2 || b - u < t2) {
        d = tol1;
        if (x >= xm) d = -d;
      }
    }
    if (std::fabs(d) >= tol1) u = x + d;
    else if (d > 0.0) u = x + tol1;
    else u = x - tol1;
    const double fu = f(u);
    if (fu <= fx) {
      if (u < x) b = x;
      else a = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) a = u;
      else b = u;
      if (fu <= fw || w == x) {
        v = w; fv = fw;
        w = u; fw = fu;
      } else if (fu <= fv || v == x || v == w) {
        v = u; fv = fu;
      }
    }
  }
  return x;
}

struct VmminResult {
  double value = 0.0;
  int fail = 0, fncount = 0, grcount = 0;
};

// b: start on entry, minimiser on exit.  Throws ObjectiveError if the initial value is not finite (R: error()).
inline VmminResult vmmin(double* b, int n, const FnN& fminfn, const GrN& fmingr, int maxit = 100,
                         double abstol = -std::numeric_limits<double>::infinity(),
                         double reltol = std::sqrt(std::numeric_limits<double>::epsilon())) {
  const double stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
  std::vector<double> B((size_t)n * n, 0.0), g(n), t(n, 0.0), X(n, 0.0), c(n, 0.0);
  auto Bm = [&](int i, int j) -> double& { return B[(size_t)i * n + j]; };
  double f = fminfn(b);
  if (!std::isfinite(f)) throw ObjectiveError{};
  double Fmin = f;
  int funcount = 1, gradcount = 1;
  fmingr(b, g.data());
  int iter = 1;
  int ilast = gradcount;
  int count = 0;
  for (;;) {
    if (ilast == gradcount) {
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Bm(i, j) = (i == j) ? 1.0 : 0.0;
    }
    for (int i = 0; i < n; ++i) {
      X[i] = b[i];
      c[i] = g[i];
    }
    double gradproj = 0.0;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j <= i; ++j) s -= Bm(i, j) * g[j];
      for (int j = i + 1; j < n; ++j) s -= Bm(j, i) * g[j];
      t[i] = s;
      gradproj += s * g[i];
    }
    if (gradproj < 0.0) {
      double steplength = 1.0;
      bool accpoint = false;
      do {
        count = 0;
        for (int i = 0; i < n; ++i) {
          b[i] = X[i] + steplength * t[i];
          if (reltest + X[i] == reltest + b[i]) ++count;
        }
        if (count < n) {
          f = fminfn(b);
          ++funcount;
          accpoint = std::isfinite(f) && (f <= Fmin + gradproj * steplength * acctol);
          if (!accpoint) steplength *= stepredn;
        }
      } while (!(count == n || accpoint));
      const bool enough = (f > abstol) && std::fabs(f - Fmin) > reltol * (std::fabs(Fmin) + reltol);
      if (!enough) {
        count = n;
        Fmin = f;
      }
      if (count < n) {
        Fmin = f;
        fmingr(b, g.data());
        ++gradcount;
        ++iter;
        double D1 = 0.0;
        for (int i = 0; i < n; ++i) {
          t[i] = steplength * t[i];
          c[i] = g[i] - c[i];
          D1 += t[i] * c[i];
        }
        if (D1 > 0) {
          double D2 = 0.0;
          for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j <= i; ++j) s += Bm(i, j) * c[j];
            for (int j = i + 1; j < n; ++j) s += Bm(j, i) * c[j];
            X[i] = s;
            D2 += s * c[i];
          }
          D2 = 1.0 + D2 / D1;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) Bm(i, j) += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
        } else {
          ilast = gradcount;
        }
      } else {
        if (ilast < gradcount) {
          count = 0;
          ilast = gradcount;
        }
      }
    } else {
      count = 0;
      if (ilast == gradcount) count = n;
      else ilast = gradcount;
    }
    if (iter >= maxit) break;
    if (gradcount - ilast > 2 * n) ilast = gradcount;
    if (!(count != n || ilast != gradcount)) break;
  }
  VmminResult r;
  r.value = Fmin;
  r.fail = (iter < maxit) ? 0 : 1;
  r.fncount = funcount;
  r.grcount = gradcount;
  return r;
}

// optim_until_error(start, f, method, ...) with control = list(fnscale = -1), R/fit.R:47-69,149-150,158: MAXIMISES f.
// An objective error scores -10000; a gradient error (or an error raised by the optimiser itself) ends the search and
// the best recorded evaluation wins; with nothing recorded the start is returned with its own value.
struct UntilErrorResult {
  std::vector<double> par;
  double value = 0.0;
  long fn_evaluations = 0, gr_evaluations = 0;
};

inline UntilErrorResult optim_until_error(const std::vector<double>& start, const FnN& f, const GrN* gr, bool brent,
                                          double lower, double upper) {
  const int n = (int)start.size();
  std::vector<std::pair<std::vector<double>, double>> record;
  UntilErrorResult out;
  auto f_new = [&](const double* p) -> double {
    ++out.fn_evaluations;
    double v;
    try {
      v = f(p);
    } catch (const ObjectiveError&) {
      return -10000.0;
    }
    if (!(v == -10000.0)) {
      std::vector<double> at((size_t)n);
      for (int i = 0; i < n; ++i) at[i] = p[i];
      record.emplace_back(std::move(at), v);
    }
    return v;
  };
  try {
    if (brent) {
      const double x = brent_fmin([&](double p) { return f_new(&p) / -1.0; }, lower, upper,
                                  std::sqrt(std::numeric_limits<double>::epsilon()));
      out.par = {x};
      out.value = f_new(&x);  // optim() re-evaluates fn at the minimiser
      return out;
    }
    std::vector<double> b(start);
    GrN g = [&](const double* p, double* gout) {
      ++out.gr_evaluations;
      (*gr)(p, gout);
      for (int i = 0; i < n; ++i) gout[i] = gout[i] / -1.0;
    };
    VmminResult r = vmmin(b.data(), n, [&](const double* p) { return f_new(p) / -1.0; }, g);
    out.par = b;
    out.value = r.value * -1.0;
    return out;
  } catch (const ObjectiveError&) {
    if (record.empty()) {
      out.par = start;
      try {
        out.value = f(start.data());
      } catch (const ObjectiveError&) {
        out.value = -10000.0;
      }
      return out;
    }
    size_t best = 0;  // which.max: the first maximum
    for (size_t i = 1; i < record.size(); ++i)
      if (record[i].second > record[best].second) best = i;
    out.par = record[best].first;
    out.value = record[best].second;
    return out;
  }
}

}  // namespace optim
}  // namespace gprc
