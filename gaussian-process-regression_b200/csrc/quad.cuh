// quad.cuh -- QUADPACK dqagi (infinite range, 15-point transformed Gauss-Kronrod, epsilon-algorithm extrapolation)
// restated for host and device, specialised by the integrand functor.  It replaces the per-test-point
//     integrate(function(z) sigmoid(z) * dnorm(z, mean = fs_bar[i], sd = Vfs[i]), -Inf, Inf)$value
// of GPC$predict_class (R/GPCclass.R:116-117).  R's integrate() is a C translation of QUADPACK's dqagie/dqk15i/
// dqpsrt/dqelg (R src/appl/integrate.c), called with rel.tol = abs.tol = .Machine$double.eps^0.25 and 100
// subdivisions; this file follows the published QUADPACK routines (Piessens, de Doncker-Kapenga, Ueberhuber,
// Kahaner 1983) statement by statement, INCLUDING their behaviour on narrow peaks (a density much narrower than the
// transformed sampling grid is missed and the result is ~0 with ier = 0: SURVEY.md A.1) -- an "accurate" quadrature
// would flip class labels against the reference.
// The file compiles as plain C++ (tests build a host library from it and compare with scipy.integrate.quad, which
// wraps the same QUADPACK routine) and as CUDA (one thread per test point).
//
// Provenance / licence: QUADPACK (R. Piessens, E. de Doncker-Kapenga, C. W. Ueberhuber, D. K. Kahaner, "QUADPACK: A
// Subroutine Package for Automatic Integration", Springer 1983; Fortran sources on netlib) is in the PUBLIC DOMAIN.  This
// file is a new C++/CUDA restatement of the published algorithms dqagie / dqk15i / dqpsrt / dqelg; no code was taken from
// R's integrate.c (GPL-2) or from /root/reference (which holds no native code).  It is part of libgprc and carries the
// repository's licence.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define GPRC_HD __host__ __device__
#else
#define GPRC_HD
#endif

namespace gprc_quad {

constexpr int LIMIT = 100;  // R's `subdivisions`

// 15-point Kronrod abscissae / weights and the embedded 7-point Gauss weights (QUADPACK dqk15i)
GPRC_HD inline double xgk(int j) {
  const double v[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                       0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                       0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                       0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
  return v[j];
}
GPRC_HD inline double wgk(int j) {
  const double v[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                       0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                       0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                       0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
  return v[j];
}
GPRC_HD inline double wg(int j) {
  const double v[8] = {0.0, 0.129484966168869693270611432679082, 0.0, 0.279705391489276667901467771423780,
                       0.0, 0.381830050505118944950369775488975, 0.0, 0.417959183673469387755102040816327};
  return v[j];
}

// dqk15i for the range (-inf, +inf) (inf = 2, bound = 0): integrates f over the transformed interval (a, b) of (0, 1]
template <class F>
GPRC_HD inline void qk15i(const F& f, double a, double b, double& result, double& abserr, double& resabs,
                          double& resasc) {
  const double epmach = DBL_EPSILON, uflow = DBL_MIN;
  double fv1[7], fv2[7];
  const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a);
  const double tabsc1 = (1.0 - centr) / centr;
  double fval1 = f(tabsc1) + f(-tabsc1);
  const double fc = (fval1 / centr) / centr;
  double resg = wg(7) * fc, resk = wgk(7) * fc;
  resabs = fabs(resk);
  for (int j = 0; j < 7; ++j) {
    const double absc = hlgth * xgk(j), absc1 = centr - absc, absc2 = centr + absc;
    const double t1 = (1.0 - absc1) / absc1, t2 = (1.0 - absc2) / absc2;
    fval1 = f(t1) + f(-t1);
    double fval2 = f(t2) + f(-t2);
    fval1 = (fval1 / absc1) / absc1;
    fval2 = (fval2 / absc2) / absc2;
    fv1[j] = fval1;
    fv2[j] = fval2;
    const double fsum = fval1 + fval2;
    resg += wg(j) * fsum;
    resk += wgk(j) * fsum;
    resabs += wgk(j) * (fabs(fval1) + fabs(fval2));
  }
  const double reskh = resk * 0.5;
  resasc = wgk(7) * fabs(fc - reskh);
  for (int j = 0; j < 7; ++j) resasc += wgk(j) * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
  result = resk * hlgth;
  resasc *= hlgth;
  resabs *= hlgth;
  abserr = fabs((resk - resg) * hlgth);
  if (resasc != 0.0 && abserr != 0.0) abserr = resasc * fmin(1.0, pow(200.0 * abserr / resasc, 1.5));
  if (resabs > uflow / (50.0 * epmach)) abserr = fmax((epmach * 50.0) * resabs, abserr);
}

// dqpsrt: maintains the descending ordering of the error estimates (1-based indices as in QUADPACK)
GPRC_HD inline void qpsrt(int limit, int last, int& maxerr, double& ermax, const double* elist, int* iord, int& nrmax) {
  if (last <= 2) {
    iord[1] = 1;
    iord[2] = 2;
  } else {
    const double errmax = elist[maxerr];
    if (nrmax != 1) {
      const int ido = nrmax - 1;
      for (int i = 1; i <= ido; ++i) {
        const int isucc = iord[nrmax - 1];
        if (errmax <= elist[isucc]) break;
        iord[nrmax] = isucc;
        --nrmax;
      }
    }
    int jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    const double errmin = elist[last];
    const int jbnd = jupbn - 1, ibeg = nrmax + 1;
    bool placed = false;
    int i = ibeg;
    for (; i <= jbnd; ++i) {
      const int isucc = iord[i];
      if (errmax >= elist[isucc]) {
        placed = true;
        break;
      }
      iord[i - 1] = isucc;
    }
    if (!placed) {
      iord[jbnd] = maxerr;
      iord[jupbn] = last;
    } else {
      iord[i - 1] = maxerr;
      int k = jbnd;
      bool done = false;
      for (int j = i; j <= jbnd; ++j) {
        const int isucc = iord[k];
        if (errmin < elist[isucc]) {
          iord[k + 1] = last;
          done = true;
          break;
        }
        iord[k + 1] = isucc;
        --k;
      }
      if (!done) iord[i] = last;
    }
  }
  maxerr = iord[nrmax];
  ermax = elist[maxerr];
}

// dqelg: the epsilon algorithm (1-based table epstab[1..52], res3la[1..3])
GPRC_HD inline void qelg(int& n, double* epstab, double& result, double& abserr, double* res3la, int& nres) {
  const double epmach = DBL_EPSILON, oflow = DBL_MAX;
  ++nres;
  abserr = oflow;
  result = epstab[n];
  if (n >= 3) {
    const int limexp = 50;
    epstab[n + 2] = epstab[n];
    const int newelm = (n - 1) / 2;
    epstab[n] = oflow;
    const int num = n;
    int k1 = n;
    bool converged = false;
    for (int i = 1; i <= newelm; ++i) {
      const int k2 = k1 - 1, k3 = k1 - 2;
      double res = epstab[k1 + 2];
      const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
      const double e1abs = fabs(e1), delta2 = e2 - e1, err2 = fabs(delta2), tol2 = fmax(fabs(e2), e1abs) * epmach;
      const double delta3 = e1 - e0, err3 = fabs(delta3), tol3 = fmax(e1abs, fabs(e0)) * epmach;
      if (!(err2 > tol2 || err3 > tol3)) {
        // e0, e1 and e2 are equal to within machine accuracy: convergence is assumed
        result = res;
        abserr = err2 + err3;
        abserr = fmax(abserr, 5.0 * epmach * fabs(result));
        converged = true;
        break;
      }
      const double e3 = epstab[k1];
      epstab[k1] = e1;
      const double delta1 = e1 - e3, err1 = fabs(delta1), tol1 = fmax(e1abs, fabs(e3)) * epmach;
      // if two elements are very close to each other, omit a part of the table by adjusting the value of n
      if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
        n = i + i - 1;
        break;
      }
      const double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
      const double epsinf = fabs(ss * e1);
      // test to detect irregular behaviour in the table, and eventually omit a part of the table
      if (!(epsinf > 1.0e-4)) {
        n = i + i - 1;
        break;
      }
      res = e1 + 1.0 / ss;
      epstab[k1] = res;
      k1 -= 2;
      const double error = err2 + fabs(res - e2) + err3;
      if (error > abserr) continue;
      abserr = error;
      result = res;
    }
    if (converged) return;  // label 100 was already applied above
    // shift the table
    if (n == limexp) n = 2 * (limexp / 2) - 1;
    int ib = ((num / 2) * 2 == num) ? 2 : 1;
    const int ie = newelm + 1;
    for (int i = 1; i <= ie; ++i) {
      const int ib2 = ib + 2;
      epstab[ib] = epstab[ib2];
      ib = ib2;
    }
    if (num != n) {
      int indx = num - n + 1;
      for (int i = 1; i <= n; ++i) {
        epstab[i] = epstab[indx];
        ++indx;
      }
    }
    if (nres >= 4) {
      abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
      res3la[1] = res3la[2];
      res3la[2] = res3la[3];
      res3la[3] = result;
    } else {
      res3la[nres] = result;
      abserr = oflow;
    }
  }
  abserr = fmax(abserr, 5.0 * epmach * fabs(result));
}

struct QuadResult {
  double result, abserr;
  int neval, ier, last;
};

// dqagie with inf = 2 (both limits infinite), limit = 100
template <class F>
GPRC_HD inline QuadResult qagi(const F& f, double epsabs, double epsrel) {
  const double epmach = DBL_EPSILON, uflow = DBL_MIN, oflow = DBL_MAX;
  const int limit = LIMIT;
  double alist[LIMIT + 1], blist[LIMIT + 1], rlist[LIMIT + 1], elist[LIMIT + 1];
  int iord[LIMIT + 2];
  double rlist2[53], res3la[4];
  QuadResult R;
  int ier = 0, last = 0;
  double result = 0.0, abserr = 0.0;
  alist[1] = 0.0;
  blist[1] = 1.0;
  rlist[1] = 0.0;
  elist[1] = 0.0;
  iord[1] = 0;
  if (epsabs <= 0.0 && epsrel < fmax(50.0 * epmach, 0.5e-28)) {
    R.result = 0.0;
    R.abserr = 0.0;
    R.neval = 0;
    R.ier = 6;
    R.last = 0;
    return R;
  }
  double defabs, resabs;
  qk15i(f, 0.0, 1.0, result, abserr, defabs, resabs);
  last = 1;
  rlist[1] = result;
  elist[1] = abserr;
  iord[1] = 1;
  double dres = fabs(result);
  double errbnd = fmax(epsabs, epsrel * dres);
  if (abserr <= 100.0 * epmach * defabs && abserr > errbnd) ier = 2;
  if (limit == 1) ier = 1;
  bool finish_sum = false;  // label 115
  if (!(ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0)) {
    rlist2[1] = result;
    double errmax = abserr;
    int maxerr = 1;
    double area = result, errsum = abserr;
    abserr = oflow;
    int nrmax = 1, nres = 0, ktmin = 0, numrl2 = 2;
    bool extrap = false, noext = false;
    int ierro = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
    int ksgn = -1;
    if (dres >= (1.0 - 50.0 * epmach) * defabs) ksgn = 1;
    double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast = 0.0;
    double reseps = 0.0, abseps = 0.0;
    bool goto115 = false;
    for (last = 2; last <= limit; ++last) {
      const double a1 = alist[maxerr], b1 = 0.5 * (alist[maxerr] + blist[maxerr]), a2 = b1, b2 = blist[maxerr];
      erlast = errmax;
      double area1, error1, area2, error2, defab1, defab2;
      qk15i(f, a1, b1, area1, error1, resabs, defab1);
      qk15i(f, a2, b2, area2, error2, resabs, defab2);
      const double area12 = area1 + area2, erro12 = error1 + error2;
      errsum = errsum + erro12 - errmax;
      area = area + area12 - rlist[maxerr];
      if (defab1 != error1 && defab2 != error2) {
        if (fabs(rlist[maxerr] - area12) <= 1.0e-5 * fabs(area12) && erro12 >= 0.99 * errmax) {
          if (extrap) ++iroff2;
          else ++iroff1;
        }
        if (last > 10 && erro12 > errmax) ++iroff3;
      }
      rlist[maxerr] = area1;
      rlist[last] = area2;
      errbnd = fmax(epsabs, epsrel * fabs(area));
      if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
      if (iroff2 >= 5) ierro = 3;
      if (last == limit) ier = 1;
      if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * epmach) * (fabs(a2) + 1000.0 * uflow)) ier = 4;
      if (error2 > error1) {
        alist[maxerr] = a2;
        alist[last] = a1;
        blist[last] = b1;
        rlist[maxerr] = area2;
        rlist[last] = area1;
        elist[maxerr] = error2;
        elist[last] = error1;
      } else {
        alist[last] = a2;
        blist[maxerr] = b1;
        blist[last] = b2;
        elist[maxerr] = error1;
        elist[last] = error2;
      }
      qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
      if (errsum <= errbnd) {
        goto115 = true;
        break;
      }
      if (ier != 0) break;
      if (last == 2) {
        small = 0.375;
        erlarg = errsum;
        ertest = errbnd;
        rlist2[2] = area;
        continue;
      }
      if (noext) continue;
      erlarg -= erlast;
      if (fabs(b1 - a1) > small) erlarg += erro12;
      if (!extrap) {
        // test whether the interval to be bisected next is the smallest interval
        if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
        extrap = true;
        nrmax = 2;
      }
      bool skip_extrap = false;
      if (ierro != 3 && erlarg > ertest) {
        // the smallest interval has the largest error: bisect the big intervals first
        const int id = nrmax;
        int jupbnd = last;
        if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
        for (int k = id; k <= jupbnd; ++k) {
          maxerr = iord[nrmax];
          errmax = elist[maxerr];
          if (fabs(blist[maxerr] - alist[maxerr]) > small) {
            skip_extrap = true;
            break;
          }
          ++nrmax;
        }
      }
      if (skip_extrap) continue;
      // perform extrapolation
      ++numrl2;
      rlist2[numrl2] = area;
      qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
      ++ktmin;
      if (ktmin > 5 && abserr < 1.0e-3 * errsum) ier = 5;
      if (abseps < abserr) {
        ktmin = 0;
        abserr = abseps;
        result = reseps;
        correc = erlarg;
        ertest = fmax(epsabs, epsrel * fabs(reseps));
        if (abserr <= ertest) break;
      }
      // prepare bisection of the smallest interval
      if (numrl2 == 1) noext = true;
      if (ier == 5) break;
      maxerr = iord[1];
      errmax = elist[maxerr];
      nrmax = 1;
      extrap = false;
      small *= 0.5;
      erlarg = errsum;
    }
    if (last > limit) last = limit;  // loop ran to completion
    // label 100: set final result and error estimate
    bool goto130 = false;
    if (!goto115) {
      if (abserr == oflow) {
        goto115 = true;
      } else if (ier + ierro != 0) {
        if (ierro == 3) abserr += correc;
        if (ier == 0) ier = 3;
        if (result != 0.0 && area != 0.0) {
          if (abserr / fabs(result) > errsum / fabs(area)) goto115 = true;
        } else if (abserr > errsum) {
          goto115 = true;
        } else if (area == 0.0) {
          goto130 = true;
        }
      }
      if (!goto115 && !goto130) {
        // label 110: test on divergence
        if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01)) {
          if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
        }
      }
    }
    if (goto115) {
      finish_sum = true;
      // label 115: compute global integral sum
      result = 0.0;
      for (int k = 1; k <= last; ++k) result += rlist[k];
      abserr = errsum;
    }
  }
  (void)finish_sum;
  // label 130
  int neval = 30 * last - 15;
  neval *= 2;  // inf = 2
  if (ier > 2) --ier;
  R.result = result;
  R.abserr = abserr;
  R.neval = neval;
  R.ier = ier;
  R.last = last;
  return R;
}

// sigmoid(z) * dnorm(z, mean, sd)  (private$.sigmoid, R/GPCclass.R:63; the reference passes the latent VARIANCE as sd)
struct LogisticGaussian {
  double mean, sd;
  GPRC_HD double operator()(double z) const {
    const double s = 1.0 / (1.0 + exp(-z));
    const double x = (z - mean) / sd;
    return s * exp(-0.5 * x * x) / (sd * 2.5066282746310002);  // sqrt(2 pi)
  }
};

GPRC_HD inline QuadResult logistic_gaussian(double mean, double sd) {
  const double tol = 1.220703125e-4;  // .Machine$double.eps^0.25 = 2^-13
  LogisticGaussian f{mean, sd};
  return qagi(f, tol, tol);
}

}  // namespace gprc_quad
