/* shim_exec.c -- EXECUTES src/gprc_shim.c (the `.Call` shim of the drop-in R package) against libgprc on a GPU, with the
 * miniature R runtime of tests/stubs/mini_r.c standing in for R (not installed here).  Entry points are resolved by name
 * and argument count from the table R_init_gprc registers, as `.Call(C_gprc_*, ...)` resolves them, and are fed the SEXPs
 * the R host code (R/GPRclass.R, R/GPCclass.R, R/fit.R of the package) builds:
 *   - the four known answers of the reference's tests/testthat/test-gpr.R:5-28 through C_gprc_gpr_fit + C_gprc_gpr_predict
 *     (one attempt of the loop R/GPRclass.R:141-148, then $predict), tolerance 1.5e-8 as expect_equivalent;
 *   - X_star with the wrong number of rows: "non-conformable arrays" raised from shim level, handle intact afterwards;
 *   - predict(pointwise_var = FALSE), $alpha / $L downloads, dens / dens_deriv / fit_family, covariance_matrix;
 *   - the first case of tests/testthat/test-gpc.R:5-10 (constructor argument order fixed) through C_gprc_gpc_fit +
 *     C_gprc_gpc_predict_class: p(-0.2) < 0.5 < p(0.2), and the values the restatement gives;
 *   - a pending interrupt during a long predict: the library returns, the shim raises from its own frame;
 *   - finalizers (R's garbage collector at exit) free every handle; R_unload_gprc frees the context.
 *   gcc -std=c11 -Itests/stubs -Iinclude src/gprc_shim.c tests/stubs/mini_r.c tests/shim_exec.c -L<pkg> -lgprc -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mini_r.h"
#include "gprc.h"

void R_init_gprc(DllInfo*);
void R_unload_gprc(DllInfo*);

typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static int failures = 0;
static void expect(int cond, const char* what) {
  printf("%-78s %s\n", what, cond ? "ok" : "FAIL");
  if (!cond) ++failures;
}
static int near(double a, double b, double tol) { return fabs(a - b) <= tol; }

/* the `gprc_kernel` attribute cov_func() attaches: list(id =, <named parameters>) */
static SEXP spec(int id, const char* p1, double v1, const char* p2, double v2) {
  const char* names[3] = {"id", p1, p2};
  SEXP vals[3] = {Rf_ScalarInteger(id), Rf_ScalarReal(v1), p2 ? Rf_ScalarReal(v2) : R_NilValue};
  return mini_r_named_list(p2 ? 3 : 2, names, vals);
}
static void* must(const char* name, int nargs) {
  void* f = (void*)mini_r_lookup(name, nargs);
  if (!f) {
    printf("routine %s with %d arguments is not registered\n", name, nargs);
    exit(2);
  }
  return f;
}

static void known_answer(F5 fit, F4 predict, SEXP k, double x0, double x1, double y0, double y1, double noise, double xs,
                         double mean, double var, const char* what) {
  const double X[2] = {x0, x1}, y[2] = {y0, y1};
  SEXP r = fit(k, mini_r_matrix(1, 2, X), mini_r_vector(2, y), Rf_ScalarReal(noise), R_NilValue);
  const double info = Rf_asReal(VECTOR_ELT(r, 2));
  SEXP p = predict(VECTOR_ELT(r, 0), mini_r_matrix(1, 1, &xs), R_NilValue, R_NilValue);
  char buf[160];
  snprintf(buf, sizeof buf, "%s: info 0, m x 2 result, mean %.12f var %.12f", what, REAL(p)[0], REAL(p)[1]);
  expect(info == 0.0 && Rf_nrows(p) == 1 && Rf_ncols(p) == 2 && near(REAL(p)[0], mean, 1.5e-8) && near(REAL(p)[1], var, 1.5e-8), buf);
}

int main(void) {
  mini_r_init();
  R_init_gprc(NULL);
  F5 fit = (F5)must("C_gprc_gpr_fit", 5);
  F4 predict = (F4)must("C_gprc_gpr_predict", 4);
  F2 predict_cov = (F2)must("C_gprc_gpr_predict_cov", 2);
  F2 gpr_get = (F2)must("C_gprc_gpr_get", 2);
  F3 cov_matrix = (F3)must("C_gprc_cov_matrix", 3);
  F4 logml = (F4)must("C_gprc_logml", 4);
  F6 logml_grad = (F6)must("C_gprc_logml_grad", 6);
  F5 fit_family = (F5)must("C_gprc_fit_family", 5);
  F2 set_option = (F2)must("C_gprc_set_option", 2);
  F6 gpc_fit = (F6)must("C_gprc_gpc_fit", 6);
  F4 gpc_latent = (F4)must("C_gprc_gpc_predict_latent", 4);
  F2 gpc_class = (F2)must("C_gprc_gpc_predict_class", 2);
  F2 gpc_get = (F2)must("C_gprc_gpc_get", 2);
  expect(mini_r_lookup("C_gprc_gpr_fit", 4) == NULL, "a call with the wrong number of arguments is not resolved");

  /* ---- tests/testthat/test-gpr.R:5-28 ---- */
  const double e1 = exp(-1.0), e2 = exp(-2.0), e3 = exp(-3.0), e4 = exp(-4.0);
  known_answer(fit, predict, spec(GPRC_POLYNOMIAL, "sigma", 0.25, "p", 1.0), -0.5, 0.5, 4, 4, 0.5, 0.0, 2.0, 0.125,
               "test-gpr.R:6-9   polynomial");
  known_answer(fit, predict, spec(GPRC_CONSTANT, "c", 1.0, NULL, 0), 1, 2, 1, 3, 1.0, 3.0, 4.0 / 3, 1.0 / 3,
               "test-gpr.R:12-15 constant");
  known_answer(fit, predict, spec(GPRC_CONSTANT, "c", 1.0, NULL, 0), 100, 54, 5, 0, 1.0, M_PI, 5.0 / 3, 1.0 / 3,
               "test-gpr.R:16-19 constant");
  known_answer(fit, predict, spec(GPRC_SQREXP, "l", 1.0, NULL, 0), 1, 2, 0, 1, 1.0, 0.0, (2 * e2 - e1) / (4 - e1),
               1 - (2 * e1 - 2 * e3 + 2 * e4) / (4 - e1), "test-gpr.R:23-27 sqrexp");
  expect(mini_r_protect_depth() == 0, "PROTECT / UNPROTECT balanced after the GPR calls");

  /* ---- a model with D = 2: shapes, bindings, error path ---- */
  enum { N = 40, D = 2, M = 7 };
  double X[D * N], y[N], Xs[D * M];
  unsigned s = 12345u;
  for (int i = 0; i < D * N; ++i) X[i] = ((s = s * 1664525u + 1013904223u) >> 8) / 8388608.0 - 1.0;
  for (int i = 0; i < N; ++i) y[i] = sin(2 * X[2 * i]) + 0.5 * X[2 * i + 1];
  for (int i = 0; i < D * M; ++i) Xs[i] = ((s = s * 1664525u + 1013904223u) >> 8) / 8388608.0 - 1.0;
  SEXP k2 = spec(GPRC_RATQUAD, "l", 0.8, "alpha", 1.5);
  SEXP Xm = mini_r_matrix(D, N, X), yv = mini_r_vector(N, y), noise = Rf_ScalarReal(0.05);
  SEXP r = fit(k2, Xm, yv, noise, R_NilValue);
  SEXP ptr = VECTOR_ELT(r, 0);
  const double logp = Rf_asReal(VECTOR_ELT(r, 1));
  SEXP lm = logml(k2, Xm, yv, noise);
  expect(Rf_asReal(VECTOR_ELT(r, 2)) == 0.0 && near(REAL(lm)[0], logp, 1e-10 * fabs(logp)) && REAL(lm)[2] == 0.0,
         "dens(v) (C_gprc_logml) equals the logp of GPR$new on the same data");
  SEXP pw = predict(ptr, mini_r_matrix(D, M, Xs), R_NilValue, R_NilValue);
  SEXP pc = predict_cov(ptr, mini_r_matrix(D, M, Xs));
  int same = Rf_nrows(pw) == M && Rf_ncols(pw) == 2 && XLENGTH(pc) == 2 && Rf_nrows(VECTOR_ELT(pc, 1)) == M &&
             Rf_ncols(VECTOR_ELT(pc, 1)) == M && Rf_ncols(VECTOR_ELT(pc, 0)) == 1;
  for (int i = 0; same && i < M; ++i)
    same = near(REAL(VECTOR_ELT(pc, 0))[i], REAL(pw)[i], 1e-12) && near(REAL(VECTOR_ELT(pc, 1))[i + i * M], REAL(pw)[M + i], 1e-10);
  expect(same, "predict(pointwise_var = FALSE): list(m x 1, m x m), diagonal = pointwise variances");
  SEXP alpha = gpr_get(ptr, Rf_ScalarInteger(GPRC_GET_ALPHA)), L = gpr_get(ptr, Rf_ScalarInteger(GPRC_GET_L));
  SEXP K = cov_matrix(k2, Xm, Xm);
  /* (K + noise I) alpha = y, and L L' = K + noise I on a few entries */
  double worst = 0.0;
  for (int i = 0; i < N; ++i) {
    double acc = 0.05 * REAL(alpha)[i];
    for (int j = 0; j < N; ++j) acc += REAL(K)[i + j * N] * REAL(alpha)[j];
    worst = fmax(worst, fabs(acc - y[i]));
  }
  double llt = 0.0;
  for (int j = 0; j <= 5; ++j) llt += REAL(L)[5 + j * N] * REAL(L)[3 + j * N] * (j <= 3);
  expect(XLENGTH(alpha) == N && Rf_nrows(L) == N && Rf_ncols(L) == N && worst < 1e-9 && near(llt, REAL(K)[5 + 3 * N], 1e-12) &&
             REAL(L)[3 + 5 * N] == 0.0,
         "$alpha, $L, covariance_matrix: (K + noise I) alpha = y, L L' = K + noise I, zeros above the diagonal");
  /* X_star with the wrong number of rows: the shim must raise before the library reads D x m doubles */
  {
    jmp_buf jb;
    mini_r_handler = &jb;
    int raised = 0;
    if (setjmp(jb) == 0) predict(ptr, mini_r_matrix(1, 2 * M, Xs), R_NilValue, R_NilValue);
    else raised = 1;
    mini_r_handler = NULL;
    expect(raised && strstr(mini_r_last_error, "non-conformable") != NULL, "X_star with nrow != nrow(X): \"non-conformable arrays\"");
    SEXP again = predict(ptr, mini_r_matrix(D, M, Xs), R_NilValue, R_NilValue);
    expect(near(REAL(again)[0], REAL(pw)[0], 0.0), "... and the model handle still works afterwards");
  }
  /* dens_deriv, textbook formula against central differences of dens */
  {
    SEXP g = logml_grad(k2, Xm, yv, noise, Rf_ScalarInteger(GPRC_GRAD_TEXTBOOK), Rf_ScalarInteger(2));
    const double h = 1e-5;
    const double up = REAL(logml(spec(GPRC_RATQUAD, "l", 0.8 + h, "alpha", 1.5), Xm, yv, noise))[0];
    const double dn = REAL(logml(spec(GPRC_RATQUAD, "l", 0.8 - h, "alpha", 1.5), Xm, yv, noise))[0];
    expect(XLENGTH(g) == 2 && near(REAL(g)[0], (up - dn) / (2 * h), 1e-5 * fmax(1.0, fabs(REAL(g)[0]))),
           "dens_deriv (C_gprc_logml_grad, textbook) = d dens / d l by central differences");
    SEXP ff = fit_family(Rf_ScalarInteger(GPRC_SQREXP), Xm, yv, noise, Rf_ScalarInteger(0));
    const double at = REAL(logml(spec(GPRC_SQREXP, "l", REAL(ff)[1], NULL, 0), Xm, yv, noise))[0];
    expect(XLENGTH(ff) == 2 && REAL(ff)[1] > 0 && REAL(ff)[1] < 10 && near(REAL(ff)[0], at, 1e-9 * fabs(at)),
           "fit_family(sqrexp): c(value, l) with value = dens(l) (Brent on [0, 10], R/fit.R:143)");
  }

  /* ---- tests/testthat/test-gpc.R:5-10: exp(-3 (x - y)^2) = sqrexp with l = 1 / sqrt(6) ---- */
  {
    double Xc[21], yc[21];
    for (int i = 0; i < 21; ++i) {
      Xc[i] = -1.0 + 0.1 * i;
      yc[i] = (i > 10) ? 1.0 : -1.0;
    }
    SEXP kc = spec(GPRC_SQREXP, "l", 1.0 / sqrt(6.0), NULL, 0);
    SEXP gr = gpc_fit(kc, mini_r_matrix(1, 21, Xc), mini_r_vector(21, yc), Rf_ScalarReal(1e-5), Rf_ScalarInteger(1), R_NilValue);
    const int iters = Rf_asInteger(VECTOR_ELT(gr, 1)), status = Rf_asInteger(VECTOR_ELT(gr, 5));
    SEXP trace = VECTOR_ELT(gr, 2);
    const double q[2] = {-0.2, 0.2};
    SEXP pr = gpc_class(VECTOR_ELT(gr, 0), mini_r_matrix(1, 2, q));
    SEXP lat = gpc_latent(VECTOR_ELT(gr, 0), mini_r_matrix(1, 2, q), R_NilValue, R_NilValue);
    SEXP fh = gpc_get(VECTOR_ELT(gr, 0), Rf_ScalarInteger(GPRC_GET_FHAT));
    const double logq = REAL(trace)[XLENGTH(trace) - 1] - Rf_asReal(VECTOR_ELT(gr, 3)); /* objective - sum(diag(L)), R/GPCclass.R:103 */
    char buf[200];
    snprintf(buf, sizeof buf, "test-gpc.R:5-10: %d iterations, p(-0.2) = %.4f < 0.5 < p(0.2) = %.4f, logq %.6f", iters,
             REAL(pr)[0], REAL(pr)[1], logq);
    expect(status == 0 && iters == 4 && XLENGTH(trace) == 4 && REAL(pr)[0] < 0.5 && REAL(pr)[1] > 0.5 &&
               near(REAL(pr)[0], 0.2798, 2e-4) && near(REAL(pr)[1], 0.6467, 2e-4) && near(logq, -30.942983985, 1e-6) &&
               near(REAL(lat)[0], -0.987961, 1e-5) && near(REAL(lat)[2], 0.447528, 1e-5) && XLENGTH(fh) == 21,
           buf);
  }

  /* ---- a pending Ctrl-C during a long predict: the library polls between chunks and returns; the shim raises ---- */
  {
    enum { NB_ = 300 };
    static double Xb[NB_], yb[NB_];
    const long mbig = (1L << 20) + 5000; /* two chunks */
    double* Xsb = (double*)malloc(sizeof(double) * (size_t)mbig);
    for (int i = 0; i < NB_; ++i) {
      Xb[i] = -6.0 + 12.0 * i / NB_;
      yb[i] = 0.1 * Xb[i] * Xb[i] * Xb[i];
    }
    for (long i = 0; i < mbig; ++i) Xsb[i] = -6.0 + 12.0 * (double)i / (double)mbig;
    SEXP rb = fit(spec(GPRC_SQREXP, "l", 1.0, NULL, 0), mini_r_matrix(1, NB_, Xb), mini_r_vector(NB_, yb), Rf_ScalarReal(0.01), R_NilValue);
    SEXP xs = Rf_allocMatrix(REALSXP, 1, (int)mbig);
    memcpy(REAL(xs), Xsb, sizeof(double) * (size_t)mbig);
    jmp_buf jb;
    mini_r_handler = &jb;
    int raised = 0;
    mini_r_set_interrupt(1);
    if (setjmp(jb) == 0) predict(VECTOR_ELT(rb, 0), xs, R_NilValue, R_NilValue);
    else raised = 1;
    mini_r_handler = NULL;
    mini_r_set_interrupt(0);
    expect(raised && strstr(mini_r_last_error, "interrupt") != NULL, "pending interrupt: the long predict stops between chunks, error raised from the shim");
    SEXP ok = predict(VECTOR_ELT(rb, 0), mini_r_matrix(1, 3, Xsb), R_NilValue, R_NilValue);
    expect(isfinite(REAL(ok)[0]) && isfinite(REAL(ok)[3]), "... and model and context are usable afterwards");
    free(Xsb);
  }
  set_option(Rf_ScalarInteger(GPRC_OPT_PREDICT_PATH), Rf_ScalarInteger(0));

  /* ---- R's garbage collector at exit: finalizers free every handle, then the package unloads ---- */
  const int freed = mini_r_gc();
  expect(freed >= 7 && mini_r_gc() == 0, "finalizers ran once for every model handle");
  R_unload_gprc(NULL);
  printf(failures ? "%d FAILURES\n" : "ALL OK (%d)\n", failures);
  return failures ? 1 : 0;
}
