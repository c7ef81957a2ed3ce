/* Plain-C client of libgprc: the same calls src/gprc_shim.c makes, without R or Python in the process.
 * Reproduces the four known answers of the reference's tests/testthat/test-gpr.R through the C ABI.
 *   gcc -std=c11 -Iinclude tests/c_abi_smoke.c -L<pkg> -lgprc -lm -o c_abi_smoke && LD_LIBRARY_PATH=<pkg> ./c_abi_smoke */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "gprc.h"

static int check(const char* what, double got, double want) {
  const int ok = fabs(got - want) <= 1.5e-8; /* expect_equivalent's tolerance */
  printf("%-28s got % .15f want % .15f %s\n", what, got, want, ok ? "ok" : "FAIL");
  return ok ? 0 : 1;
}

static int one(gprc_ctx* ctx, gprc_kernel k, const double* X, const double* y, double noise, double xs, double mean,
               double var, const char* name) {
  gprc_gpr* g = NULL;
  double logp = 0, m = 0, v = 0;
  long info = 0;
  if (gprc_gpr_fit(ctx, &k, X, 1, 2, y, noise, &g, &logp, &info) != 0 || info != 0 || !g) {
    printf("%s: fit failed: %s (info %ld)\n", name, gprc_last_error(), info);
    return 1;
  }
  if (gprc_gpr_predict(g, &xs, 1, &m, &v) != 0) {
    printf("%s: predict failed: %s\n", name, gprc_last_error());
    return 1;
  }
  gprc_gpr_free(g);
  char buf[64];
  int bad = 0;
  snprintf(buf, sizeof buf, "%s mean", name);
  bad += check(buf, m, mean);
  snprintf(buf, sizeof buf, "%s var", name);
  bad += check(buf, v, var);
  return bad;
}

int main(void) {
  gprc_ctx* ctx = NULL;
  if (gprc_ctx_create(&ctx, 0) != 0) {
    printf("no context: %s\n", gprc_last_error());
    return 2;
  }
  int bad = 0;
  gprc_kernel k;
  memset(&k, 0, sizeof k);
  { /* test-gpr.R:6-9 */
    const double X[2] = {-0.5, 0.5}, y[2] = {4, 4};
    k.id = GPRC_POLYNOMIAL; k.sigma = 0.25; k.p = 1;
    bad += one(ctx, k, X, y, 0.5, 0.0, 2.0, 0.125, "polynomial");
  }
  { /* :12-15 */
    const double X[2] = {1, 2}, y[2] = {1, 3};
    k.id = GPRC_CONSTANT; k.c = 1;
    bad += one(ctx, k, X, y, 1.0, 3.0, 4.0 / 3, 1.0 / 3, "constant (1)");
  }
  { /* :16-19 */
    const double X[2] = {100, 54}, y[2] = {5, 0};
    bad += one(ctx, k, X, y, 1.0, 3.14159265358979323846, 5.0 / 3, 1.0 / 3, "constant (2)");
  }
  { /* :23-27 */
    const double X[2] = {1, 2}, y[2] = {0, 1};
    k.id = GPRC_SQREXP; k.l = 1;
    const double e1 = exp(-1.0), e2 = exp(-2.0), e3 = exp(-3.0), e4 = exp(-4.0);
    bad += one(ctx, k, X, y, 1.0, 0.0, (2 * e2 - e1) / (4 - e1), 1 - (2 * e1 - 2 * e3 + 2 * e4) / (4 - e1), "sqrexp");
  }
  gprc_ctx_free(ctx);
  printf(bad ? "FAILED (%d)\n" : "all known answers reproduced through the C ABI\n", bad);
  return bad ? 1 : 0;
}
