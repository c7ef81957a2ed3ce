/* mini_r.c -- a miniature implementation of the part of R's C API that src/gprc_shim.c uses, so that the shim's `.Call`
 * entry points can be EXECUTED (tests/shim_exec.c) where R itself is not installed.  Test infrastructure only.
 *
 * What it models: SEXPs as tagged heap records (REALSXP / INTSXP / LGLSXP / STRSXP / VECSXP / CHARSXP / EXTPTRSXP / NILSXP)
 * with dim and names attributes; PROTECT as a counter; Rf_error as a longjmp to the nearest mini_r_try() -- like R's
 * error longjmp, it unwinds through the shim's frames, which is exactly what the shim's "release first, then raise"
 * discipline must survive; external pointers with C finalizers that mini_r_gc() runs (onexit = TRUE: at shutdown);
 * R_ToplevelExec / R_CheckUserInterrupt with a test-controlled pending-interrupt flag; the routine table handed to
 * R_registerRoutines, so that the driver looks entry points up by NAME and argument count as `.Call` does. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "Rinternals.h"
#include "R_ext/Rdynload.h"
#include "mini_r.h"

struct SEXPREC {
  int type;
  R_xlen_t len;
  void* data;          /* double[] / int[] / SEXP[] / char[] / external address */
  int nrow, ncol;      /* dim attribute (ncol < 0: none) */
  SEXP names;          /* names attribute (STRSXP) or NULL */
  R_CFinalizer_t fin;  /* EXTPTRSXP */
  struct SEXPREC* next_ext;
};

static struct SEXPREC nil_rec = {0, 0, NULL, 0, -1, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec;
static struct SEXPREC names_sym = {1, 0, (void*)"names", 0, -1, NULL, NULL, NULL};
SEXP R_NamesSymbol = &names_sym;
double R_NaReal;
static SEXP ext_list = NULL;
static int protect_depth = 0;
static const R_CallMethodDef* routines = NULL;
static int interrupt_flag = 0;
static int option_device = -1;

jmp_buf* mini_r_handler = NULL;
char mini_r_last_error[512];

static SEXP mk(int type, R_xlen_t len, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof *s);
  s->type = type;
  s->len = len;
  s->ncol = -1;
  s->data = len > 0 ? calloc((size_t)len, elt) : NULL;
  return s;
}

void mini_r_init(void) {
  union { unsigned long long u; double d; } na = {0x7FF00000000007A2ull}; /* R's NA_real_ payload 1954 */
  R_NaReal = na.d;
}
SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  switch (type) {
    case REALSXP: return mk(REALSXP, n, sizeof(double));
    case INTSXP: case LGLSXP: return mk((int)type, n, sizeof(int));
    case VECSXP: case STRSXP: {
      SEXP s = mk((int)type, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
      return s;
    }
  }
  Rf_error("mini_r: allocVector of type %u", type);
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
  s->nrow = nr;
  s->ncol = nc;
  return s;
}
double* REAL(SEXP s) {
  if (s->type != REALSXP) Rf_error("mini_r: REAL() on a non-double");
  return (double*)s->data;
}
int* INTEGER(SEXP s) { return (int*)s->data; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int Rf_length(SEXP s) { return (int)s->len; }
int Rf_nrows(SEXP s) { return s->ncol >= 0 ? s->nrow : (int)s->len; }
int Rf_ncols(SEXP s) { return s->ncol >= 0 ? s->ncol : 1; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP*)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { return ((SEXP*)s->data)[i] = v; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP*)s->data)[i]; }
const char* CHAR(SEXP s) { return (const char*)s->data; }
SEXP Rf_getAttrib(SEXP s, SEXP what) { return (what == R_NamesSymbol && s->names) ? s->names : R_NilValue; }
SEXP Rf_install(const char* name) {
  SEXP s = mk(1, (R_xlen_t)strlen(name) + 1, 1);
  memcpy(s->data, name, strlen(name) + 1);
  return s;
}
SEXP Rf_GetOption1(SEXP sym) {
  if (!strcmp((const char*)sym->data, "gprc.device") && option_device >= 0) return Rf_ScalarInteger(option_device);
  return R_NilValue;
}
static double as_real(SEXP s) {
  if (s->len < 1) return R_NaReal;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) return (double)((int*)s->data)[0];
  Rf_error("mini_r: cannot coerce to a number");
}
double Rf_asReal(SEXP s) { return as_real(s); }
int Rf_asInteger(SEXP s) { return (int)as_real(s); }
int Rf_asLogical(SEXP s) { return as_real(s) != 0.0; }
SEXP Rf_ScalarReal(double v) {
  SEXP s = Rf_allocVector(REALSXP, 1);
  REAL(s)[0] = v;
  return s;
}
SEXP Rf_ScalarInteger(int v) {
  SEXP s = Rf_allocVector(INTSXP, 1);
  INTEGER(s)[0] = v;
  return s;
}
SEXP Rf_protect(SEXP s) {
  ++protect_depth;
  return s;
}
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) {
    fprintf(stderr, "mini_r: PROTECT stack underflow\n");
    abort();
  }
}
int mini_r_protect_depth(void) { return protect_depth; }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(mini_r_last_error, sizeof mini_r_last_error, fmt, ap);
  va_end(ap);
  if (!mini_r_handler) {
    fprintf(stderr, "mini_r: uncaught error: %s\n", mini_r_last_error);
    abort();
  }
  protect_depth = 0; /* R unwinds the protect stack to the context it jumps to */
  longjmp(*mini_r_handler, 1);
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag;
  (void)prot;
  SEXP s = mk(22 /* EXTPTRSXP */, 0, 1);
  s->data = p;
  s->next_ext = ext_list;
  ext_list = s;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fin, Rboolean onexit) {
  (void)onexit;
  s->fin = fin;
}
int mini_r_gc(void) { /* run every pending finalizer, as R does at collection / exit; returns how many ran */
  int n = 0;
  for (SEXP s = ext_list; s; s = s->next_ext)
    if (s->fin && s->data) {
      s->fin(s);
      ++n;
    }
  return n;
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
void mini_r_set_interrupt(int pending) { interrupt_flag = pending; }
void R_CheckUserInterrupt(void) {
  if (interrupt_flag) {
    interrupt_flag = 0;
    Rf_error("interrupt");
  }
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) { /* FALSE if `fun` jumped */
  jmp_buf here, *saved = mini_r_handler;
  const int depth = protect_depth;
  Rboolean ok = TRUE;
  mini_r_handler = &here;
  if (setjmp(here) == 0) fun(data);
  else ok = FALSE;
  mini_r_handler = saved;
  protect_depth = depth;
  return ok;
}
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)dll; (void)c; (void)f; (void)e;
  routines = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) {
  (void)dll; (void)v;
  return 1;
}
DL_FUNC mini_r_lookup(const char* name, int nargs) { /* what .Call(C_name, ...) resolves through */
  for (const R_CallMethodDef* r = routines; r && r->name; ++r)
    if (!strcmp(r->name, name)) return r->numArgs == nargs ? r->fun : NULL;
  return NULL;
}
/* ---- constructors for the driver ---- */
SEXP mini_r_matrix(int nr, int nc, const double* v) {
  SEXP s = Rf_allocMatrix(REALSXP, nr, nc);
  memcpy(REAL(s), v, sizeof(double) * (size_t)nr * nc);
  return s;
}
SEXP mini_r_vector(int n, const double* v) {
  SEXP s = Rf_allocVector(REALSXP, n);
  memcpy(REAL(s), v, sizeof(double) * (size_t)n);
  return s;
}
SEXP mini_r_named_list(int n, const char** names, SEXP* values) {
  SEXP s = Rf_allocVector(VECSXP, n), nm = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; ++i) {
    SET_VECTOR_ELT(s, i, values[i]);
    ((SEXP*)nm->data)[i] = Rf_install(names[i]);
  }
  s->names = nm;
  return s;
}
