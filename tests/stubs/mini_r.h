/* driver-side interface of the miniature R runtime (tests/stubs/mini_r.c) */
#ifndef MINI_R_H
#define MINI_R_H
#include <setjmp.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"
extern jmp_buf* mini_r_handler;
extern char mini_r_last_error[512];
void mini_r_init(void);
int mini_r_gc(void);
int mini_r_protect_depth(void);
void mini_r_set_interrupt(int pending);
DL_FUNC mini_r_lookup(const char* name, int nargs);
SEXP mini_r_matrix(int nr, int nc, const double* v);
SEXP mini_r_vector(int n, const double* v);
SEXP mini_r_named_list(int n, const char** names, SEXP* values);
int* INTEGER(SEXP);
#endif
