/* Minimal stand-in for R's C API, ONLY to type-check rshim/lqg_rcall.c in an image without R
 * (tests/test_host.py).  Declarations follow the documented signatures in "Writing R Extensions";
 * nothing here is ever linked or run. */
#ifndef LQG_TEST_R_STUB_H
#define LQG_TEST_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef unsigned int SEXPTYPE;
typedef ptrdiff_t R_xlen_t;
#define REALSXP 14
#define VECSXP 19
#define INTSXP 13
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_length(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_isMatrix(SEXP);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_ScalarLogical(int);
SEXP Rf_ScalarInteger(int);
int Rf_isNull(SEXP);
extern SEXP R_NilValue;
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
#endif
