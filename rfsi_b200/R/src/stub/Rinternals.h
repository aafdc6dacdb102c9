/* Minimal stand-in for R's C API: just enough declarations to syntax-check r_shim.c in an
 * image without R (tests/test_r_shim.py).  NOT used on a machine that has R. */
#ifndef RFSI_R_STUB_H
#define RFSI_R_STUB_H
#include <stddef.h>
#include <string.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
enum { INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };
extern SEXP R_NilValue, R_NamesSymbol, R_ClassSymbol, R_RowNamesSymbol;
double* REAL(SEXP); int* INTEGER(SEXP); R_xlen_t XLENGTH(SEXP);
int Rf_nrows(SEXP); int Rf_ncols(SEXP); int Rf_isReal(SEXP); int Rf_asInteger(SEXP); int Rf_asLogical(SEXP);
SEXP Rf_allocMatrix(int, int, int); SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); SEXP VECTOR_ELT(SEXP, R_xlen_t); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_mkChar(const char*); SEXP Rf_mkString(const char*); SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_install(const char*);
void Rf_error(const char*, ...); void Rf_warning(const char*, ...);
char* R_alloc(size_t, int);
int TYPEOF(SEXP);
extern int R_NaInt; extern double R_NaN;
#define NA_INTEGER R_NaInt
SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
