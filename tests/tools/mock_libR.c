/*
 * mock_libR.c -- TEST infrastructure: a working miniature of the part of R's C API that
 * rfsi_b200/R/src/r_shim.c uses, so that the .Call entries (C_rfsi_near_obs, C_rfsi_predict,
 * C_rfsi_pred_fused, C_rfsi_pred_maps ...) run END TO END -- on the GPU, against the oracle --
 * in an image that has no R (tests/test_gpu_r_shim.py, tests/test_r_shim.py).
 *
 * An SEXP is a tagged struct (type, length, dim, names, payload).  PROTECT/UNPROTECT keep a
 * counter that the tests check for balance.  Rf_error records the message and longjmps to the
 * trampoline mock_call() every test goes through, exactly as R unwinds to its top level: the
 * shim's rule "free native temporaries BEFORE Rf_error" is therefore exercised for real.
 * R_alloc memory is released when the trampoline returns (R frees it when .Call returns).
 * Not a re-implementation of R: no garbage collector (objects live until mock_reset()), no
 * evaluation, no ALTREP.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"

struct SEXPREC {
    int type;            /* INTSXP, REALSXP, STRSXP, VECSXP, or the private tags below */
    R_xlen_t length;
    int has_dim, nrow, ncol;   /* dim attribute */
    void* data;
    SEXP names, klass, rownames;
    void* extptr;
    R_CFinalizer_t finalizer;
    struct SEXPREC* next;
};
enum { NILSXP_ = 0, SYMSXP_ = 1, CHARSXP_ = 9, EXTPTRSXP_ = 22 };

static struct SEXPREC nil_obj = {NILSXP_, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC names_sym = {SYMSXP_, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC class_sym = {SYMSXP_, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC rownames_sym = {SYMSXP_, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_NamesSymbol = &names_sym;
SEXP R_ClassSymbol = &class_sym;
SEXP R_RowNamesSymbol = &rownames_sym;
int R_NaInt = INT32_MIN;
double R_NaN;

static SEXP all_objects = NULL;
static int protect_depth = 0, protect_max = 0;
static jmp_buf* top = NULL;
static char last_error[1024], last_warning[1024];
static int nwarnings = 0;
static void* ralloc_list[4096];
static int ralloc_n = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t elt) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type; s->length = n; s->names = s->klass = s->rownames = R_NilValue;
    s->data = (elt && n > 0) ? calloc((size_t)n, elt) : NULL;
    s->next = all_objects; all_objects = s;
    return s;
}

/* ---- the API the shim uses ------------------------------------------------------------ */
double* REAL(SEXP s) { if (s->type != REALSXP) Rf_error("REAL() can only be applied to a 'numeric', not a type-%d object", s->type); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP) Rf_error("INTEGER() can only be applied to a 'integer', not a type-%d object", s->type); return (int*)s->data; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
int TYPEOF(SEXP s) { return s->type; }
int Rf_isReal(SEXP s) { return s->type == REALSXP; }
int Rf_nrows(SEXP s) { return s->has_dim ? s->nrow : (int)s->length; }
int Rf_ncols(SEXP s) { return s->has_dim ? s->ncol : 1; }
int Rf_asInteger(SEXP s) {
    if (s->length < 1) return R_NaInt;
    if (s->type == INTSXP) return ((int*)s->data)[0];
    if (s->type == REALSXP) return (int)((double*)s->data)[0];
    return R_NaInt;
}
int Rf_asLogical(SEXP s) { return Rf_asInteger(s) != 0; }
SEXP Rf_allocVector(int type, R_xlen_t n) {
    switch (type) {
        case REALSXP: return new_obj(type, n, sizeof(double));
        case INTSXP: return new_obj(type, n, sizeof(int));
        case STRSXP: case VECSXP: {
            SEXP s = new_obj(type, n, sizeof(SEXP));
            for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
            return s;
        }
        default: Rf_error("mock allocVector: type %d", type); return R_NilValue;
    }
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->has_dim = 1; s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_protect(SEXP s) { if (++protect_depth > protect_max) protect_max = protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
SEXP SET_VECTOR_ELT(SEXP v, R_xlen_t i, SEXP x) {
    if (v->type != VECSXP || i < 0 || i >= v->length) Rf_error("SET_VECTOR_ELT: bad list or index");
    ((SEXP*)v->data)[i] = x; return x;
}
SEXP VECTOR_ELT(SEXP v, R_xlen_t i) {
    if (v->type != VECSXP || i < 0 || i >= v->length) Rf_error("VECTOR_ELT: bad list or index");
    return ((SEXP*)v->data)[i];
}
void SET_STRING_ELT(SEXP v, R_xlen_t i, SEXP x) {
    if (v->type != STRSXP || i < 0 || i >= v->length) Rf_error("SET_STRING_ELT: bad vector or index");
    ((SEXP*)v->data)[i] = x;
}
SEXP Rf_mkChar(const char* c) { SEXP s = new_obj(CHARSXP_, (R_xlen_t)strlen(c), 0); s->data = strdup(c); return s; }
SEXP Rf_install(const char* c) { SEXP s = Rf_mkChar(c); s->type = SYMSXP_; return s; }
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP v) {
    if (what == R_NamesSymbol) x->names = v;
    else if (what == R_ClassSymbol) x->klass = v;
    else if (what == R_RowNamesSymbol) x->rownames = v;
    return v;
}
SEXP Rf_mkString(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); ((SEXP*)s->data)[0] = Rf_mkChar(c); return s; }
char* R_alloc(size_t n, int size) {
    void* p = calloc(n ? n : 1, (size_t)size);
    if (ralloc_n < 4096) ralloc_list[ralloc_n++] = p;
    return (char*)p;
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = new_obj(EXTPTRSXP_, 0, 0); s->extptr = p; return s; }
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP_ ? s->extptr : NULL; }
void R_ClearExternalPtr(SEXP s) { s->extptr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->finalizer = f; }

void Rf_warning(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(last_warning, sizeof last_warning, fmt, ap); va_end(ap);
    ++nwarnings;
}
void Rf_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(last_error, sizeof last_error, fmt, ap); va_end(ap);
    if (top) longjmp(*top, 1);
    fprintf(stderr, "mock libR: Rf_error outside mock_call: %s\n", last_error);
    abort();
}

/* ---- what the tests use --------------------------------------------------------------- */
typedef SEXP (*call12_t)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* .Call(fn, args...): NULL when the callee raised an R error (message in mock_last_error()) */
SEXP mock_call(void* fn, int nargs, SEXP* args) {
    SEXP a[12];
    for (int i = 0; i < 12; ++i) a[i] = i < nargs ? args[i] : R_NilValue;
    jmp_buf here;
    jmp_buf* saved = top;
    const int depth = protect_depth;
    last_error[0] = 0;
    SEXP out = NULL;
    top = &here;
    if (setjmp(here) == 0) out = ((call12_t)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]);
    else protect_depth = depth;            /* R resets the protect stack when it unwinds */
    top = saved;
    for (int i = 0; i < ralloc_n; ++i) free(ralloc_list[i]);
    ralloc_n = 0;
    return out;
}
const char* mock_last_error(void) { return last_error; }
const char* mock_last_warning(void) { return last_warning; }
int mock_warning_count(void) { return nwarnings; }
int mock_protect_depth(void) { return protect_depth; }
SEXP mock_nil(void) { return R_NilValue; }

SEXP mock_real(const double* x, R_xlen_t n, int nrow, int ncol) {
    SEXP s = ncol ? Rf_allocMatrix(REALSXP, nrow, ncol) : Rf_allocVector(REALSXP, n);
    if (n > 0) memcpy(s->data, x, (size_t)n * sizeof(double));
    return s;
}
SEXP mock_int(const int* x, R_xlen_t n, int nrow, int ncol) {
    SEXP s = ncol ? Rf_allocMatrix(INTSXP, nrow, ncol) : Rf_allocVector(INTSXP, n);
    if (n > 0) memcpy(s->data, x, (size_t)n * sizeof(int));
    return s;
}
SEXP mock_list(SEXP* elts, int n) {
    SEXP s = Rf_allocVector(VECSXP, n);
    for (int i = 0; i < n; ++i) ((SEXP*)s->data)[i] = elts[i];
    return s;
}
int mock_type(SEXP s) { return s->type; }
R_xlen_t mock_length(SEXP s) { return s->length; }
int mock_nrow(SEXP s) { return s->has_dim ? s->nrow : 0; }
int mock_ncol(SEXP s) { return s->has_dim ? s->ncol : 0; }
void* mock_data(SEXP s) { return s->data; }
SEXP mock_elt(SEXP s, int i) { return ((SEXP*)s->data)[i]; }
const char* mock_name(SEXP s, int i) {
    if (s->names == R_NilValue || i >= s->names->length) return "";
    return (const char*)((SEXP*)s->names->data)[i]->data;
}
const char* mock_class(SEXP s) { return s->klass == R_NilValue ? "" : (const char*)((SEXP*)s->klass->data)[0]->data; }
SEXP mock_rownames(SEXP s) { return s->rownames; }
/* gc(): run the finalizers of external pointers, free everything */
void mock_reset(void) {
    for (SEXP s = all_objects; s; s = s->next)
        if (s->type == EXTPTRSXP_ && s->finalizer && s->extptr) s->finalizer(s);
    while (all_objects) {
        SEXP s = all_objects; all_objects = s->next;
        free(s->data); free(s);
    }
    protect_depth = 0; nwarnings = 0; last_warning[0] = 0;
}
/* the finalizer of one external pointer, as R's gc would run it */
void mock_finalize(SEXP s) { if (s->type == EXTPTRSXP_ && s->finalizer) s->finalizer(s); }

static void __attribute__((constructor)) mock_init(void) {
    union { uint64_t u; double d; } v; v.u = 0x7FF8000000000000ULL; R_NaN = v.d;
}

/* R_registerRoutines / R_useDynamicSymbols: the table R_init_rfsib200() registers is kept so that
 * the tests can check the registration itself (names, arities) the way R's loader would see it */
#include <R_ext/Rdynload.h>
static const R_CallMethodDef* registered = NULL;
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)dll; (void)c; (void)f; (void)e; registered = call; return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; return v; }
int mock_registered(int i, const char** name, int* nargs) {
    if (!registered || !registered[i].name) return 0;
    *name = registered[i].name; *nargs = registered[i].numArgs; return 1;
}
