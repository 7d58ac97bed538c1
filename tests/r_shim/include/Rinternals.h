/* Stand-in for R's <Rinternals.h>: the subset of the R API glue/msim_glue.c uses, with R's names, types and
 * semantics (vectors with a type tag and attributes; a global environment; non-returning Rf_error). */
#ifndef MSIM_RSHIM_RINTERNALS_H
#define MSIM_RSHIM_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
#define NILSXP 0
#define SYMSXP 1
#define ENVSXP 4
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_UnboundValue, R_GlobalEnv, R_NamesSymbol, R_ClassSymbol, R_RowNamesSymbol;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
int TYPEOF(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
int Rf_length(SEXP x);
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP Rf_mkChar(const char *s);
SEXP Rf_mkString(const char *s);
const char *R_CHAR(SEXP x);
#define CHAR(x) R_CHAR(x)
SEXP Rf_install(const char *name);
SEXP Rf_findVar(SEXP symbol, SEXP env);
void Rf_defineVar(SEXP symbol, SEXP value, SEXP env);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value);
SEXP Rf_getAttrib(SEXP x, SEXP name);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#ifdef __cplusplus
}
#endif
#endif
