/* Stand-in for R's <R.h>, just enough for glue/msim_glue.c (test infrastructure: R is not in the image). */
#ifndef MSIM_RSHIM_R_H
#define MSIM_RSHIM_R_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void REprintf(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#ifdef __cplusplus
}
#endif
#endif
