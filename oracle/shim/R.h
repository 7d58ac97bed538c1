/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for R's <R.h>, needed only so the reference's
 * src/ssim.cc (which includes <R.h> for Rprintf in a debug dump,
 * ssim.cc:29, 81-86) compiles in this R-less image. */
#ifndef MSIM_ORACLE_R_SHIM_H
#define MSIM_ORACLE_R_SHIM_H
#include <stdio.h>
#define Rprintf(...) printf(__VA_ARGS__)
#define REprintf(...) fprintf(stderr, __VA_ARGS__)
#endif
