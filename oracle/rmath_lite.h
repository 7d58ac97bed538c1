// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// The reference reaches its sampling arithmetic through R's C API
// (`R::runif`, `R::rweibull`, `R::rexp`, `R::rnorm`, `R::gammafn`,
// `unif_rand`), i.e. libR's nmath/*.c and main/RNG.c.  libR is a third-party
// dependency that is NOT under /root/reference and R is not installed in this
// image; DESCRIPTION:13 pins no R version.  This file restates the published
// algorithms:
//   * set.seed() scrambling + Mersenne-Twister MT19937 (R's default kind)
//   * RNGkind("user"): unif_rand() = *user_unif_rand()  (src/microsimulation.cc:77-85)
//   * runif, rweibull, rexp/exp_rand (Ahrens-Dieter 1972), rnorm/norm_rand
//     (INVERSION) with qnorm (Wichura AS241), gammafn (Chebyshev series, 0<x<=10)
// Parity is anchored on the reference's own call sites and printed outputs
// (tests/test_oracle_golden.py), all produced with set.seed(12345).
#ifndef MSIM_ORACLE_RMATH_LITE_H
#define MSIM_ORACLE_RMATH_LITE_H

#include <stdint.h>

namespace rlite {

enum RngKind { MERSENNE_TWISTER = 0, USER_UNIF = 1 };

void RNGkind(RngKind k);
RngKind current_kind();
void set_seed(uint32_t seed);  // R: set.seed(seed) for Mersenne-Twister
uint64_t draws_taken();        // uniforms handed out since the last set_seed / reset
void reset_draw_counter();

// hook the "user" generator (R resolves `user_unif_rand` by symbol name)
void set_user_unif_rand(double *(*fn)());

}  // namespace rlite

double unif_rand();
double norm_rand();
double exp_rand();

namespace R {
double runif(double a, double b);
double rweibull(double shape, double scale);
double rexp(double scale);
double rnorm(double mu, double sigma);
double qnorm(double p, double mu, double sigma, int lower_tail, int log_p);
double gammafn(double x);
}  // namespace R

#endif
