// ORACLE — TEST INFRASTRUCTURE ONLY (see rmath_lite.h).
// Restated from the published R algorithms;
// compile with -ffp-contract=off: every expression is evaluated as written.
#include "rmath_lite.h"

#include <math.h>

namespace rlite {

namespace {

const int MT_N = 624, MT_M = 397;
uint32_t mt_state[MT_N];
int mt_pos = MT_N + 1;
RngKind g_kind = MERSENNE_TWISTER;
uint64_t g_draws = 0;
double *(*g_user)() = 0;

void mt_refill() {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t y;
  int kk;
  for (kk = 0; kk < MT_N - MT_M; kk++) {
    y = (mt_state[kk] & 0x80000000u) | (mt_state[kk + 1] & 0x7fffffffu);
    mt_state[kk] = mt_state[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
  }
  for (; kk < MT_N - 1; kk++) {
    y = (mt_state[kk] & 0x80000000u) | (mt_state[kk + 1] & 0x7fffffffu);
    mt_state[kk] = mt_state[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
  }
  y = (mt_state[MT_N - 1] & 0x80000000u) | (mt_state[0] & 0x7fffffffu);
  mt_state[MT_N - 1] = mt_state[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
  mt_pos = 0;
}

double mt_next() {
  if (mt_pos >= MT_N) {
    if (mt_pos == MT_N + 1) set_seed(4357);  // never seeded
    mt_refill();
  }
  uint32_t y = mt_state[mt_pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return (double)y * 2.3283064365386963e-10;  // [0,1)
}

// keep results strictly inside (0,1)
double fixup(double x) {
  const double i2_32m1 = 2.328306437080797e-10;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

}  // namespace

void RNGkind(RngKind k) { g_kind = k; }
RngKind current_kind() { return g_kind; }
uint64_t draws_taken() { return g_draws; }
void reset_draw_counter() { g_draws = 0; }
void set_user_unif_rand(double *(*fn)()) { g_user = fn; }

void set_seed(uint32_t seed) {
  // initial scrambling, then one LCG step per seed word; word 0 is the
  // position counter, which FixupSeeds forces to N ("regenerate first")
  for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
  seed = 69069u * seed + 1u;  // dummy[0], overwritten by mti = N
  for (int j = 0; j < MT_N; j++) {
    seed = 69069u * seed + 1u;
    mt_state[j] = seed;
  }
  mt_pos = MT_N;
  g_draws = 0;
}

double next_uniform();
double next_uniform() {
  g_draws++;
  if (g_kind == USER_UNIF) return fixup(*g_user());
  return fixup(mt_next());
}

}  // namespace rlite

double unif_rand() { return rlite::next_uniform(); }

// Ahrens & Dieter (1972) exponential sampler, as in R's sexp.c
double exp_rand() {
  // q[k-1] = sum_{i=1..k} ln(2)^i / i!
  static const double q[] = {
      0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
      0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
      0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
      0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
  double a = 0.;
  double u = unif_rand();
  while (u <= 0. || u >= 1.) u = unif_rand();
  for (;;) {
    u += u;
    if (u > 1.) break;
    a += q[0];
  }
  u -= 1.;
  if (u <= q[0]) return a + u;
  int i = 0;
  double ustar = unif_rand(), umin = ustar;
  do {
    ustar = unif_rand();
    if (umin > ustar) umin = ustar;
    i++;
  } while (u > q[i]);
  return a + umin * q[0];
}

// default normal kind INVERSION: two uniforms -> 53-bit-ish p -> qnorm
double norm_rand() {
  const double BIG = 134217728;  // 2^27
  double u = unif_rand();
  u = (int)(BIG * u) + unif_rand();
  return R::qnorm(u / BIG, 0.0, 1.0, 1, 0);
}

namespace R {

double runif(double a, double b) {
  if (a == b) return a;
  double u;
  do {
    u = unif_rand();
  } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}

double rweibull(double shape, double scale) { return scale * pow(-log(unif_rand()), 1. / shape); }

double rexp(double scale) { return scale * exp_rand(); }

double rnorm(double mu, double sigma) {
  if (sigma == 0.) return mu;  // no draw
  return mu + sigma * norm_rand();
}

// Wichura's AS241 (PPND16); only the lower-tail, non-log form is reachable
// from the reference's call sites.
double qnorm(double p, double mu, double sigma, int lower_tail, int log_p) {
  (void)lower_tail;
  (void)log_p;
  if (p == 0) return -INFINITY;
  if (p == 1) return INFINITY;
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = .180625 - q * q;
    val = q *
          (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
               45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
            133.14166789178437745) * r + 3.387132872796366608) /
          (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
               21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
            42.313330701600911252) * r + 1.);
  } else {
    r = (q < 0) ? p : 1 - p;
    r = sqrt(-log(r));
    if (r <= 5.) {
      r += -1.6;
      val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
              4.6303378461565452959) * r + 1.42343711074968357734) /
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
              2.05319162663775882187) * r + 1.);
    } else {
      r += -5.;
      val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
              5.4637849111641143699) * r + 6.6579046435011037772) /
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
              .59983220655588793769) * r + 1.);
    }
    if (q < 0.0) val = -val;
  }
  return mu + sigma * val;
}

// gamma function for 0 < x <= 10 (the reference only evaluates
// gammafn(1+1/a) for a = 3, 4: illness-death.cpp:17-19)
double gammafn(double x) {
  static const double gamcs[22] = {
      +.8571195590989331421920062399942e-2,  +.4415381324841006757191315771652e-2,
      +.5685043681599363378632664588789e-1,  -.4219835396418560501012500186624e-2,
      +.1326808181212460220584006796352e-2,  -.1893024529798880432523947023886e-3,
      +.3606925327441245256578082217225e-4,  -.6056761904460864218485548290365e-5,
      +.1055829546302283344731823509093e-5,  -.1811967365542384048291855891166e-6,
      +.3117724964715322277790254593169e-7,  -.5354219639019687140874081024347e-8,
      +.9193275519859588946887786825940e-9,  -.1577941280288339761767423273953e-9,
      +.2707980622934954543266540433089e-10, -.4646818653825730144081661058933e-11,
      +.7973350192007419656460767175359e-12, -.1368078209830916025799499172309e-12,
      +.2347319486563800657233471771688e-13, -.4027432614949066932766570534699e-14,
      +.6910051747372100912138336975257e-15, -.1185584500221992907052387126192e-15};
  const int ngam = 22;
  int n = (int)x;
  if (x < 0) --n;
  double y = x - n;  // 0 <= y < 1
  --n;
  // chebyshev_eval(2y-1, gamcs, ngam)
  double t = y * 2 - 1, twox = t * 2, b0 = 0, b1 = 0, b2 = 0;
  for (int i = 1; i <= ngam; i++) {
    b2 = b1;
    b1 = b0;
    b0 = twox * b1 - b2 + gamcs[ngam - i];
  }
  double value = .9375 + (b0 - b2) * 0.5;
  if (n == 0) return value;  // 1 <= x < 2
  if (n < 0) {               // 0 < x < 1
    n = -n;
    for (int i = 0; i < n; i++) value /= (x + i);
    return value;
  }
  for (int i = 1; i <= n; i++) value *= (y + i);  // 2 <= x <= 10
  return value;
}

}  // namespace R
