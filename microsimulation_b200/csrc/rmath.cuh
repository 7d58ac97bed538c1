// Device restatement of the R samplers the reference's models call through
// `R::` (libR nmath; call sites person-r.cc:100-167, illness-death.cpp:18-62,
// simple-example.cc:29-59, calibperson-r.cc:96-152, microsimulation.cc:6-8).
// Algorithms: SURVEY.md Appendix A.3.  All arithmetic is IEEE double evaluated
// in the written order; this translation unit is compiled with -fmad=false so
// no multiply-add is contracted.
//
// `Src` is any type with `double unif_rand()` returning a value strictly
// inside (0,1): the per-person MRG32k3a substream (RNGkind("user"),
// microsimulation.cc:77-85) or a cursor into a pre-generated Mersenne-Twister
// stream (RNGkind("Mersenne-Twister")).
#pragma once
#include <math.h>

#include "fastmath.cuh"
#include "mrg32k3a.cuh"

namespace msim {

// Ahrens & Dieter (1972): q[k-1] = sum_{i=1..k} ln(2)^i / i!
#ifdef __CUDACC__
__constant__
#else
static
#endif
    const double MSIM_EXP_Q[16] = {0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
                                   0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
                                   0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
                                   0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};

// exp_rand's first loop adds ln 2 once per doubling of u: MSIM_EXP_A[k] = ln 2 + ln 2 + ... (k times, rounded at every
// addition as the loop does: from k = 25 on that is not k * ln 2)
#ifdef __CUDACC__
__constant__
#else
static
#endif
    const double MSIM_EXP_A[64] = {0.0, 0.6931471805599453, 1.3862943611198906, 2.0794415416798357, 2.772588722239781, 3.4657359027997265, 4.1588830833596715, 4.852030263919617, 5.545177444479562, 6.238324625039508, 6.931471805599453, 7.6246189861593985, 8.317766166719343, 9.010913347279288, 9.704060527839234, 10.39720770839918, 11.090354888959125, 11.78350206951907, 12.476649250079015, 13.16979643063896, 13.862943611198906, 14.556090791758852, 15.249237972318797, 15.942385152878742, 16.635532333438686, 17.32867951399863, 18.021826694558573, 18.714973875118517, 19.40812105567846, 20.101268236238404, 20.794415416798348, 21.48756259735829, 22.180709777918235, 22.87385695847818, 23.567004139038122, 24.260151319598066, 24.95329850015801, 25.646445680717953, 26.339592861277897, 27.03274004183784, 27.725887222397784, 28.419034402957728, 29.11218158351767, 29.805328764077615, 30.49847594463756, 31.191623125197502, 31.884770305757446, 32.57791748631739, 33.271064666877336, 33.964211847437284, 34.65735902799723, 35.35050620855718, 36.043653389117125, 36.73680056967707, 37.42994775023702, 38.12309493079697, 38.816242111356914, 39.50938929191686, 40.20253647247681, 40.895683653036755, 41.5888308335967, 42.28197801415665, 42.9751251947166, 43.668272375276544};

// the `R::` namespace as seen from a model: R.runif(a,b), R.rweibull(...), ...
template <class Src>
struct RMath {
  Src *src;
  MSIM_DEV explicit RMath(Src *s) : src(s) {}

  MSIM_DEV double unif_rand() { return src->unif_rand(); }

  MSIM_DEV double runif(double a, double b) {
    if (a == b) return a;
    double u;
    do {
      u = unif_rand();
    } while (u <= 0 || u >= 1);
    return a + (b - a) * u;
  }

  MSIM_DEV double rweibull(double shape, double scale) {
    return scale * fm_pow(-fm_log(unif_rand()), 1. / shape);  // exp / log / pow: fastmath.cuh
  }

  MSIM_DEV double exp_rand() {
    const double q0 = 0.6931471805599453;  // MSIM_EXP_Q[0] = ln 2
    const double *q = MSIM_EXP_Q;
    double u = unif_rand();
    while (u <= 0. || u >= 1.) u = unif_rand();
    // R's loop — for (;;) { u += u; if (u > 1.) break; a += q0; } — doubles u until it exceeds 1 and adds ln 2 for every
    // doubling but the last.  Doubling is exact, so the number of doublings j follows from u's exponent (u = m 2^e,
    // 1 <= m < 2, e < 0: j = -e, one more if m = 1), the doubled u is u with j added to its exponent field, and a is
    // the (j - 1)-fold sum from the table.  (As a loop a warp runs as many passes as its unluckiest lane needs.)
    const int64_t ub = fm_bits(u);
    const int j = 1023 - (int)(ub >> 52) + ((ub & 0x000fffffffffffffll) == 0 ? 1 : 0);
    double a = MSIM_EXP_A[j - 1 < 63 ? j - 1 : 63];
    // u < 2^-64: no generator here produces it (RngStream's and R's uniforms stay above 2^-33), so the loop is kept as a
    // loop: unrolled, as the compiler has it otherwise, it sat — 78 instructions, never executed — in the middle of
    // the person-r kernel's loop
#pragma unroll 1
    for (int k = 63; k < j - 1; k++) a += q0;
    u = fm_from_bits(ub + ((int64_t)j << 52));
    u -= 1.;
    if (u <= q0) return a + u;
    int i = 0;
    double ustar = unif_rand(), umin = ustar;
    do {
      ustar = unif_rand();
      if (umin > ustar) umin = ustar;
      i++;
    } while (u > q[i]);
    return a + umin * q0;
  }

  MSIM_DEV double rexp(double scale) { return scale * exp_rand(); }

  // default normal kind INVERSION: always two uniforms
  MSIM_DEV double norm_rand() {
    const double BIG = 134217728;  // 2^27
    double u = unif_rand();
    u = (int)(BIG * u) + unif_rand();
    return qnorm_std(u / BIG);
  }

  MSIM_DEV double rnorm(double mu, double sigma) {
    if (sigma == 0.) return mu;  // no draw
    return mu + sigma * norm_rand();
  }

  // rweibullHR, src/microsimulation.cc:6-8
  MSIM_DEV double rweibullHR(double shape, double scale, double hr) {
    return rweibull(shape, scale * pow(hr, 1.0 / shape));
  }

  // Wichura AS241, lower tail, mu = 0, sigma = 1 (0.0 + 1.0*val == val)
  MSIM_DEV_NOINLINE static double qnorm_std(double p) {
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
    return val;
  }
};

// gammafn(x) for 0 < x <= 10 is only needed for model constants
// (illness-death.cpp:17-19) and is evaluated on the host: see host_gammafn in
// capi.cu.

}  // namespace msim
