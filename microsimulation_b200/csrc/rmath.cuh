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
    double a = 0.;
    double u = unif_rand();
    while (u <= 0. || u >= 1.) u = unif_rand();
    for (;;) {
      u += u;
      if (u > 1.) break;
      a += q0;
    }
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
