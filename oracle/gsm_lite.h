// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's generalised survival model:
//   SplineBasis / bs / ns     src/splines.cpp:11-229, inst/include/splines.h
//   gsm                        src/gsm.cpp:6-93, inst/include/gsm.h
//   R_zeroin2_functor_ptr      inst/include/c_optim.h:9-105 — in the `ref` build the
//                              reference's own header is included as it lies;
//                              the portable build restates it below.
// Armadillo (absent here) is replaced by std::vector; statement order follows the
// reference.  PARITY UNPINNED: the reference holds no expected value for this
// code (test/test_microsimulation.R:737-774 only times it against rstpm2, which
// is absent); tests cross-check the spline basis against scipy's BSpline and the
// inversion through S(randU(u)) = u.
#ifndef MSIM_ORACLE_GSM_LITE_H
#define MSIM_ORACLE_GSM_LITE_H

#include <math.h>
#include <stddef.h>

#include <tuple>
#include <vector>

#ifdef MSIM_ORACLE_USE_REFERENCE
#include <c_optim.h>  // the reference's, from $(REF)/inst/include
#else
#include <float.h>
namespace ssim {
template <typename Functor>
std::tuple<double, double, int> R_zeroin2_functor_ptr(double ax, double bx, Functor *f, double Tol, int Maxit) {
  using Result = std::tuple<double, double, int>;
  double a, b, c, fa, fb, fc, tol;
  int maxit;
  a = ax; b = bx; fa = (*f)(a); fb = (*f)(b);
  c = a; fc = fa;
  maxit = Maxit + 1; tol = Tol;
  if (fa == 0.0) return Result(a, 0.0, 0);
  if (fb == 0.0) return Result(b, 0.0, 0);
  while (maxit--) {
    double prev_step = b - a, tol_act, p, q, new_step;
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    tol_act = 2 * DBL_EPSILON * fabs(b) + tol / 2;
    new_step = (c - b) / 2;
    if (fabs(new_step) <= tol_act || fb == (double)0) return Result(b, Maxit - maxit, fabs(c - b));
    if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
      double t1, cb, t2;
      cb = c - b;
      if (a == c) {
        t1 = fb / fa;
        p = cb * t1;
        q = 1.0 - t1;
      } else {
        q = fa / fc; t1 = fb / fc; t2 = fb / fa;
        p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
        q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
      }
      if (p > (double)0) q = -q;
      else p = -p;
      if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2)) new_step = p / q;
    }
    if (fabs(new_step) < tol_act) {
      if (new_step > (double)0) new_step = tol_act;
      else new_step = -tol_act;
    }
    a = b; fa = fb;
    b += new_step; fb = (*f)(b);
    if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) {
      c = a; fc = fa;
    }
  }
  return Result(b, fabs(c - b), -1);
}
}  // namespace ssim
#endif

namespace orc {

typedef std::vector<double> vec;

class SplineBasis {
 public:
  int order, ordm1, nknots, curs, boundary, ncoef;
  vec ldel, rdel, knots, a;
  SplineBasis(int order = 4);
  int set_cursor(double x);
  void diff_table(double x, int ndiff);
  double slow_evaluate(double x, int nder);
  vec basis_funcs(double x);
  vec eval(double x, int ders = 0);
};

class bs : public SplineBasis {
 public:
  vec boundary_knots, interior_knots;
  int intercept, df;
  bs() {}
  bs(vec boundary_knots, vec interior_knots, int intercept = 0);
  vec eval(double x, int ders = 0);
};

class ns : public bs {
 public:
  vec tl0, tl1, tr0, tr1;
  std::vector<vec> q_matrix;  // rows
  int cure;
  ns() {}
  ns(vec boundary_knots, vec interior_knots, std::vector<vec> q_matrix, int intercept = 0, int cure = 0);
  vec eval(double x, int der);
};

struct gsm_term {
  ns ns1;
  vec gamma, x;
};

class gsm {
 public:
  double tmin, tmax, target, target0, t0;
  vec etap, etap0;
  std::vector<gsm_term> terms;
  int index;
  bool log_time;
  double link(double S) { return log(-log(S)); }
  double linkinv(double eta) { return exp(-exp(eta)); }
  double eta(double y);
  double eta0(double y);
  double operator()(double y);
  double randU(double u, double tentry = 0.0, int index = 0, double scale = 10.0);
  double randU0(double u, int index = 0, double scale = 10.0);
};

}  // namespace orc

#endif
