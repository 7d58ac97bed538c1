// Generalised survival model: survival-time inversion as a device function
// (north_star (e)).  Follows the reference's
//   gsm::eta / operator() / randU / randU0      src/gsm.cpp:15-58
//   ns::eval, bs::eval, SplineBasis::set_cursor /
//   basis_funcs / slow_evaluate                  src/splines.cpp:24-118, 145-167, 204-219
//   R_zeroin2_functor_ptr (Brent)                inst/include/c_optim.h:9-105
// with Armadillo vectors replaced by fixed-capacity arrays.  S(t) = u is
// solved for a natural-cubic-spline log cumulative hazard eta(log t) (link PH:
// log(-log S)), one inversion per thread.
//
// The model is a flat struct (no pointers except the per-covariate-pattern
// vectors) so that it can be passed by value to a kernel; msim_gsm (msim.h) is
// its serialised form.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "mrg32k3a.cuh"

namespace msim {

constexpr int GSM_MAX_TERMS = 4;
constexpr int GSM_MAX_INTERIOR = 12;                  // interior knots per term
constexpr int GSM_MAX_COEF = GSM_MAX_INTERIOR + 4;    // B-spline coefficients (order 4)
constexpr int GSM_MAX_KNOTS = GSM_MAX_INTERIOR + 8;   // full knot vector
constexpr int GSM_MAX_NS = GSM_MAX_COEF;              // rows of the natural-spline projection

struct GsmTerm {
  int n_interior, intercept, cure, n_ns, df;  // df = intercept + 3 + n_interior = length of bs::eval
  int nknots, ncoef;
  double knots[GSM_MAX_KNOTS];
  double b0, b1;                          // Boundary_knots
  double q[GSM_MAX_NS][GSM_MAX_COEF];     // q_matrix, n_ns x df
  double gamma[GSM_MAX_NS];
  double tl0[GSM_MAX_NS], tl1[GSM_MAX_NS], tr0[GSM_MAX_NS], tr1[GSM_MAX_NS];
};

struct GsmModel {
  int n_terms, log_time, n_index;
  double tmin, tmax, t0;
  GsmTerm terms[GSM_MAX_TERMS];
  const double *etap, *etap0;          // [n_index]
  const double *x[GSM_MAX_TERMS];      // [n_index] each
};

// ---- SplineBasis (order 4) ------------------------------------------------

struct SplineCursor {
  int curs, boundary;
};

// splines.cpp:24-41
MSIM_HD SplineCursor spline_set_cursor(const GsmTerm &T, double x) {
  SplineCursor c;
  c.curs = -1;
  c.boundary = 0;
  for (int i = 0; i < T.nknots; i++) {
    if (T.knots[i] >= x) c.curs = i;
    if (T.knots[i] > x) break;
  }
  if (c.curs > T.nknots - 4) {
    int lastLegit = T.nknots - 4;
    if (x == T.knots[lastLegit]) {
      c.boundary = 1;
      c.curs = lastLegit;
    }
  }
  return c;
}

// splines.cpp:74-97: the four cubic B-splines that are non-zero at x
MSIM_HD void spline_basis_funcs(const GsmTerm &T, int curs, double x, double b[4]) {
  double rdel[3], ldel[3];
  for (int i = 0; i < 3; i++) {  // diff_table, :42-50
    rdel[i] = T.knots[curs + i] - x;
    ldel[i] = x - T.knots[curs - (i + 1)];
  }
  b[0] = 1.0;
  for (int j = 1; j <= 3; j++) {
    double saved = 0.0;
    for (int r = 0; r < j; r++) {
      double den = rdel[r] + ldel[j - 1 - r];
      if (den != 0.0) {
        double term = b[r] / den;
        b[r] = saved + rdel[r] * term;
        saved = ldel[j - 1 - r] * term;
      } else {
        if (r != 0 || rdel[r] != 0.0) b[r] = saved;
        saved = 0.0;
      }
    }
    b[j] = saved;
  }
}

// splines.cpp:51-73: value of the nder-th derivative of sum_i a[i] B_i at x (a is scratch, overwritten)
MSIM_HD double spline_slow_evaluate(const GsmTerm &T, const SplineCursor &c, double x, int nder, double a[4]) {
  int ti = c.curs, lpt, apt, rpt, inner, outer = 3;
  if (c.boundary && nder == 3) return 0.0;
  while (nder--) {
    for (inner = outer, apt = 0, lpt = ti - outer; inner--; apt++, lpt++)
      a[apt] = (double)outer * (a[apt + 1] - a[apt]) / (T.knots[lpt + outer] - T.knots[lpt]);
    outer--;
  }
  double rdel[3], ldel[3];
  for (int i = 0; i < outer; i++) {
    rdel[i] = T.knots[c.curs + i] - x;
    ldel[i] = x - T.knots[c.curs - (i + 1)];
  }
  while (outer--)
    for (apt = 0, lpt = outer, rpt = 0, inner = outer + 1; inner--; lpt--, rpt++, apt++)
      a[apt] = (a[apt + 1] * ldel[lpt] + a[apt] * rdel[rpt]) / (rdel[rpt] + ldel[lpt]);
  return a[0];
}

// SplineBasis::eval (splines.cpp:98-118): all ncoef basis values (or derivatives) at x
MSIM_HD void spline_eval(const GsmTerm &T, double x, int ders, double val[GSM_MAX_COEF]) {
  for (int i = 0; i < T.ncoef; i++) val[i] = 0.0;
  SplineCursor c = spline_set_cursor(T, x);
  int io = c.curs - 4;
  if (io < 0 || io > T.nknots) return;
  if (ders > 0) {
    for (int i = 0; i < 4; i++) {
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      a[i] = 1.0;
      val[i + io] = spline_slow_evaluate(T, c, x, ders, a);
    }
  } else {
    double b[4];
    spline_basis_funcs(T, c.curs, x, b);
    for (int i = 0; i < 4; i++) val[i + io] = b[i];
  }
}

// bs::eval for x inside the boundary knots (splines.cpp:163-165): drop the first column unless intercept
MSIM_HD void bs_eval_inside(const GsmTerm &T, double x, int ders, double v[GSM_MAX_COEF]) {
  double val[GSM_MAX_COEF];
  spline_eval(T, x, ders, val);
  for (int i = 0; i < T.df; i++) v[i] = val[i + 1 - T.intercept];
}

// ns::eval(x, 0) . gamma   (splines.cpp:204-219 and gsm.cpp:19)
MSIM_HD double ns_dot_gamma(const GsmTerm &T, double x) {
  double s = 0.0;
  if (x < T.b0) {
    for (int r = 0; r < T.n_ns; r++) s += (T.tl0[r] + (x - T.b0) * T.tl1[r]) * T.gamma[r];
  } else if (x > T.b1) {
    for (int r = 0; r < T.n_ns; r++) s += (T.tr0[r] + (x - T.b1) * T.tr1[r]) * T.gamma[r];
  } else {
    double v[GSM_MAX_COEF];
    bs_eval_inside(T, x, 0, v);
    for (int r = 0; r < T.n_ns; r++) {
      double e = 0.0;
      for (int j = 0; j < T.df; j++) e += T.q[r][j] * v[j];  // q_matrix * bs::eval
      s += e * T.gamma[r];
    }
  }
  return s;
}

// ---- gsm ------------------------------------------------------------------

MSIM_HD double gsm_link(double S) { return log(-log(S)); }          // PH, gsm.cpp:8-10
MSIM_HD double gsm_linkinv(double eta) { return exp(-exp(eta)); }   // gsm.cpp:11-13

MSIM_HD double gsm_eta(const GsmModel &M, int index, double y, bool baseline0) {  // gsm.cpp:15-28
  double eta = baseline0 ? M.etap0[index] : M.etap[index];
  for (int i = 0; i < M.n_terms; i++) {
    double xi = M.x[i][index];
    if (xi != 0.0) eta += xi * ns_dot_gamma(M.terms[i], y);
  }
  return eta;
}

struct GsmTarget {
  const GsmModel *M;
  int index;
  double target, target0;
  MSIM_HD double operator()(double y) const {  // gsm.cpp:29-31
    return ((M->log_time ? exp(y) : y) < M->t0) ? (gsm_eta(*M, index, y, false) - target)
                                               : (gsm_eta(*M, index, y, true) - target0);
  }
};

// R_zeroin2_functor_ptr (c_optim.h:9-105): Brent's root finder exactly as R's zeroin.c, written as a state machine
// that ASKS for function values (x = where f is wanted next) instead of calling f: the arithmetic and its order are
// zeroin.c's, but a scheduler may interleave the evaluations of many searches (kernels.cuh, gsm_kernel: every lane
// evaluates the spline once per pass, whatever stage its search is in).  zeroin2() below is the plain loop over it.
struct Brent {
  double a, b, c, fa, fb, fc, tol;
  double x;     // evaluate f here and feed the value
  double root;  // valid once done
  int maxit, phase;
  bool done;
};
MSIM_HD void brent_begin(Brent &s, double ax, double bx, double Tol, int Maxit) {
  s.a = ax;
  s.b = bx;
  s.tol = Tol;
  s.maxit = Maxit + 1;
  s.phase = 0;
  s.x = ax;
  s.done = false;
  s.root = bx;
}
MSIM_HD void brent_feed(Brent &s, double value) {
  if (s.phase == 0) {  // fa = f(a)
    s.fa = value;
    s.phase = 1;
    s.x = s.b;
    return;
  }
  if (s.phase == 1) {  // fb = f(b); c = a
    s.fb = value;
    s.c = s.a;
    s.fc = s.fa;
    s.phase = 2;
    if (s.fa == 0.0) {
      s.root = s.a;
      s.done = true;
      return;
    }
    if (s.fb == 0.0) {
      s.root = s.b;
      s.done = true;
      return;
    }
  } else {  // the end of the loop body: fb = f(b) after the step
    s.fb = value;
    if ((s.fb > 0 && s.fc > 0) || (s.fb < 0 && s.fc < 0)) {
      s.c = s.a;
      s.fc = s.fa;
    }
  }
  // while (maxit--)
  if (s.maxit == 0) {
    s.root = s.b;  // not converged within Maxit: the last approximation, as the reference returns
    s.done = true;
    return;
  }
  s.maxit--;
  double prev_step = s.b - s.a, tol_act, p, q, new_step;
  if (fabs(s.fc) < fabs(s.fb)) {
    s.a = s.b; s.b = s.c; s.c = s.a;
    s.fa = s.fb; s.fb = s.fc; s.fc = s.fa;
  }
  tol_act = 2 * DBL_EPSILON * fabs(s.b) + s.tol / 2;
  new_step = (s.c - s.b) / 2;
  if (fabs(new_step) <= tol_act || s.fb == 0.0) {
    s.root = s.b;
    s.done = true;
    return;
  }
  if (fabs(prev_step) >= tol_act && fabs(s.fa) > fabs(s.fb)) {
    double t1, cb, t2;
    cb = s.c - s.b;
    if (s.a == s.c) {
      t1 = s.fb / s.fa;
      p = cb * t1;
      q = 1.0 - t1;
    } else {
      q = s.fa / s.fc; t1 = s.fb / s.fc; t2 = s.fb / s.fa;
      p = t2 * (cb * q * (q - t1) - (s.b - s.a) * (t1 - 1.0));
      q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
    }
    if (p > 0.0) q = -q;
    else p = -p;
    if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2)) new_step = p / q;
  }
  if (fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
  s.a = s.b; s.fa = s.fb;
  s.b += new_step;
  s.x = s.b;  // fb = f(b) comes with the next feed
}
template <class F>
MSIM_HD double zeroin2(double ax, double bx, const F &f, double Tol, int Maxit) {
  Brent s;
  brent_begin(s, ax, bx, Tol, Maxit);
  while (!s.done) brent_feed(s, f(s.x));
  return s.root;
}

// gsm::randU (gsm.cpp:38-47)
MSIM_HD double gsm_randU(const GsmModel &M, double u, double tentry, int index, double scale) {
  double ymin = tentry == 0.0 ? (M.log_time ? log(M.tmin / scale) : M.tmin / scale) : (M.log_time ? log(tentry) : tentry);
  double ymax = M.log_time ? log(M.tmax * scale) : M.tmax * scale;
  GsmTarget f;
  f.M = &M;
  f.index = index;
  f.target = (tentry == 0.0) ? gsm_link(u) : gsm_link(u * gsm_linkinv(gsm_eta(M, index, ymin, false)));
  f.target0 = 0.0;
  double root = zeroin2(ymin, ymax, f, 1.0e-8, 100);
  return M.log_time ? exp(root) : root;
}

// gsm::randU0 (gsm.cpp:48-58)
MSIM_HD double gsm_randU0(const GsmModel &M, double u, int index, double scale) {
  double ymin = M.log_time ? log(M.tmin / scale) : M.tmin / scale;
  double ymax = M.log_time ? log(M.tmax * scale) : M.tmax * scale;
  GsmTarget f;
  f.M = &M;
  f.index = index;
  f.target = gsm_link(u);
  const double y0 = M.log_time ? log(M.t0) : M.t0;
  f.target0 = gsm_link(u * gsm_linkinv(gsm_eta(M, index, y0, true)) / gsm_linkinv(gsm_eta(M, index, y0, false)));
  double root = zeroin2(ymin, ymax, f, 1.0e-8, 100);
  return M.log_time ? exp(root) : root;
}

// ns::ns (splines.cpp:194-203): the linear-extrapolation vectors, from bs::eval and its first derivative at the
// boundary knots.  Host only (model set-up).
inline void gsm_term_finish(GsmTerm &T) {
  T.df = T.intercept + 3 + T.n_interior;
  T.nknots = T.n_interior + 8;
  T.ncoef = T.nknots - 4;
  double v0[GSM_MAX_COEF], v1[GSM_MAX_COEF];
  for (int side = 0; side < 2; side++) {
    double xb = side ? T.b1 : T.b0;
    bs_eval_inside(T, xb, 0, v0);
    bs_eval_inside(T, xb, 1, v1);
    for (int r = 0; r < T.n_ns; r++) {
      double e0 = 0, e1 = 0;
      for (int j = 0; j < T.df; j++) {
        e0 += T.q[r][j] * v0[j];
        e1 += T.q[r][j] * v1[j];
      }
      (side ? T.tr0 : T.tl0)[r] = e0;
      (side ? T.tr1 : T.tl1)[r] = e1;
    }
  }
}

}  // namespace msim
