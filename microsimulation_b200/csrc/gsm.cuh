// Generalised survival model: survival-time inversion as a device function
// (north_star (e)).  What the reference computes with
//   gsm::eta / operator() / randU / randU0      src/gsm.cpp:15-58
//   ns::eval (natural cubic spline basis, linear beyond the boundary knots)   src/splines.cpp:194-219
//   R_zeroin2_functor_ptr (Brent)                inst/include/c_optim.h:9-105
// S(t) = u is solved for a natural-cubic-spline log cumulative hazard eta(log t) (link PH: log(-log S)), one
// inversion per thread.
//
// The spline is NOT evaluated the reference's way (B-spline basis by the de Boor recurrence at every call, then the
// projection matrix, then the coefficients: ~n_ns x df multiply-adds per evaluation).  A term's contribution
// f(y) = ns::eval(y) . gamma is a piecewise cubic with a line on either side of the boundary knots, so the host
// expands it ONCE per model into polynomial coefficients per knot interval (Cox-de Boor carried out on polynomials,
// projection matrix and gamma folded in: gsm_term_build below), and an evaluation on the device is a short search
// over the breakpoints and three multiply-adds.  Parity for this row is unpinned in the reference (no expected value
// in the tree); the oracle restates the reference's evaluation and the two agree to rounding (tests/test_gsm.py).
//
// The model is a flat struct (no pointers except the per-covariate-pattern
// vectors) so that it can be passed by value to a kernel; msim_gsm (msim.h) is
// its serialised form.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "fastmath.cuh"
#include "mrg32k3a.cuh"

namespace msim {

constexpr int GSM_MAX_TERMS = 4;
constexpr int GSM_MAX_INTERIOR = 12;                 // interior knots per term
constexpr int GSM_MAX_PIECES = GSM_MAX_INTERIOR + 3;  // a line, up to n_interior + 1 cubics, a line
constexpr int GSM_MAX_NS = GSM_MAX_INTERIOR + 4;      // rows of the natural-spline projection the host accepts

struct GsmTerm {
  int n_pieces;                 // piece 0: y < b0; pieces 1 .. n_pieces-2: [brk[p], brk[p+1]); piece n_pieces-1: y > b1
  double b0, b1;                // Boundary_knots
  double brk[GSM_MAX_PIECES];   // left end of piece p (p >= 1), ascending; brk[n_pieces-1] = b1
  double org[GSM_MAX_PIECES];   // origin of piece p's polynomial (b0 for piece 0, b1 for the last, else brk[p])
  double c[GSM_MAX_PIECES][4];  // f(y) = c0 + h (c1 + h (c2 + h c3)), h = y - org[p]
};

struct GsmModel {
  int n_terms, log_time, n_index;
  double tmin, tmax, t0;
  GsmTerm terms[GSM_MAX_TERMS];
  const double *etap, *etap0;          // [n_index]
  const double *x[GSM_MAX_TERMS];      // [n_index] each
};

// ns::eval(y, 0) . gamma of one term (splines.cpp:204-219 and gsm.cpp:19), from the expanded form
MSIM_HD double ns_dot_gamma(const GsmTerm &T, double y) {
  int p;
  if (y < T.b0) {
    p = 0;
  } else if (y > T.b1) {
    p = T.n_pieces - 1;
  } else {
    p = 1;
    while (p + 2 < T.n_pieces && y >= T.brk[p + 1]) p++;
  }
  const double h = y - T.org[p];
  return MSIM_FMA(h, MSIM_FMA(h, MSIM_FMA(h, T.c[p][3], T.c[p][2]), T.c[p][1]), T.c[p][0]);
}

// ---- model set-up (host) ----------------------------------------------------

// What ns::ns + bs::bs hold for a term (splines.cpp:145-203): boundary and interior knots, whether the first
// B-spline is kept (intercept), the projection matrix q (n_ns x df, df = intercept + 3 + n_interior) and gamma.
struct GsmTermSpec {
  int n_interior, intercept, n_ns;
  double b0, b1;
  const double *interior;  // ascending, strictly inside (b0, b1)
  const double *q;         // row-major n_ns x df
  const double *gamma;     // n_ns
};

// Expands a term into per-interval polynomials.  On the knot interval [t_j, t_j+1) the four cubic B-splines that are
// non-zero there are built by the Cox-de Boor recurrence
//   N_{i,k}(x) = (x - t_i)/(t_{i+k-1} - t_i) N_{i,k-1}(x) + (t_{i+k} - x)/(t_{i+k} - t_{i+1}) N_{i+1,k-1}(x)
// with the N's held as polynomials in h = x - t_j (terms with a zero denominator vanish), then weighted by
// w_i = sum_r gamma_r q[r][i - 1 + intercept] (bs::eval drops the first B-spline unless intercept) and summed.
// The lines beyond the boundary knots continue value and slope there (ns::eval, splines.cpp:205-217).
inline bool gsm_term_build(GsmTerm &T, const GsmTermSpec &S) {
  const int ni = S.n_interior, m = ni + 8, ncoef = m - 4, df = S.intercept + 3 + ni;
  if (ni < 0 || ni > GSM_MAX_INTERIOR || S.n_ns < 1 || S.n_ns > GSM_MAX_NS || !(S.b0 < S.b1)) return false;
  double t[GSM_MAX_INTERIOR + 8], w[GSM_MAX_INTERIOR + 4];
  for (int k = 0; k < 4; k++) t[k] = S.b0, t[m - 1 - k] = S.b1;
  for (int k = 0; k < ni; k++) {
    t[4 + k] = S.interior[k];
    if (!(t[4 + k] > t[3 + k]) || !(t[4 + k] < S.b1)) return false;  // ascending, strictly inside
  }
  for (int i = 0; i < ncoef; i++) {
    const int col = i - 1 + S.intercept;
    w[i] = 0.0;
    if (col >= 0 && col < df)
      for (int r = 0; r < S.n_ns; r++) w[i] += S.gamma[r] * S.q[r * df + col];
  }
  T.b0 = S.b0;
  T.b1 = S.b1;
  int np = 1;  // piece 0 is the left line, filled in below
  for (int j = 3; j < m - 4; j++) {
    if (!(t[j] < t[j + 1])) continue;
    // N[k][a]: polynomial (coefficients of h^0..h^3) of N_{j-k+1+a, k} for a = 0..k-1, on this interval
    double N[5][4][4] = {};
    N[1][0][0] = 1.0;
    for (int k = 2; k <= 4; k++)
      for (int a = 0; a < k; a++) {
        const int i = j - k + 1 + a;
        double out[4] = {0, 0, 0, 0};
        if (a >= 1) {  // first term: (x - t_i)/(t_{i+k-1} - t_i) * N_{i,k-1}; N_{i,k-1} is entry a-1 of order k-1
          const double den = t[i + k - 1] - t[i];
          if (den != 0.0) {
            const double *P = N[k - 1][a - 1];
            const double s0 = (t[j] - t[i]) / den, s1 = 1.0 / den;  // (x - t_i)/den = s0 + s1 h
            for (int d = 0; d < 4; d++) {
              out[d] += s0 * P[d];
              if (d + 1 < 4) out[d + 1] += s1 * P[d];
            }
          }
        }
        if (a <= k - 2) {  // second term: (t_{i+k} - x)/(t_{i+k} - t_{i+1}) * N_{i+1,k-1}; entry a of order k-1
          const double den = t[i + k] - t[i + 1];
          if (den != 0.0) {
            const double *P = N[k - 1][a];
            const double s0 = (t[i + k] - t[j]) / den, s1 = -1.0 / den;
            for (int d = 0; d < 4; d++) {
              out[d] += s0 * P[d];
              if (d + 1 < 4) out[d + 1] += s1 * P[d];
            }
          }
        }
        for (int d = 0; d < 4; d++) N[k][a][d] = out[d];
      }
    if (np >= GSM_MAX_PIECES - 1) return false;
    T.brk[np] = T.org[np] = t[j];
    for (int d = 0; d < 4; d++) {
      double v = 0.0;
      for (int a = 0; a < 4; a++) v += w[j - 3 + a] * N[4][a][d];
      T.c[np][d] = v;
    }
    np++;
  }
  // the lines: value and slope at the boundary knots
  const int last = np - 1;
  const double hl = S.b1 - T.org[last];
  const double *cl = T.c[last];
  T.brk[0] = T.org[0] = S.b0;
  T.c[0][0] = T.c[1][0];
  T.c[0][1] = T.c[1][1];
  T.c[0][2] = T.c[0][3] = 0.0;
  T.brk[np] = T.org[np] = S.b1;
  T.c[np][0] = cl[0] + hl * (cl[1] + hl * (cl[2] + hl * cl[3]));
  T.c[np][1] = cl[1] + hl * (2.0 * cl[2] + hl * 3.0 * cl[3]);
  T.c[np][2] = T.c[np][3] = 0.0;
  T.n_pieces = np + 1;
  for (int p = T.n_pieces; p < GSM_MAX_PIECES; p++) {
    T.brk[p] = T.org[p] = S.b1;
    for (int d = 0; d < 4; d++) T.c[p][d] = 0.0;
  }
  return true;
}

// ---- gsm ------------------------------------------------------------------

MSIM_HD double gsm_link(double S) { return fm_log(-fm_log(S)); }    // PH, gsm.cpp:8-10 (exp / log: fastmath.cuh)
MSIM_HD double gsm_linkinv(double eta) { return fm_exp(-fm_exp(eta)); }   // gsm.cpp:11-13

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

}  // namespace msim
