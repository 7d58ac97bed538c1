// Survival-time utilities and lookup tables of the reference as device
// functions (what the external FHCRC screening model is
// built from):
//   find_interval                      inst/include/microsimulation.h:505-509 (R's findInterval)
//   Rpexp        H, h, S, p, d, rand   inst/include/microsimulation.h:527-617
//   Rpexp_gamma  H, p, d, h, rand      inst/include/microsimulation.h:626-704
//   NumericInterpolate approx / () / invert / invert_decreasing   inst/include/rcpp_table.h:52-111
//   Table<keys..., Outcome>::operator() inst/include/rcpp_table.h:114-324 (floor lookup on every axis)
// Vectors are plain pointers (device memory in kernels, host memory in the CPU
// emulation); the host builds the cumulative vectors exactly as the
// constructors do.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mrg32k3a.cuh"

namespace msim {

// index of the last element <= x, 0 if there is none (vec ascending)
MSIM_HD int find_interval(double x, const double *vec, int n) {
  int lo = 0, hi = n;  // count of elements <= x
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (vec[mid] <= x) lo = mid + 1;
    else hi = mid;
  }
  return lo == 0 ? 0 : lo - 1;
}

// ---- Rpexp: piecewise constant hazards hi on [ti[i], ti[i+1]) ---------------
struct Rpexp {
  int n;
  const double *Hi, *hi, *ti;  // Hi[i] = cumulative hazard at ti[i]
};
// Hi as the constructor computes it (:541-543)
inline void rpexp_cumulate(int n, const double *hi, const double *ti, double *Hi) {
  Hi[0] = 0.0;
  for (int i = 1; i < n; i++) Hi[i] = Hi[i - 1] + (ti[i] - ti[i - 1]) * hi[i - 1];
}
MSIM_HD double rpexp_H(const Rpexp &T, double s) {
  if (s == 0.0) return 0.0;
  int i = find_interval(s, T.ti, T.n);
  return T.Hi[i] + (s - T.ti[i]) * T.hi[i];
}
MSIM_HD double rpexp_h(const Rpexp &T, double s) { return T.hi[find_interval(s, T.ti, T.n)]; }
MSIM_HD double rpexp_S(const Rpexp &T, double s) { return exp(-rpexp_H(T, s)); }
MSIM_HD double rpexp_p(const Rpexp &T, double s, bool lower_tail) {
  return lower_tail ? 1.0 - exp(-rpexp_H(T, s)) : exp(-rpexp_H(T, s));
}
MSIM_HD double rpexp_d(const Rpexp &T, double s) { return rpexp_h(T, s) * rpexp_S(T, s); }
// rand(u, from) (:595-607).  *failed is set where the reference calls Rcpp::stop ("zero hazards").
MSIM_HD double rpexp_rand(const Rpexp &T, double u, double from, int *failed) {
  double v = -log(u) + rpexp_H(T, from);
  int i = find_interval(v, T.Hi, T.n);
  double tstar = T.ti[i];
  bool success = false;
  for (int j = i; j < T.n && !success; j++) {
    if (T.hi[j] > 0.0) {
      tstar += (v - T.Hi[i]) / T.hi[j];
      success = true;
    }
  }
  if (!success) *failed = 1;
  return tstar;
}

// ---- Rpexp_gamma: the same with a gamma frailty of variance theta ----------
struct RpexpGamma {
  int n;
  double theta;
  const double *H0i, *mui, *ti;
};
inline void rpexp_gamma_cumulate(int n, double theta, const double *mui, const double *ti, double *H0i) {  // :645-649
  H0i[0] = 0.0;
  double Mu = 0.0;
  for (int i = 1; i < n; i++) {
    Mu += mui[i - 1] * (ti[i] - ti[i - 1]);
    H0i[i] = (exp(theta * Mu) - 1.0) / theta;
  }
}
MSIM_HD double rpexp_gamma_H(const RpexpGamma &T, double a, double Z) {  // :669-672
  int k = find_interval(a, T.ti, T.n);
  return Z * ((1 + T.theta * T.H0i[k]) * exp(T.theta * T.mui[k] * (a - T.ti[k])) - 1.0) / T.theta;
}
MSIM_HD double rpexp_gamma_p(const RpexpGamma &T, double a, double Z, bool lower_tail) {
  return lower_tail ? -expm1(-rpexp_gamma_H(T, a, Z)) : exp(-rpexp_gamma_H(T, a, Z));
}
MSIM_HD double rpexp_gamma_h(const RpexpGamma &T, double a, double Z) {  // :682-685
  int k = find_interval(a, T.ti, T.n);
  return Z * (1 + T.theta * T.H0i[k]) * exp(T.theta * T.mui[k] * (a - T.ti[k])) * T.mui[k];
}
MSIM_HD double rpexp_gamma_rand(const RpexpGamma &T, double u, double Z, double from) {  // :686-690
  double v = -log(u) / Z + rpexp_gamma_H(T, from, 1.0);
  int k = find_interval(v, T.H0i, T.n);
  return T.ti[k] + 1.0 / T.mui[k] / T.theta * log((1 + T.theta * v) / (1 + T.theta * T.H0i[k]));
}

// ---- NumericInterpolate ------------------------------------------------------
struct Interp {
  int n;
  const double *x, *y, *slope;  // slope[i] = (y[i+1]-y[i])/(x[i+1]-x[i]), n-1 values
};
// std::lower_bound: first index with v[i] >= key (ascending v)
MSIM_HD int lower_bound_idx(const double *v, int n, double key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (v[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
MSIM_HD double interp_approx(const Interp &T, double xfind) {  // rcpp_table.h:75-83
  if (xfind <= T.x[0]) return T.y[0];
  if (xfind >= T.x[T.n - 1]) return T.y[T.n - 1] + T.slope[T.n - 2] * (xfind - T.x[T.n - 1]);
  int i = lower_bound_idx(T.x, T.n, xfind) - 1;
  return T.y[i] + T.slope[i] * (xfind - T.x[i]);
}
MSIM_HD double interp_step(const Interp &T, double xfind) {  // operator(), :84-88
  if (xfind <= T.x[0]) return T.y[0];
  int i = lower_bound_idx(T.x, T.n, xfind);
  return T.y[i - 1];
}
MSIM_HD double interp_invert(const Interp &T, double yfind) {  // :89-97, y increasing
  if (yfind <= T.y[0]) return T.x[0];
  if (yfind >= T.y[T.n - 1]) return T.x[T.n - 1] + (yfind - T.y[T.n - 1]) / T.slope[T.n - 2];
  int i = lower_bound_idx(T.y, T.n, yfind) - 1;
  return T.x[i] + (yfind - T.y[i]) / T.slope[i];
}
MSIM_HD double interp_invert_decreasing(const Interp &T, double yfind) {  // :102-110
  if (yfind >= T.y[0]) return T.x[0];
  if (yfind < T.y[T.n - 1]) return T.x[T.n - 1] + (yfind - T.y[T.n - 1]) / T.slope[T.n - 2];
  int lo = 0, hi = T.n;  // lower_bound with greater<>: first index with !(y[i] > yfind)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (T.y[mid] > yfind) lo = mid + 1;
    else hi = mid;
  }
  int i = lo - 1;
  return T.x[i] + (yfind - T.y[i]) / T.slope[i];
}

// ---- Table<I0..Ik, Outcome> ---------------------------------------------------
// Every axis holds its sorted distinct key values; a lookup takes, on each axis,
// the largest key <= the argument (the smallest key if the argument lies below
// all of them: set_lower_bound, rcpp_table.h:114-117) and reads the outcome of
// that combination — 0 where the data frame held no such row (std::map's
// default-constructed value).  Dense row-major grid over the axes.
constexpr int TABLE_MAX_AXES = 5;
struct DenseTable {
  int n_axes;
  int n[TABLE_MAX_AXES];
  const double *axis[TABLE_MAX_AXES];
  const double *values;
};
MSIM_HD double table_lookup(const DenseTable &T, const double *key) {
  int flat = 0;
  for (int a = 0; a < T.n_axes; a++) flat = flat * T.n[a] + find_interval(key[a], T.axis[a], T.n[a]);
  return T.values[flat];
}

}  // namespace msim
