// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's survival-time utilities and lookup tables:
//   find_interval (R's findInterval)        inst/include/microsimulation.h:505-509
//   Rpexp, Rpexp_gamma                      inst/include/microsimulation.h:527-704
//   NumericInterpolate, set_lower_bound, Table   inst/include/rcpp_table.h:52-324
// std::vector / std::map / std::set as in the reference, Rcpp types replaced by
// plain vectors.  PARITY UNPINNED by printed values in the reference; the dev
// script's R implementation rpexp2 (test/test_microsimulation.R:268-295) states
// the intended semantics and is re-done in numpy by tests/test_lookups.py.
#ifndef MSIM_ORACLE_LOOKUPS_LITE_H
#define MSIM_ORACLE_LOOKUPS_LITE_H

#include <math.h>

#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <vector>

namespace orc {

// R's findInterval(vec, n, x, rightmost_closed = FALSE, all_inside = FALSE): number of elements <= x
inline int findInterval(const double *xt, int n, double x) {
  return (int)(std::upper_bound(xt, xt + n, x) - xt);
}
inline int find_interval(double x, std::vector<double> vec) {  // microsimulation.h:505-509
  int new_ilo = findInterval(&vec[0], (int)vec.size(), x);
  return new_ilo == 0 ? 0 : new_ilo - 1;
}

class Rpexp {  // microsimulation.h:527-617
 public:
  std::vector<double> Hi, hi, ti;
  size_t n;
  Rpexp(const double *hin, const double *tin, int nin) : n(nin) {
    Hi.resize(n);
    ti.resize(n);
    hi.resize(n);
    Hi[0] = 0.0;
    hi[0] = hin[0];
    ti[0] = tin[0];
    if (n > 1) {
      for (size_t i = 1; i < n; i++) {
        hi[i] = hin[i];
        ti[i] = tin[i];
      }
      for (size_t i = 1; i < n; i++) Hi[i] = Hi[i - 1] + (ti[i] - ti[i - 1]) * hi[i - 1];
    }
  }
  double H(double s) {
    if (s == 0.0) return 0.0;
    int i = find_interval(s, ti);
    return Hi[i] + (s - ti[i]) * hi[i];
  }
  double h(double s) { return hi[find_interval(s, ti)]; }
  double S(double s) { return exp(-H(s)); }
  double p(double s, bool lower_tail = true) { return lower_tail ? 1.0 - exp(-H(s)) : exp(-H(s)); }
  double d(double s) { return h(s) * S(s); }
  double rand(double u, double from = 0.0) {
    double v = -log(u) + H(from);
    int i = find_interval(v, Hi);
    double tstar = ti[i];
    bool success = false;
    for (size_t j = i; j < n && !success; j++) {
      if (hi[j] > 0.0) {
        tstar += (v - Hi[i]) / hi[j];
        success = true;
      }
    }
    if (!success) throw std::runtime_error("Rpexp.rand() failed due to zero hazards");
    return tstar;
  }
};

class Rpexp_gamma {  // microsimulation.h:626-704
 public:
  std::vector<double> H0i, mui, ti;
  double theta;
  size_t n;
  Rpexp_gamma(double theta, const double *muin, const double *tin, size_t n) : theta(theta), n(n) {
    H0i.resize(n);
    ti.resize(n);
    mui.resize(n);
    H0i[0] = 0.0;
    mui[0] = muin[0];
    ti[0] = tin[0];
    if (n > 1) {
      for (size_t i = 1; i < n; i++) {
        mui[i] = muin[i];
        ti[i] = tin[i];
      }
      double Mu = 0.0;
      for (size_t i = 1; i < n; i++) {
        Mu += mui[i - 1] * (ti[i] - ti[i - 1]);
        H0i[i] = (exp(theta * Mu) - 1.0) / theta;
      }
    }
  }
  double H(double a, double Z) {
    size_t index = find_interval(a, ti);
    return Z * ((1 + theta * H0i[index]) * exp(theta * mui[index] * (a - ti[index])) - 1.0) / theta;
  }
  double p(double a, double Z, bool lower_tail = true) { return lower_tail ? -expm1(-H(a, Z)) : exp(-H(a, Z)); }
  double h(double a, double Z) {
    size_t index = find_interval(a, ti);
    return Z * (1 + theta * H0i[index]) * exp(theta * mui[index] * (a - ti[index])) * mui[index];
  }
  double rand(double u, double Z, double from = 0.0) {
    double v = -log(u) / Z + H(from, 1.0);
    size_t index = find_interval(v, H0i);
    return ti[index] + 1.0 / mui[index] / theta * log((1 + theta * v) / (1 + theta * H0i[index]));
  }
};

class NumericInterpolate {  // rcpp_table.h:52-111
 public:
  std::vector<double> x, y, slope;
  int n;
  NumericInterpolate(const double *xin, const double *yin, int nin) : x(xin, xin + nin), y(yin, yin + nin), n(nin) {
    for (int i = 0; i < n - 1; i++) slope.push_back((y[i + 1] - y[i]) / (x[i + 1] - x[i]));
  }
  double approx(double xfind) {
    int i;
    if (xfind <= x[0]) return y[0];
    else if (xfind >= x[n - 1]) return y[n - 1] + slope[n - 2] * (xfind - x[n - 1]);
    else {
      i = (int)(std::lower_bound(x.begin(), x.end(), xfind) - 1 - x.begin());
      return y[i] + slope[i] * (xfind - x[i]);
    }
  }
  double operator()(double xfind) {
    if (xfind <= x[0]) return y[0];
    int i = (int)(std::lower_bound(x.begin(), x.end(), xfind) - x.begin());
    return y[--i];
  }
  double invert(double yfind) {
    int i;
    if (yfind <= y[0]) return x[0];
    else if (yfind >= y[n - 1]) return x[n - 1] + (yfind - y[n - 1]) / slope[n - 2];
    else {
      i = (int)(std::lower_bound(y.begin(), y.end(), yfind) - 1 - y.begin());
      return x[i] + (yfind - y[i]) / slope[i];
    }
  }
  double invert_decreasing(double yfind) {
    int i;
    if (yfind >= y[0]) return x[0];
    else if (yfind < y[n - 1]) return x[n - 1] + (yfind - y[n - 1]) / slope[n - 2];
    else {
      i = (int)(std::lower_bound(y.begin(), y.end(), yfind, std::greater<double>()) - 1 - y.begin());
      return x[i] + (yfind - y[i]) / slope[i];
    }
  }
};

template <class T>
T set_lower_bound(std::set<T, std::greater<T> > aset, T value) {  // rcpp_table.h:114-117
  return value < *aset.rbegin() ? *aset.rbegin() : *aset.lower_bound(value);
}

// Table<double x k, double> for any k <= 5 (the reference spells out one template specialisation per k)
class Table {
 public:
  std::vector<std::set<double, std::greater<double> > > axis;
  std::map<std::vector<double>, double> data;
  explicit Table(int k) : axis(k) {}
  void insert(const std::vector<double> &key, double outcome) {
    for (size_t a = 0; a < axis.size(); a++) axis[a].insert(key[a]);
    data[key] = outcome;
  }
  double operator()(const std::vector<double> &k) {
    std::vector<double> key(k.size());
    for (size_t a = 0; a < axis.size(); a++) key[a] = set_lower_bound(axis[a], k[a]);
    return data[key];
  }
};

}  // namespace orc

#endif
