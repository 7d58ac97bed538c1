// ORACLE — TEST INFRASTRUCTURE ONLY.  See gsm_lite.h.
#include "gsm_lite.h"

namespace orc {

static vec mat_vec(const std::vector<vec> &m, const vec &v) {
  vec out(m.size(), 0.0);
  for (size_t r = 0; r < m.size(); r++)
    for (size_t j = 0; j < v.size(); j++) out[r] += m[r][j] * v[j];
  return out;
}

SplineBasis::SplineBasis(int order) : order(order) {  // splines.cpp:11-16
  ordm1 = order - 1;
  rdel = vec(ordm1);
  ldel = vec(ordm1);
  a = vec(order);
  nknots = curs = boundary = ncoef = 0;
}
int SplineBasis::set_cursor(double x) {  // splines.cpp:25-41
  curs = -1;
  boundary = 0;
  for (int i = 0; i < nknots; i++) {
    if (knots[i] >= x) curs = i;
    if (knots[i] > x) break;
  }
  if (curs > nknots - order) {
    int lastLegit = nknots - order;
    if (x == knots[lastLegit]) {
      boundary = 1;
      curs = lastLegit;
    }
  }
  return curs;
}
void SplineBasis::diff_table(double x, int ndiff) {  // splines.cpp:42-50
  for (int i = 0; i < ndiff; i++) {
    rdel[i] = knots[curs + i] - x;
    ldel[i] = x - knots[curs - (i + 1)];
  }
}
double SplineBasis::slow_evaluate(double x, int nder) {  // splines.cpp:51-73
  int ti = curs, lpt, apt, rpt, inner, outer = ordm1;
  if (boundary && nder == ordm1) return double(0);
  while (nder--) {
    for (inner = outer, apt = 0, lpt = ti - outer; inner--; apt++, lpt++)
      a[apt] = double(outer) * (a[apt + 1] - a[apt]) / (knots[lpt + outer] - knots[lpt]);
    outer--;
  }
  diff_table(x, outer);
  while (outer--)
    for (apt = 0, lpt = outer, rpt = 0, inner = outer + 1; inner--; lpt--, rpt++, apt++)
      a[apt] = (a[apt + 1] * ldel[lpt] + a[apt] * rdel[rpt]) / (rdel[rpt] + ldel[lpt]);
  return a[0];
}
vec SplineBasis::basis_funcs(double x) {  // splines.cpp:75-97
  vec b(order);
  diff_table(x, ordm1);
  b[0] = double(1);
  for (size_t j = 1; j <= (size_t)ordm1; j++) {
    double saved = double(0);
    for (size_t r = 0; r < j; r++) {
      double den = rdel[r] + ldel[j - 1 - r];
      if (den != double(0)) {
        double term = b[r] / den;
        b[r] = saved + rdel[r] * term;
        saved = ldel[j - 1 - r] * term;
      } else {
        if (r != double(0) || rdel[r] != double(0)) b[r] = saved;
        saved = double(0);
      }
    }
    b[j] = saved;
  }
  return b;
}
vec SplineBasis::eval(double x, int ders) {  // splines.cpp:98-118
  vec val(ncoef, 0.0);
  set_cursor(x);
  int io = curs - order;
  if (io < 0 || io > nknots) {
    // the reference writes zeros at val(j+io): out of range for io < 0; nothing to do
  } else if (ders > 0) {
    for (size_t i = 0; i < (size_t)order; i++) {
      for (size_t j = 0; j < (size_t)order; j++) a[j] = double(0);
      a[i] = double(1);
      val[i + io] = slow_evaluate(x, ders);
    }
  } else {
    vec valtmp = basis_funcs(x);
    for (size_t i = 0; i < valtmp.size(); i++) val[i + io] = valtmp[i];
  }
  return val;
}

bs::bs(vec boundary_knots, vec interior_knots, int intercept)  // splines.cpp:131-144
    : SplineBasis(4), boundary_knots(boundary_knots), interior_knots(interior_knots), intercept(intercept) {
  df = intercept + 3 + (int)interior_knots.size();
  this->nknots = (int)interior_knots.size() + 8;
  this->ncoef = this->nknots - this->order;
  this->knots = vec(this->nknots);
  for (size_t i = 0; i < 4; i++) {
    this->knots[i] = boundary_knots[0];
    this->knots[this->nknots - i - 1] = boundary_knots[1];
  }
  for (size_t i = 0; i < interior_knots.size(); i++) this->knots[i + 4] = interior_knots[i];
}
static vec axpy(const vec &a, const vec &b, double s) {
  vec o(a);
  for (size_t i = 0; i < o.size(); i++) o[i] += b[i] * s;
  return o;
}
vec bs::eval(double x, int ders) {  // splines.cpp:145-167
  vec v;
  if (x < boundary_knots[0] || x > boundary_knots[1]) {
    double k_pivot = x < boundary_knots[0] ? double(0.75) * boundary_knots[0] + double(0.25) * interior_knots[0]
                                           : double(0.75) * boundary_knots[1] + double(0.25) * interior_knots[interior_knots.size() - 1];
    double delta = x - k_pivot;
    v = bs::eval(k_pivot, 0);
    v = axpy(v, bs::eval(k_pivot, 1), delta);
    v = axpy(v, bs::eval(k_pivot, 2), delta * delta / 2.);
    v = axpy(v, bs::eval(k_pivot, 3), delta * delta * delta / 6.);
  } else {
    vec full = SplineBasis::eval(x, ders);
    v = vec(full.begin() + (1 - intercept), full.begin() + (df - intercept) + 1);  // subvec(1-intercept, df-intercept)
  }
  return v;
}

ns::ns(vec boundary_knots, vec interior_knots, std::vector<vec> _q_matrix, int intercept, int cure)  // splines.cpp:194-203
    : bs(boundary_knots, interior_knots, intercept), q_matrix(_q_matrix), cure(cure) {
  if (!q_matrix.empty() && q_matrix[0].size() < q_matrix.size()) {  // q_matrix = q_matrix.t()
    std::vector<vec> t(q_matrix[0].size(), vec(q_matrix.size()));
    for (size_t i = 0; i < q_matrix.size(); i++)
      for (size_t j = 0; j < q_matrix[0].size(); j++) t[j][i] = q_matrix[i][j];
    q_matrix = t;
  }
  tl0 = mat_vec(q_matrix, bs::eval(boundary_knots[0], 0));
  tl1 = mat_vec(q_matrix, bs::eval(boundary_knots[0], 1));
  tr0 = mat_vec(q_matrix, bs::eval(boundary_knots[1], 0));
  tr1 = mat_vec(q_matrix, bs::eval(boundary_knots[1], 1));
}
vec ns::eval(double x, int der) {  // splines.cpp:204-219
  if (x < this->boundary_knots[0]) {
    if (der == 0) return axpy(tl0, tl1, x - this->boundary_knots[0]);
    else if (der == 1) return tl1;
    else return vec(tl1.size(), 0.0);
  } else if (x > this->boundary_knots[1]) {
    if (der == 0) return axpy(tr0, tr1, x - this->boundary_knots[1]);
    else if (der == 1) return tr1;
    else return vec(tr1.size(), 0.0);
  } else
    return mat_vec(q_matrix, bs::eval(x, der));
}

static double dot(const vec &a, const vec &b) {
  double s = 0;
  for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
  return s;
}
double gsm::eta(double y) {  // gsm.cpp:15-21
  double eta = etap[index];
  for (size_t i = 0; i < terms.size(); i++)
    if (terms[i].x[index] != 0.0) eta += terms[i].x[index] * dot(terms[i].ns1.eval(y, 0), terms[i].gamma);
  return eta;
}
double gsm::eta0(double y) {  // gsm.cpp:22-28
  double eta = etap0[index];
  for (size_t i = 0; i < terms.size(); i++)
    if (terms[i].x[index] != 0.0) eta += terms[i].x[index] * dot(terms[i].ns1.eval(y, 0), terms[i].gamma);
  return eta;
}
double gsm::operator()(double y) {  // gsm.cpp:29-31
  return ((log_time ? exp(y) : y) < t0) ? (eta(y) - target) : (eta0(y) - target0);
}
double gsm::randU(double u, double tentry, int index, double scale) {  // gsm.cpp:38-47
  double ymin = tentry == 0.0 ? (log_time ? log(tmin / scale) : tmin / scale) : (log_time ? log(tentry) : tentry);
  double ymax = log_time ? log(tmax * scale) : tmax * scale;
  this->index = index;
  this->target = (tentry == 0.0 ? link(u) : link(u * linkinv(eta(ymin))));
  this->target0 = 0.0;
  double root = std::get<0>(ssim::R_zeroin2_functor_ptr<gsm>(ymin, ymax, this, 1.0e-8, 100));
  return log_time ? exp(root) : root;
}
double gsm::randU0(double u, int index, double scale) {  // gsm.cpp:48-58
  double ymin = log_time ? log(tmin / scale) : tmin / scale;
  double ymax = log_time ? log(tmax * scale) : tmax * scale;
  this->index = index;
  this->target = link(u);
  this->target0 = link(u * linkinv(eta0(log_time ? log(t0) : t0)) / linkinv(eta(log_time ? log(t0) : t0)));
  double root = std::get<0>(ssim::R_zeroin2_functor_ptr<gsm>(ymin, ymax, this, 1.0e-8, 100));
  return log_time ? exp(root) : root;
}

}  // namespace orc
