// Host side of the generalised survival model (gsm.cuh): model upload and batched inversions.
#pragma once
#include "host_runtime.h"

namespace msim_host {
using namespace msim;

// ---- generalised survival model

// msim_gsm (the Rcpp::List fields, gsm.cpp:60-92) -> the flat device model; vectors indexed by covariate pattern
// go to one device buffer
int gsm_upload(const msim_gsm *m) {
  int rc;
  if (!m || m->link != 0) return fail(MSIM_ERR_ARG, "only link \"PH\" is implemented (as in the reference)");
  if (m->n_terms < 0 || m->n_terms > GSM_MAX_TERMS || m->n_index < 1 || !m->etap || !m->etap0 || !(m->inflate > 0))
    return fail(MSIM_ERR_ARG, "bad gsm model");
  std::vector<GsmModel> hostv(1);
  GsmModel &M = hostv[0];
  memset(&M, 0, sizeof M);
  M.n_terms = m->n_terms;
  M.log_time = m->log_time != 0;
  M.n_index = m->n_index;
  M.tmin = m->tmin / m->inflate;
  M.tmax = m->tmax * m->inflate;
  M.t0 = m->t0;
  const size_t ni = (size_t)m->n_index;
  std::vector<double> data((2 + (size_t)m->n_terms) * ni);
  memcpy(&data[0], m->etap, sizeof(double) * ni);
  memcpy(&data[ni], m->etap0, sizeof(double) * ni);
  for (int i = 0; i < m->n_terms; i++) {
    const msim_gsm_term &t = m->terms[i];
    GsmTerm &T = M.terms[i];
    if (t.n_knots < 1 || t.n_knots > GSM_MAX_INTERIOR || !t.gamma || !t.knots || !t.q_const || !t.x)
      return fail(MSIM_ERR_ARG, "gsm term: 1..12 interior knots and non-null vectors are required");
    // q_matrix = q_const, transposed if it has fewer columns than rows (splines.cpp:197-198)
    const int df = (t.intercept != 0) + 3 + t.n_knots;
    const bool tr = t.q_cols < t.q_rows;
    const int rows = tr ? t.q_cols : t.q_rows, cols = tr ? t.q_rows : t.q_cols;
    if (rows != t.n_gamma || cols != df || rows > GSM_MAX_NS) return fail(MSIM_ERR_ARG, "gsm term: q_const does not match gamma / knots");
    std::vector<double> q((size_t)rows * cols);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) q[(size_t)r * cols + c] = tr ? t.q_const[c * t.q_cols + r] : t.q_const[r * t.q_cols + c];
    GsmTermSpec spec;
    spec.n_interior = t.n_knots;
    spec.intercept = t.intercept != 0;
    spec.n_ns = rows;
    spec.b0 = t.Boundary_knots[0];
    spec.b1 = t.Boundary_knots[1];
    spec.interior = t.knots;
    spec.q = q.data();
    spec.gamma = t.gamma;
    // the term as polynomials per knot interval, projection and coefficients folded in (gsm.cuh)
    if (!gsm_term_build(T, spec)) return fail(MSIM_ERR_ARG, "gsm term: interior knots must ascend strictly inside the boundary knots");
    memcpy(&data[(2 + (size_t)i) * ni], t.x, sizeof(double) * ni);
  }
  if ((rc = ensure(g.gsm_data, sizeof(double) * data.size()))) return rc;
  if ((rc = ensure(g.gsm_model, sizeof(GsmModel)))) return rc;
  const double *dd = (const double *)g.gsm_data.p;
  M.etap = dd;
  M.etap0 = dd + ni;
  for (int i = 0; i < m->n_terms; i++) M.x[i] = dd + (2 + (size_t)i) * ni;
  CK(cudaMemcpy(g.gsm_data.p, data.data(), sizeof(double) * data.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.gsm_model.p, &M, sizeof M, cudaMemcpyHostToDevice));
  return MSIM_OK;
}

int gsm_run(const msim_gsm *m, int mode, int64_t n, const double *u, const double *tentry, const int32_t *index, double scale,
            double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n < 0 || (n > 0 && (!u || !out))) return fail(MSIM_ERR_ARG, "bad arguments");
  if (index)
    for (int64_t i = 0; i < n; i++)
      if (index[i] < 0 || (m && index[i] >= m->n_index)) return fail(MSIM_ERR_ARG, "index out of range");
  if ((rc = gsm_upload(m))) return rc;
  if (n == 0) return MSIM_OK;
  // inputs: u | tentry | index(int32, padded to doubles)
  const size_t nn = (size_t)n;
  if ((rc = ensure(g.gsm_in, sizeof(double) * 3 * nn))) return rc;
  if ((rc = ensure(g.gsm_out, sizeof(double) * nn))) return rc;
  double *d_u = (double *)g.gsm_in.p, *d_t = d_u + nn;
  int *d_i = (int *)(d_u + 2 * nn);
  CK(cudaMemcpyAsync(d_u, u, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  if (tentry) CK(cudaMemcpyAsync(d_t, tentry, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  if (index) CK(cudaMemcpyAsync(d_i, index, sizeof(int) * nn, cudaMemcpyHostToDevice, g.stream));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)g.sm_count * 16));
  CK(cudaEventRecord(g.ev0, g.stream));
  gsm_kernel<<<grid, 128, 0, g.stream>>>((const GsmModel *)g.gsm_model.p, mode, n, d_u, tentry ? d_t : nullptr,
                                         index ? d_i : nullptr, scale, (double *)g.gsm_out.p);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev1, g.stream));
  g.launches++;
  CK(cudaMemcpyAsync(out, g.gsm_out.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
  return MSIM_OK;
}

}  // namespace msim_host
