// Host side of the survival-time utilities and lookup tables (lookups.cuh): tables up, one evaluation per thread.
#pragma once
#include "host_runtime.h"

namespace msim_host {
using namespace msim;

// uploads `vecs` (concatenated) to g.gsm_data and returns device pointers to each
inline int upload_vectors(const std::vector<std::vector<double> > &vecs, std::vector<const double *> *ptrs) {
  int rc;
  size_t total = 0;
  for (const auto &v : vecs) total += v.size();
  std::vector<double> flat;
  flat.reserve(total);
  for (const auto &v : vecs) flat.insert(flat.end(), v.begin(), v.end());
  if ((rc = ensure(g.gsm_data, sizeof(double) * std::max<size_t>(total, 1)))) return rc;
  CK(cudaMemcpy(g.gsm_data.p, flat.data(), sizeof(double) * total, cudaMemcpyHostToDevice));
  const double *d = (const double *)g.gsm_data.p;
  size_t off = 0;
  for (const auto &v : vecs) {
    ptrs->push_back(d + off);
    off += v.size();
  }
  return MSIM_OK;
}

// inputs a, b, c (n_in doubles each; b, c may be null) up, lookup_kernel, out down
inline int lookup_run(LookupParams &p, int64_t n, int64_t n_in, const double *a, const double *b, const double *c, double *out) {
  int rc;
  if (n == 0) return MSIM_OK;
  const size_t nn = (size_t)n_in;
  if ((rc = ensure(g.gsm_in, sizeof(double) * 3 * nn))) return rc;
  if ((rc = ensure(g.gsm_out, sizeof(double) * (size_t)n))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  double *d = (double *)g.gsm_in.p;
  CK(cudaMemcpyAsync(d, a, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  if (b) CK(cudaMemcpyAsync(d + nn, b, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  if (c) CK(cudaMemcpyAsync(d + 2 * nn, c, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  p.n = n;
  p.a = d;
  p.b = b ? d + nn : nullptr;
  p.c = c ? d + 2 * nn : nullptr;
  p.out = (double *)g.gsm_out.p;
  p.error = small_view().error;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)g.sm_count * 8));
  lookup_kernel<<<grid, 256, 0, g.stream>>>(p);
  CK(cudaGetLastError());
  g.launches++;
  CK(cudaMemcpyAsync(out, g.gsm_out.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  int err = 0;
  CK(cudaMemcpy(&err, small_view().error, sizeof err, cudaMemcpyDeviceToHost));
  if (err) return fail(MSIM_ERR_ARG, "Rpexp.rand() failed due to zero hazards");  // Rcpp::stop in the reference
  return MSIM_OK;
}

}  // namespace msim_host
