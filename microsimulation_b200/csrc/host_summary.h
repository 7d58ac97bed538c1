// Host side of the SummaryReport model (Sick-Sicker, summary_kernels.cuh).
#pragma once
#include "host_cohort.h"

namespace msim_host {
using namespace msim;

// ---- Sick-Sicker (SummaryReport)

// One pass over the n persons: first-pass codes, carry scan, the persons (both regimes)
template <bool ORDERED>
int sick_launch(SickParams &p, int64_t n, bool substreams, const HostStream &hs, int64_t first_person, int grid, size_t smem,
                int64_t *end_word) {
  int rc;
  typedef SickTraitsT<ORDERED> Traits;
  if (!substreams) {
    const int max_draws = 250;
    bool short_plan = false;
    int64_t mt_end = 0;
    // 36.4 draws per person at the README's rates; plan for 48 and fall back to the per-person maximum
    int64_t n_pos = std::min<int64_t>(48 * n + 8192, (int64_t)max_draws * n);
    if ((rc = stream_offsets_attempt<Traits>(n, p.par, n_pos, max_draws, true, &short_plan, end_word, &mt_end))) return rc;
    if (short_plan && n_pos < (int64_t)max_draws * n) {
      n_pos = (int64_t)max_draws * n;
      if ((rc = stream_offsets_attempt<Traits>(n, p.par, n_pos, max_draws, true, &short_plan, end_word, &mt_end))) return rc;
    }
    if (short_plan) return fail(MSIM_ERR_STREAM, "a person drew more uniforms than the sequential-stream plan allows (250)");
    CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
    p.mt_words = (const uint32_t *)g.mt_words.p;
    p.mt_end = mt_end;
    p.starts = (const int64_t *)g.starts.p;
    gather_code_kernel<<<g.sm_count * 4, 256, 0, g.stream>>>((const unsigned char *)g.aux.p, p.starts, (int64_t)g.rng.mt[0], n,
                                                             p.code_out);
    CK(cudaGetLastError());
    {
      const unsigned tiles = (unsigned)((n + CARRY_TILE - 1) / CARRY_TILE);
      if ((rc = ensure(g.carry_tiles, tiles))) return rc;
      carry_tile_kernel<<<tiles, 256, 0, g.stream>>>(p.code_out, n, (unsigned char *)g.carry_tiles.p);
      CK(cudaGetLastError());
      carry_apply_kernel<<<tiles, 256, 0, g.stream>>>(p.code_out, n, 0, (const unsigned char *)g.carry_tiles.p,
                                                      (unsigned char *)g.code_in.p);
      CK(cudaGetLastError());
      g.launches++;
    }
    auto kern = sick_kernel<true, ORDERED>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches += 3;
  } else {
    // A shard that does not start at person 0 inherits the report's utility and cost from the persons before it
    // (the reference never resets them between persons): the last set_uc() code of persons [0, first_person), found
    // by running the code pass backwards from first_person in windows until one holds a non-zero code.
    int code0 = 0;
    for (int64_t hi_p = first_person, w = 256; hi_p > 0 && code0 == 0; w *= 16) {
      const int64_t lo_p = std::max<int64_t>(0, hi_p - w), cnt = hi_p - lo_p;
      SickParams q = p;
      q.n = cnt;
      q.first_person = lo_p;
      if ((rc = ensure(g.aux, (size_t)cnt))) return rc;
      q.code_out = (unsigned char *)g.aux.p;
      const int wgrid = (int)std::max<int64_t>(1, std::min<int64_t>((cnt + BLOCK - 1) / BLOCK, grid));
      stream_start(hs.Cg.s, lo_p, (int64_t)wgrid * BLOCK, &q.st);
      sick_code_kernel<ORDERED><<<wgrid, BLOCK, 0, g.stream>>>(q);
      CK(cudaGetLastError());
      g.launches++;
      std::vector<unsigned char> codes((size_t)cnt);
      CK(cudaMemcpyAsync(codes.data(), g.aux.p, (size_t)cnt, cudaMemcpyDeviceToHost, g.stream));
      CK(cudaStreamSynchronize(g.stream));
      for (int64_t k = cnt - 1; k >= 0 && code0 == 0; k--) code0 = codes[(size_t)k];
      hi_p = lo_p;
    }
    stream_start(hs.Cg.s, first_person, (int64_t)grid * BLOCK, &p.st);
    sick_code_kernel<ORDERED><<<grid, BLOCK, 0, g.stream>>>(p);
    CK(cudaGetLastError());
    {
      const unsigned tiles = (unsigned)((n + CARRY_TILE - 1) / CARRY_TILE);
      if ((rc = ensure(g.carry_tiles, tiles))) return rc;
      carry_tile_kernel<<<tiles, 256, 0, g.stream>>>(p.code_out, n, (unsigned char *)g.carry_tiles.p);
      CK(cudaGetLastError());
      carry_apply_kernel<<<tiles, 256, 0, g.stream>>>(p.code_out, n, code0, (const unsigned char *)g.carry_tiles.p,
                                                      (unsigned char *)g.code_in.p);
      CK(cudaGetLastError());
      g.launches++;
    }
    auto kern = sick_kernel<false, ORDERED>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches += 3;
  }
  return MSIM_OK;
}

int sick_sicker_run(int64_t n, const double par17[17], int indivp, int substreams, int64_t first_person, msim_report *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!out || !par17 || n < 0 || first_person < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  SickParams p;
  memset(&p, 0, sizeof p);
  sick_sicker::Par &P = p.par;
  P.r_HS1 = par17[0]; P.r_HD = par17[1]; P.r_S1H = par17[2]; P.r_S1S2 = par17[3]; P.r_S1D = par17[4]; P.r_S2D = par17[5];
  P.c_H = par17[6]; P.c_S1 = par17[7]; P.c_S2 = par17[8]; P.c_Trt = par17[9];
  P.u_H = par17[10]; P.u_S1 = par17[11]; P.u_S2 = par17[12]; P.u_Trt = par17[13];
  const double rate = par17[14], by = par17[15];
  P.Trt = par17[16] != 0.0;
  if (!(rate >= 0)) return fail(MSIM_ERR_ARG, "discountRate must be >= 0");
  SummaryGeom G;
  // report.setPartition(0.0, 31.0, partitionBy); report.setDiscountRate(discountRate)   (README.org:294-295)
  if (!make_summary_geom(&G, sick_sicker::N_STATE, sick_sicker::N_EVENT, 0.0, 31.0, by, 1.0e100, rate, rate))
    return fail(MSIM_ERR_ARG, "partitionBy gives more age bins than the report holds");
  if ((rc = ensure(g.sgeom, sizeof G))) return rc;
  CK(cudaMemcpyAsync(g.sgeom.p, &G, sizeof G, cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));  // G is a stack object
  const int64_t n_indiv = indivp ? n : 1;
  if ((rc = ensure(g.dense, sizeof(double) * G.total))) return rc;
  if ((rc = ensure(g.indiv, sizeof(double) * 2 * (size_t)std::max<int64_t>(n_indiv, 1)))) return rc;
  if ((rc = ensure(g.code_in, (size_t)std::max<int64_t>(n, 1)))) return rc;
  if ((rc = ensure(g.code_out, (size_t)std::max<int64_t>(n, 1)))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * G.total, g.stream));
  CK(cudaMemsetAsync(g.indiv.p, 0, sizeof(double) * 2 * (size_t)std::max<int64_t>(n_indiv, 1), g.stream));
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  p.first_person = first_person;
  p.n = n;
  p.geom = (const SummaryGeom *)g.sgeom.p;
  p.code_in = (const unsigned char *)g.code_in.p;
  p.code_out = (unsigned char *)g.code_out.p;
  p.dense = (double *)g.dense.p;
  p.indiv = (double *)g.indiv.p;
  p.indivp = indivp ? 1 : 0;
  p.error = small_view().error;
  const size_t smem = sick_kernel_smem(G);
  const int64_t tiles = (n + BLOCK - 1) / BLOCK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)g.sm_count * 4, tiles));
  int64_t end_word = 0;
  p.tie = (int *)((double *)g.small.p + 13);
  if (n > 0 && !substreams && !g.rng.mt_seeded)
    return fail(MSIM_ERR_STATE, "Mersenne-Twister state not set: call msim_mt_set_seed or msim_mt_set_state first");
  // rng = new Rng(): takes the package seed, which then moves on one stream; person i draws from substream
  // first_person + i + 1 (the regime the oracle's `substreams` flag selects)
  HostStream hs;
  if (n > 0 && substreams) hs = g.rng.new_stream();
  // Schedule 0 keeps the live actions in a small ordered array (engine.cuh, OrderedEvents; cancelled actions are
  // removed), which pops in the heap's order unless two pending actions compare equal; a run that saw such a pair is
  // repeated with heap.h's binary heap and lingering cancelled entries (schedule 1).  Schedule 2 (tests) takes the
  // repeat path unconditionally.
  for (int attempt = g.person_kernel == 1 ? 1 : 0; attempt < 2; attempt++) {
    CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * G.total, g.stream));
    CK(cudaMemsetAsync(g.indiv.p, 0, sizeof(double) * 2 * (size_t)std::max<int64_t>(n_indiv, 1), g.stream));
    CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
    CK(cudaEventRecord(g.ev0, g.stream));
    if (n > 0) {
      rc = attempt == 0 ? sick_launch<true>(p, n, substreams != 0, hs, first_person, grid, smem, &end_word)
                        : sick_launch<false>(p, n, substreams != 0, hs, first_person, grid, smem, &end_word);
      if (rc) return rc;
    }
    CK(cudaEventRecord(g.ev1, g.stream));
    g.timing_pending = true;
    g.have_last = false;
    if ((rc = finish_run(nullptr))) return rc;
    int tie = 0;
    CK(cudaMemcpy(&tie, p.tie, sizeof tie, cudaMemcpyDeviceToHost));
    if (attempt == 0 && !tie && g.person_kernel != 2) break;
  }
  std::vector<double> d(G.total), ind(2 * (size_t)std::max<int64_t>(n_indiv, 1));
  CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ind.data(), g.indiv.p, sizeof(double) * ind.size(), cudaMemcpyDeviceToHost));
  dense_to_summary_impl(d.data(), G, ind.data(), n_indiv, out);
  if (n > 0 && !substreams && end_word > 0) {  // leave R's generator after the last person's last draw
    int64_t batch = (end_word - 1) / 624;
    int64_t mti = end_word - batch * 624;
    CK(cudaMemcpy(&g.rng.mt[1], (const uint32_t *)g.mt_words.p + batch * 624, sizeof(uint32_t) * 624, cudaMemcpyDeviceToHost));
    g.rng.mt[0] = (uint32_t)mti;
  }
  return MSIM_OK;
}

// ---- cancer management (SummaryReport + addPointCost)

int cancer_management_run(int64_t n, double odShape, double rate, double by, int indivp, int64_t first_person, msim_report *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!out || n < 0 || first_person < 0 || !(odShape > 0)) return fail(MSIM_ERR_ARG, "bad arguments");
  if (!(rate >= 0)) return fail(MSIM_ERR_ARG, "discountRate must be >= 0");
  SummaryGeom G;
  // report.setPartition(0.0, 100.0, partitionBy); report.setDiscountRate(discountRate)   (test_microsimulation.R:1029-1030)
  if (!make_summary_geom(&G, cancer_mgmt::N_STATE, cancer_mgmt::N_EVENT, 0.0, 100.0, by, 1.0e100, rate, rate))
    return fail(MSIM_ERR_ARG, "partitionBy gives more age bins than the report holds");
  if ((rc = ensure(g.sgeom, sizeof G))) return rc;
  CK(cudaMemcpyAsync(g.sgeom.p, &G, sizeof G, cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));  // G is a stack object
  const int64_t n_indiv = indivp ? n : 1;
  if ((rc = ensure(g.dense, sizeof(double) * G.total))) return rc;
  if ((rc = ensure(g.indiv, sizeof(double) * 2 * (size_t)std::max<int64_t>(n_indiv, 1)))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * G.total, g.stream));
  CK(cudaMemsetAsync(g.indiv.p, 0, sizeof(double) * 2 * (size_t)std::max<int64_t>(n_indiv, 1), g.stream));
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  MgmtParams p;
  memset(&p, 0, sizeof p);
  p.first_person = first_person;
  p.n = n;
  p.par.odShape = odShape;
  p.geom = (const SummaryGeom *)g.sgeom.p;
  p.dense = (double *)g.dense.p;
  p.indiv = (double *)g.indiv.p;
  p.indivp = indivp ? 1 : 0;
  p.error = small_view().error;
  const size_t smem = sick_kernel_smem(G);
  const int64_t tiles = (n + BLOCK - 1) / BLOCK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)g.sm_count * 4, tiles));
  CK(cudaEventRecord(g.ev0, g.stream));
  if (n > 0) {
    // rng = new Rng(): takes the package seed, which then moves on one stream; person i draws from substream first_person + i + 1
    HostStream hs = g.rng.new_stream();
    stream_start(hs.Cg.s, first_person, (int64_t)grid * BLOCK, &p.st);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(mgmt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mgmt_kernel<<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches++;
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  g.have_last = false;
  if ((rc = finish_run(nullptr))) return rc;
  std::vector<double> d(G.total), ind(2 * (size_t)std::max<int64_t>(n_indiv, 1));
  CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ind.data(), g.indiv.p, sizeof(double) * ind.size(), cudaMemcpyDeviceToHost));
  dense_to_summary_impl(d.data(), G, ind.data(), n_indiv, out);
  return MSIM_OK;
}

// ---- CostReport: a batch of point costs

int cost_report_run(double rate, double start_age, double p_start, double p_finish, double p_delta, int32_t n_state, int64_t n,
                    const int32_t *state, const double *time, const double *cost, msim_table *out, double *current) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!out || !current || n < 0 || n_state < 1 || n_state > 64 || (n > 0 && (!state || !time || !cost)) || !(rate >= 0))
    return fail(MSIM_ERR_ARG, "bad arguments");
  SummaryGeom G;
  if (!make_summary_geom(&G, n_state, 1, p_start, p_finish, p_delta, 1.0e100, 0.0, rate))
    return fail(MSIM_ERR_ARG, "the partition has more age bins than the report holds");
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  if ((rc = ensure(g.sgeom, sizeof G))) return rc;
  if ((rc = ensure(g.dense, sizeof(double) * (G.total + 1)))) return rc;
  if ((rc = ensure(g.gsm_in, sizeof(double) * 2 * nn))) return rc;
  if ((rc = ensure(g.aux, sizeof(int) * nn))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemcpyAsync(g.sgeom.p, &G, sizeof G, cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * (G.total + 1), g.stream));
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  double *d_time = (double *)g.gsm_in.p, *d_cost = d_time + nn;
  if (n > 0) {
    CK(cudaMemcpyAsync(d_time, time, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(d_cost, cost, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.aux.p, state, sizeof(int) * n, cudaMemcpyHostToDevice, g.stream));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)g.sm_count * 8));
    cost_report_kernel<<<grid, 256, 0, g.stream>>>((const SummaryGeom *)g.sgeom.p, n, (const int *)g.aux.p, d_time, d_cost, rate,
                                                   start_age, (double *)g.dense.p, (double *)g.dense.p + G.total, small_view().error);
    CK(cudaGetLastError());
    g.launches++;
  }
  g.have_last = false;
  if ((rc = finish_run(nullptr))) return rc;
  std::vector<double> d(G.total + 1);
  CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
  *current = d[G.total];
  std::vector<double> rows;  // CostReport::wrap(): (Key, age, cost) for every cell a cost was booked to, key-sorted
  for (int s = 0; s < G.n_state; s++)
    for (int b = 0; b < G.n_bin; b++)
      if (d[G.off_pcn + s * G.n_bin + b] > 0) rows.insert(rows.end(), {(double)s, G.part[b], d[G.off_cost + s * G.n_bin + b]});
  table_from_rows(rows, 3, COST_COLS, out);
  return MSIM_OK;
}

}  // namespace msim_host
