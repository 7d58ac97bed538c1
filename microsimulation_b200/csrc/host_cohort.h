// Host side of the generic cohort kernel (kernels.cuh): launch, the row-log expand pass, result fetches, and the
// sequential-stream pipeline (Mersenne-Twister stream -> draws per position -> chain resolution -> persons).
#pragma once
#include "host_runtime.h"

namespace msim_host {
using namespace msim;

// ---------------------------------------------------------------- cohort launch

template <class Traits, bool STREAM>
inline int cohort_grid(size_t smem, int64_t n, int *grid) {
  auto kern = cohort_kernel<Traits, STREAM>;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
  if (per_sm < 1) return fail(MSIM_ERR_CUDA, "cohort kernel does not fit on an SM");
  int64_t tiles = (n + BLOCK - 1) / BLOCK;
  int64_t gmax = (int64_t)per_sm * g.sm_count;
  if (gmax > JUMP_TAB * (int64_t)JUMP_TAB / BLOCK) gmax = JUMP_TAB * (int64_t)JUMP_TAB / BLOCK;
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>(gmax, tiles));
  return MSIM_OK;
}

// Allocates and zeroes the accumulators a run needs and fills the pointers in p.
template <class Traits, bool STREAM>
inline int cohort_prepare(CohortParams<typename Traits::Consts> &p, const uint32_t stream_state[6], int *grid_out,
                   size_t *smem_out) {
  int rc;
  size_t smem = cohort_kernel_smem<Traits>(p.geom, p.outputs);
  int grid = 1;
  if ((rc = cohort_grid<Traits, STREAM>(smem, p.n, &grid))) return rc;
  if (!STREAM) stream_start(stream_state, p.first_person, (int64_t)grid * BLOCK, &p.st);
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  SmallView sv = small_view();
  p.error = sv.error;
  // one accumulator vector: the dense report cells (if requested) followed by the N_COUNTS count cells,
  // so that a single all-reduce merges everything a shard produces
  const int cells = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  if ((rc = ensure(g.dense, sizeof(double) * (cells + N_COUNTS)))) return rc;
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * (cells + N_COUNTS), g.stream));
  p.dense = (double *)g.dense.p;
  p.counts = p.dense + cells;
  g.last_cells = cells;
  if (p.outputs & OUT_ROWLOG) {
    const int64_t wtiles = std::max<int64_t>((p.n + 31) / 32, 1);
    const size_t slots = (size_t)wtiles * 32 * Traits::MAX_ROWS;
    if ((rc = ensure(g.prov_start, sizeof(double) * slots))) return rc;
    if ((rc = ensure(g.prov_end, sizeof(double) * slots))) return rc;
    if ((rc = ensure(g.prov_meta, sizeof(int) * slots))) return rc;
    // the kernel's last block tile may be partial: its idle warps still store a 0 count
    if ((rc = ensure(g.tile_count, sizeof(int) * (size_t)(wtiles + WARPS)))) return rc;
    if ((rc = ensure(g.partial, sizeof(unsigned long long) * (size_t)((wtiles + EXPAND_TILES - 1) / EXPAND_TILES + 1)))) return rc;
    if ((rc = ensure(g.rowlog, sizeof(double) * 5 * (size_t)std::max<int64_t>(g.want_capacity, 1)))) return rc;
    p.prov_start = (double *)g.prov_start.p;
    p.prov_end = (double *)g.prov_end.p;
    p.prov_meta = (int *)g.prov_meta.p;
    p.tile_count = (int *)g.tile_count.p;
  }
  *grid_out = grid;
  *smem_out = smem;
  return MSIM_OK;
}

// second pass of the row log: tile counts -> offsets -> final columns
inline int rowlog_expand(int64_t n, int64_t first_person, int max_rows, int64_t capacity) {
  ExpandParams e;
  memset(&e, 0, sizeof e);
  e.prov_start = (const double *)g.prov_start.p;
  e.prov_end = (const double *)g.prov_end.p;
  e.prov_meta = (const int *)g.prov_meta.p;
  e.tile_count = (const int *)g.tile_count.p;
  e.n_warp_tiles = (n + 31) / 32;
  e.tile_cap = 32 * max_rows;
  e.partial = (unsigned long long *)g.partial.p;
  e.n_rows = small_view().n_rows;
  e.n_chunks = (e.n_warp_tiles + EXPAND_TILES - 1) / EXPAND_TILES;
  e.first_person = first_person;
  e.rowlog = (double *)g.rowlog.p;
  e.capacity = capacity;
  if (e.n_chunks == 0) return MSIM_OK;
  rowlog_partial_kernel<<<(unsigned)e.n_chunks, 256, 0, g.stream>>>(e);
  CK(cudaGetLastError());
  rowlog_chunk_scan_kernel<<<1, 1024, 0, g.stream>>>(e);
  CK(cudaGetLastError());
  rowlog_expand_kernel<<<(unsigned)e.n_chunks, 256, 0, g.stream>>>(e);
  CK(cudaGetLastError());
  g.launches += 3;
  return MSIM_OK;
}

template <class Traits, bool STREAM>
inline int cohort_fire(CohortParams<typename Traits::Consts> &p, int grid, size_t smem, int64_t capacity) {
  if (p.n > 0) {
    cohort_kernel<Traits, STREAM><<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches++;
    if (p.outputs & OUT_ROWLOG) {
      int rc = rowlog_expand(p.n, p.first_person, Traits::MAX_ROWS, capacity);
      if (rc) return rc;
    }
  }
  g.last_geom = p.geom;
  g.last_outputs = p.outputs;
  g.last_capacity = capacity;
  g.last_n = p.n;
  g.last_first = p.first_person;
  g.last_max_rows = Traits::MAX_ROWS;
  g.last_staged = false;
  g.have_last = true;
  return MSIM_OK;
}

// The staged person-r launch left the times at which its two waves of blocks were done (StagedParams::wave_clock):
// move the first wave's share of the tiles towards the value at which both finish together.  Measured on a B200
// (report + counts, 1e8 persons): at equal shares the first wave is done 14% before the second, and a share of 0.55
// evens them out — 0.36 of the relative gap, or 0.72 of it times the share; 0.6 is applied, so the share comes up to
// its value from one side (0.500, 0.542, 0.548, ... instead of overshooting).
constexpr int WAVE_MIN_TILES = 16;    // tiles per warp below which a launch is not split into waves
constexpr int WAVE_ADAPT_TILES = 64;  // ... below which its clocks say more about the launch's fixed costs than about the share
inline void wave_feedback(const double *small) {
  if (g.last_wave_slot < 0 || !g.last_wave_adapt) return;
  const unsigned long long *c = (const unsigned long long *)small;
  const unsigned long long start = ~c[0];
  if (c[0] == 0 || c[1] <= start || c[2] <= start) return;
  const double t1 = (double)(c[1] - start), t2 = (double)(c[2] - start), f = g.last_wave_share;
  double df = t1 < t2 ? 0.6 * f * (t2 - t1) / t1 : -0.6 * (1.0 - f) * (t1 - t2) / t2;
  if (fabs(t2 - t1) < 0.004 * std::max(t1, t2)) df = 0;
  g.wave_share[g.last_wave_slot] = std::min(0.65, std::max(0.40, f + df));
  g.last_wave_slot = -1;  // one adjustment per launch
}

inline int finish_run(int64_t *n_rows_out) {
  CK(cudaStreamSynchronize(g.stream));
  if (g.timing_pending) {
    CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
    g.timing_pending = false;
  }
  double *small = g.last_small;  // kept: callers read their own flags from it (e.g. the calibration sweep's tie flag)
  CK(cudaMemcpy(small, g.small.p, sizeof g.last_small, cudaMemcpyDeviceToHost));
  wave_feedback(small);
  int err = *(int *)(small + 10);
  unsigned long long nr = *(unsigned long long *)(small + 9);
  g.last_n_rows = (int64_t)nr;
  g.last_ov_rows = (int64_t) * (unsigned long long *)(small + 12);
  if (n_rows_out) *n_rows_out = (int64_t)nr;
  if (err == MSIM_ERR_HEAP) return fail(MSIM_ERR_HEAP, "a person scheduled more pending events / rows than the device capacity");
  if (err == MSIM_ERR_STREAM) return fail(MSIM_ERR_STREAM, "a person drew more uniforms than the sequential-stream plan allows");
  if (err == MSIM_ERR_MODEL) return fail(MSIM_ERR_MODEL, "a model reported a (state, event) pair outside the set it declared reachable");
  if (err == MSIM_ERR_PEER) return fail(MSIM_ERR_PEER, "peer merge: another rank did not deliver its result vector in time");
  if (err) return fail(err, "device-side error");
  return MSIM_OK;
}

// The staged person-r schedule keeps rows beyond a person's first in a region per chunk of persons; a chunk that
// logged more than the region holds has dropped rows.  msim_person_run_device repeats such a run with the exact
// size; the asynchronous entries (msim_person_enqueue + msim_sync / msim_person_fetch_rowlog) report it instead.
inline int rowlog_overflow_check() {
  if (g.have_last && g.last_staged && (g.last_outputs & OUT_ROWLOG) && g.last_ov_rows > g.last_ov_cap)
    return fail(MSIM_ERR_HEAP, "row log: a chunk of persons logged more rows beyond their first than planned for "
                               "(msim_set_rowlog_plan); use msim_person_run_device, which repeats such a run with the exact size");
  return MSIM_OK;
}

inline int64_t rowlog_capacity_guess(int64_t n, int max_rows, double mean_rows) {
  if (n <= 65536 && mean_rows >= 1.0) return n * max_rows + 1;
  int64_t c = (int64_t)((double)n * mean_rows) + (mean_rows >= 1.0 ? 65536 : 16);
  return std::min<int64_t>(c, n * max_rows + 1);
}

inline int fetch_rowlog(const char *const *cols, msim_table *out) {
  int64_t n = g.last_n_rows, cap = g.last_capacity;
  out->nrow = n;
  out->ncol = 5;
  out->colnames = cols;
  out->data = nullptr;
  int rc;
  if ((rc = pinned_alloc(sizeof(double) * 5 * (size_t)std::max<int64_t>(n, 1), &out->data))) return rc;
  if (n > 0) {
    CK(cudaMemcpy2DAsync(out->data, sizeof(double) * n, g.rowlog.p, sizeof(double) * cap, sizeof(double) * n, 5,
                         cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
  }
  return MSIM_OK;
}

inline int fetch_report(msim_report *out) {
  std::vector<double> d(g.last_geom.total);
  CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
  dense_to_report_impl(d.data(), g.last_geom, out);
  return MSIM_OK;
}

// the guessed row capacity was too small: the packed rows are still there, expand again
int staged_rowpasses(int64_t n, int64_t first_person, int64_t capacity, bool count_first);  // host_person.h

inline int reexpand(int64_t rows) {
  int rc;
  if ((rc = ensure(g.rowlog, sizeof(double) * 5 * (size_t)rows))) return rc;
  if (g.last_staged) {
    if ((rc = staged_rowpasses(g.last_n, g.last_first, rows, false))) return rc;
  } else if ((rc = rowlog_expand(g.last_n, g.last_first, g.last_max_rows, rows))) {
    return rc;
  }
  g.last_capacity = rows;
  return finish_run(nullptr);
}


// ---- sequential-stream models

template <class Traits>
inline int sequential_run(int64_t n, const typename Traits::Consts &consts, int max_draws, unsigned outputs,
                   const ReportGeom &geom, int max_rows, double mean_rows) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n < 0) return fail(MSIM_ERR_ARG, "n < 0");
  if (!g.rng.mt_seeded)
    return fail(MSIM_ERR_STATE, "Mersenne-Twister state not set: call msim_mt_set_seed or msim_mt_set_state first");
  if (max_draws > CHAIN_MAX_DRAWS) return fail(MSIM_ERR_ARG, "max_draws too large");
  const int64_t mti0 = g.rng.mt[0];  // next word to hand out, 1..624 (624 = regenerate first)
  const int64_t n_pos = (int64_t)max_draws * n;
  const int64_t words_needed = mti0 + n_pos + max_draws + 1;
  const int64_t n_batches = (words_needed + 623) / 624;  // including batch 0 = current state
  const int64_t mt_end = n_batches * 624;
  if ((rc = ensure(g.mt_words, sizeof(uint32_t) * (size_t)mt_end))) return rc;
  CK(cudaMemcpyAsync(g.mt_words.p, &g.rng.mt[1], sizeof(uint32_t) * 624, cudaMemcpyHostToDevice, g.stream));
  CK(cudaEventRecord(g.ev0, g.stream));
  if (n_batches > 1) {
    if ((rc = mt_generate((uint32_t *)g.mt_words.p, mt_end))) return rc;
  }
  CohortParams<typename Traits::Consts> p;
  memset(&p, 0, sizeof p);
  g.last_simple_names = true;
  if (n > 0) {
    const int64_t n_chunks = (n_pos + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
    if ((rc = ensure(g.len, (size_t)n_pos))) return rc;
    if ((rc = ensure(g.starts, sizeof(int64_t) * (size_t)n))) return rc;
    if ((rc = ensure(g.chain_exit, (size_t)n_chunks * CHAIN_MAX_DRAWS))) return rc;
    if ((rc = ensure(g.chain_count, sizeof(int) * (size_t)n_chunks * CHAIN_MAX_DRAWS))) return rc;
    if ((rc = ensure(g.chain_entry, (size_t)n_chunks))) return rc;
    if ((rc = ensure(g.chain_base, sizeof(int64_t) * (size_t)n_chunks))) return rc;
    if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
    int lgrid = (int)std::min<int64_t>((n_pos + BLOCK - 1) / BLOCK, (int64_t)g.sm_count * 8);
    seq_len_kernel<Traits><<<lgrid, BLOCK, 0, g.stream>>>((const uint32_t *)g.mt_words.p, mt_end, mti0, n_pos, consts,
                                                         (unsigned char *)g.len.p, nullptr);
    CK(cudaGetLastError());
    chain_chunk_kernel<<<(unsigned)n_chunks, 64, 0, g.stream>>>((const unsigned char *)g.len.p, n_pos, max_draws, CHAIN_MAX_DRAWS,
                                                                (unsigned char *)g.chain_exit.p, (int *)g.chain_count.p);
    CK(cudaGetLastError());
    chain_scan_kernel<<<1, 256, 0, g.stream>>>((const unsigned char *)g.chain_exit.p, (const int *)g.chain_count.p,
                                               n_chunks, n, CHAIN_MAX_DRAWS, (unsigned char *)g.chain_entry.p, (int64_t *)g.chain_base.p);
    CK(cudaGetLastError());
    g.launches += 3;
  }
  p.first_person = 0;
  p.n = n;
  p.outputs = outputs;
  p.consts = consts;
  p.geom = geom;
  const int64_t capacity = rowlog_capacity_guess(n, max_rows, mean_rows);
  g.want_capacity = capacity;
  p.mt_words = (const uint32_t *)g.mt_words.p;
  p.mt_end = mt_end;
  p.starts = (const int64_t *)g.starts.p;
  int grid;
  size_t smem;
  uint32_t unused_state[6] = {0, 0, 0, 0, 0, 0};
  if ((rc = cohort_prepare<Traits, true>(p, unused_state, &grid, &smem))) return rc;  // zeroes `small`
  if (n > 0) {
    const int64_t n_chunks = (n_pos + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
    SmallView sv = small_view();
    chain_emit_kernel<<<(unsigned)n_chunks, 64, 0, g.stream>>>(
        (const unsigned char *)g.len.p, n_pos, mti0, (const unsigned char *)g.chain_entry.p,
        (const int64_t *)g.chain_base.p, n, (int64_t *)g.starts.p, sv.end_word, sv.error);
    CK(cudaGetLastError());
    g.launches++;
  }
  if ((rc = cohort_fire<Traits, true>(p, grid, smem, capacity))) return rc;
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  int64_t rows = 0;
  if ((rc = finish_run(&rows))) return rc;
  int64_t end_word = 0;
  CK(cudaMemcpy(&end_word, small_view().end_word, sizeof end_word, cudaMemcpyDeviceToHost));
  if ((outputs & OUT_ROWLOG) && rows > g.last_capacity && (rc = reexpand(rows))) return rc;
  // leave R's generator where the reference would: after the last person's last draw
  if (n > 0 && end_word > 0) {
    int64_t batch = (end_word - 1) / 624;
    int64_t mti = end_word - batch * 624;
    CK(cudaMemcpy(&g.rng.mt[1], (const uint32_t *)g.mt_words.p + batch * 624, sizeof(uint32_t) * 624,
                  cudaMemcpyDeviceToHost));
    g.rng.mt[0] = (uint32_t)mti;
  }
  return MSIM_OK;
}

// Stream regime for models that may draw many uniforms per person (up to 250; chunk tables with a stride of 256
// instead of sequential_run's 8).  One attempt with `n_pos` start positions planned for; *short_plan is set when the chain
// of persons ran off the planned positions before person n-1 (the caller retries with the maximum).
template <class Traits>
inline int stream_offsets_attempt(int64_t n, const typename Traits::Consts &consts, int64_t n_pos, int max_draws, bool want_aux,
                           bool *short_plan, int64_t *end_word_out, int64_t *mt_end_out) {
  int rc;
  *short_plan = false;
  const int64_t mti0 = g.rng.mt[0];
  const int64_t words_needed = mti0 + n_pos + max_draws + 1;
  const int64_t n_batches = (words_needed + 623) / 624;
  const int64_t mt_end = n_batches * 624;
  *mt_end_out = mt_end;
  if ((rc = ensure(g.mt_words, sizeof(uint32_t) * (size_t)mt_end))) return rc;
  CK(cudaMemcpyAsync(g.mt_words.p, &g.rng.mt[1], sizeof(uint32_t) * 624, cudaMemcpyHostToDevice, g.stream));
  if (n_batches > 1) {
    if ((rc = mt_generate((uint32_t *)g.mt_words.p, mt_end))) return rc;
  }
  const int stride = 256;
  const int64_t n_chunks = (n_pos + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
  if ((rc = ensure(g.len, (size_t)n_pos))) return rc;
  if (want_aux && (rc = ensure(g.aux, (size_t)n_pos))) return rc;
  if ((rc = ensure(g.starts, sizeof(int64_t) * (size_t)n))) return rc;
  if ((rc = ensure(g.chain_exit, (size_t)n_chunks * stride))) return rc;
  if ((rc = ensure(g.chain_count, sizeof(int) * (size_t)n_chunks * stride))) return rc;
  if ((rc = ensure(g.chain_entry, (size_t)n_chunks))) return rc;
  if ((rc = ensure(g.chain_base, sizeof(int64_t) * (size_t)n_chunks))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  SmallView sv = small_view();
  int lgrid = (int)std::min<int64_t>((n_pos + BLOCK - 1) / BLOCK, (int64_t)g.sm_count * 8);
  seq_len_kernel<Traits><<<lgrid, BLOCK, 0, g.stream>>>((const uint32_t *)g.mt_words.p, mt_end, mti0, n_pos, consts,
                                                        (unsigned char *)g.len.p, want_aux ? (unsigned char *)g.aux.p : nullptr);
  CK(cudaGetLastError());
  chain_chunk_kernel<<<(unsigned)n_chunks, 256, 0, g.stream>>>((const unsigned char *)g.len.p, n_pos, max_draws, stride,
                                                               (unsigned char *)g.chain_exit.p, (int *)g.chain_count.p);
  CK(cudaGetLastError());
  if (n_chunks > 4 * CHAIN_GROUP) {
    // groups of chunks composed in parallel, the scan over the groups, then the chunks of every group
    const int64_t n_groups = (n_chunks + CHAIN_GROUP - 1) / CHAIN_GROUP;
    if ((rc = ensure(g.chain_gexit, (size_t)n_groups * stride))) return rc;
    if ((rc = ensure(g.chain_gcount, sizeof(int) * (size_t)n_groups * stride))) return rc;
    if ((rc = ensure(g.chain_gentry, (size_t)n_groups))) return rc;
    if ((rc = ensure(g.chain_gbase, sizeof(int64_t) * (size_t)n_groups))) return rc;
    chain_group_kernel<<<(unsigned)n_groups, 256, 0, g.stream>>>((const unsigned char *)g.chain_exit.p, (const int *)g.chain_count.p,
                                                                 n_chunks, stride, (unsigned char *)g.chain_gexit.p, (int *)g.chain_gcount.p);
    CK(cudaGetLastError());
    chain_scan_kernel<<<1, 256, 0, g.stream>>>((const unsigned char *)g.chain_gexit.p, (const int *)g.chain_gcount.p, n_groups, n,
                                               stride, (unsigned char *)g.chain_gentry.p, (int64_t *)g.chain_gbase.p);
    CK(cudaGetLastError());
    chain_expand_kernel<<<(unsigned)((n_groups + 63) / 64), 64, 0, g.stream>>>(
        (const unsigned char *)g.chain_exit.p, (const int *)g.chain_count.p, n_chunks, n_groups, n, stride,
        (const unsigned char *)g.chain_gentry.p, (const int64_t *)g.chain_gbase.p, (unsigned char *)g.chain_entry.p,
        (int64_t *)g.chain_base.p);
    CK(cudaGetLastError());
    g.launches += 2;
  } else {
    chain_scan_kernel<<<1, 256, 0, g.stream>>>((const unsigned char *)g.chain_exit.p, (const int *)g.chain_count.p, n_chunks, n,
                                               stride, (unsigned char *)g.chain_entry.p, (int64_t *)g.chain_base.p);
    CK(cudaGetLastError());
  }
  chain_emit_kernel<<<(unsigned)n_chunks, 64, 0, g.stream>>>((const unsigned char *)g.len.p, n_pos, mti0,
                                                             (const unsigned char *)g.chain_entry.p, (const int64_t *)g.chain_base.p,
                                                             n, (int64_t *)g.starts.p, sv.end_word, sv.error);
  CK(cudaGetLastError());
  g.launches += 4;
  CK(cudaStreamSynchronize(g.stream));
  double *small = g.last_small;  // kept: callers read their own flags from it (e.g. the calibration sweep's tie flag)
  CK(cudaMemcpy(small, g.small.p, sizeof g.last_small, cudaMemcpyDeviceToHost));
  const int64_t end_word = *(int64_t *)(small + 11);
  const int err = *(int *)(small + 10);
  if (end_word == 0 || err == MSIM_ERR_STREAM) {  // person n-1 was never reached
    *short_plan = true;
    return MSIM_OK;
  }
  if (err) return fail(err, "device-side error while resolving stream offsets");
  *end_word_out = end_word;
  return MSIM_OK;
}


}  // namespace msim_host
