// C ABI of the B200 engine (include/msim.h): host-side orchestration in C++,
// every simulation step on the GPU.  There is no CPU execution path: without
// a CUDA device each simulation entry returns MSIM_ERR_NO_DEVICE.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/msim.h"
#include "host_state.h"
#include "report_host.h"
#include "kernels.cuh"
#include "person_staged.cuh"
#include "summary_kernels.cuh"

using namespace msim;

namespace {

// ---------------------------------------------------------------- context

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct Ctx {
  bool inited = false;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  MrgJump *d_lo = nullptr, *d_hi = nullptr;
  DevBuf rowlog, prov_start, prov_end, prov_meta, tile_count, partial, dense, small, mt_words, len, starts, chain_exit, chain_count, chain_entry, chain_base, calib_out,
      calib_par, uniforms, first_end, first_tag, ov_fill, ov_tag, ov_start, ov_end, aux, code_in, code_out, sgeom,
      indiv, gsm_model, gsm_data, gsm_in, gsm_out, slices;
  int person_kernel = 0;  // 0 = staged schedule (person_staged.cuh), 1 = generic cohort kernel
  bool last_staged = false;
  double plan_rows = 1.25, plan_extra_rows = 0.26;  // row-log buffer planning (msim_set_rowlog_plan)
  int64_t last_ov_cap = 0, last_ov_rows = 0;
  HostRngState rng;
  // last cohort run
  int64_t last_n_rows = 0, last_capacity = 0;
  ReportGeom last_geom;
  unsigned last_outputs = 0;
  bool last_simple_names = false;
  int64_t want_capacity = 0;
  int64_t last_n = 0, last_first = 0;
  int last_max_rows = 0;
  int last_cells = 0;  // dense report cells of the last run (0 when no report was requested)
  bool have_last = false;
  float last_ms = 0.f;
  bool timing_pending = false;
  int64_t launches = 0;
  // pinned host cache (one buffer) so steady-state row-log fetches do not re-pin
  void *pin_cache = nullptr;
  size_t pin_cache_bytes = 0;
  std::unordered_map<void *, size_t> pinned;  // live pinned allocations -> bytes
};

Ctx g;
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(MSIM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
  } while (0)

int ensure(DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return MSIM_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes < 256 ? 256 : bytes;
  CK(cudaMalloc(&b.p, want));
  b.bytes = want;
  return MSIM_OK;
}

int ensure_init() {
  if (g.inited) return MSIM_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(MSIM_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
  if (g.device >= ndev) return fail(MSIM_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(g.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, g.device));
  g.sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
  g.stream = g.own_stream;
  CK(cudaEventCreate(&g.ev0));
  CK(cudaEventCreate(&g.ev1));
  // per-thread start tables: lo[j] = M^j, hi[j] = M^(JUMP_TAB*j)
  std::vector<MrgJump> lo(JUMP_TAB), hi(JUMP_TAB);
  const MrgJump &M = mrg_substream_jump();
  lo[0] = mrg_jump_identity();
  for (int j = 1; j < JUMP_TAB; j++) lo[j] = mrg_jump_mul(M, lo[j - 1]);
  MrgJump Mt = mrg_jump_mul(M, lo[JUMP_TAB - 1]);  // M^JUMP_TAB
  hi[0] = mrg_jump_identity();
  for (int j = 1; j < JUMP_TAB; j++) hi[j] = mrg_jump_mul(Mt, hi[j - 1]);
  CK(cudaMalloc(&g.d_lo, sizeof(MrgJump) * JUMP_TAB));
  CK(cudaMalloc(&g.d_hi, sizeof(MrgJump) * JUMP_TAB));
  CK(cudaMemcpy(g.d_lo, lo.data(), sizeof(MrgJump) * JUMP_TAB, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_hi, hi.data(), sizeof(MrgJump) * JUMP_TAB, cudaMemcpyHostToDevice));
  g.inited = true;
  return MSIM_OK;
}

// substream state of person `first_person` of the stream that starts at `state`
void stream_start(const uint32_t state[6], int64_t first_person, int64_t n_threads, StreamStart *st) {
  Mrg32k3a b;
  memcpy(b.s, state, sizeof b.s);
  b.jump(mrg_jump_pow(mrg_substream_jump(), (uint64_t)first_person + 1));
  memcpy(st->base, b.s, sizeof b.s);
  st->lo = g.d_lo;
  st->hi = g.d_hi;
  st->stride = mrg_jump_pow(mrg_substream_jump(), (uint64_t)n_threads);
}

// ---------------------------------------------------------------- host memory

int pinned_alloc(size_t bytes, double **out) {
  if (bytes == 0) bytes = 8;
  if (g.pin_cache && g.pin_cache_bytes >= bytes) {
    *out = (double *)g.pin_cache;
    g.pinned[g.pin_cache] = g.pin_cache_bytes;
    g.pin_cache = nullptr;
    g.pin_cache_bytes = 0;
    return MSIM_OK;
  }
  void *p = nullptr;
  CK(cudaMallocHost(&p, bytes));
  g.pinned[p] = bytes;
  *out = (double *)p;
  return MSIM_OK;
}

void host_free(double *p) {
  if (!p) return;
  auto it = g.pinned.find(p);
  if (it == g.pinned.end()) {
    free(p);
    return;
  }
  size_t bytes = it->second;
  g.pinned.erase(it);
  if (bytes > g.pin_cache_bytes) {  // keep the largest buffer for the next call
    if (g.pin_cache) cudaFreeHost(g.pin_cache);
    g.pin_cache = p;
    g.pin_cache_bytes = bytes;
  } else {
    cudaFreeHost(p);
  }
}

ReportGeom geom_from_dims(const msim_dense_dims *dims, double discount_rate) {
  return ages_geom(dims->n_state, dims->n_event, discount_rate);
}

// ---------------------------------------------------------------- cohort launch

// `small` buffer layout (8-byte cells): [9] n_rows (u64), [10] error (int), [11] end_word (i64), [12] overflow rows (u64)
struct SmallView {
  unsigned long long *n_rows;
  int *error;
  int64_t *end_word;
  unsigned long long *ov_count;
};
SmallView small_view() {
  SmallView v;
  double *b = (double *)g.small.p;
  v.n_rows = (unsigned long long *)(b + 9);
  v.error = (int *)(b + 10);
  v.end_word = (int64_t *)(b + 11);
  v.ov_count = (unsigned long long *)(b + 12);
  return v;
}

template <class Traits, bool STREAM>
int cohort_grid(size_t smem, int64_t n, int *grid) {
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
int cohort_prepare(CohortParams<typename Traits::Consts> &p, const uint32_t stream_state[6], int *grid_out,
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
int rowlog_expand(int64_t n, int64_t first_person, int max_rows, int64_t capacity) {
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
int cohort_fire(CohortParams<typename Traits::Consts> &p, int grid, size_t smem, int64_t capacity) {
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

int finish_run(int64_t *n_rows_out) {
  CK(cudaStreamSynchronize(g.stream));
  if (g.timing_pending) {
    CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
    g.timing_pending = false;
  }
  double small[16];
  CK(cudaMemcpy(small, g.small.p, sizeof small, cudaMemcpyDeviceToHost));
  int err = *(int *)(small + 10);
  unsigned long long nr = *(unsigned long long *)(small + 9);
  g.last_n_rows = (int64_t)nr;
  g.last_ov_rows = (int64_t) * (unsigned long long *)(small + 12);
  if (n_rows_out) *n_rows_out = (int64_t)nr;
  if (err == MSIM_ERR_HEAP) return fail(MSIM_ERR_HEAP, "a person scheduled more pending events / rows than the device capacity");
  if (err == MSIM_ERR_STREAM) return fail(MSIM_ERR_STREAM, "a person drew more uniforms than the sequential-stream plan allows");
  if (err) return fail(err, "device-side error");
  return MSIM_OK;
}

int64_t rowlog_capacity_guess(int64_t n, int max_rows, double mean_rows) {
  if (n <= 65536 && mean_rows >= 1.0) return n * max_rows + 1;
  int64_t c = (int64_t)((double)n * mean_rows) + (mean_rows >= 1.0 ? 65536 : 16);
  return std::min<int64_t>(c, n * max_rows + 1);
}

int fetch_rowlog(const char *const *cols, msim_table *out) {
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

int fetch_report(msim_report *out) {
  std::vector<double> d(g.last_geom.total);
  CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
  dense_to_report_impl(d.data(), g.last_geom, out);
  return MSIM_OK;
}

// the guessed row capacity was too small: the packed rows are still there, expand again
int staged_rowpasses(int64_t n, int64_t first_person, int64_t capacity, bool count_first);

int reexpand(int64_t rows) {
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

// ---- person-r

// the bandwidth-bound passes that turn first rows + further rows into the final columns
int staged_rowpasses(int64_t n, int64_t first_person, int64_t capacity, bool count_first) {
  RowPassParams r;
  memset(&r, 0, sizeof r);
  SmallView sv = small_view();
  r.first_end = (const double *)g.first_end.p;
  r.first_tag = (const unsigned short *)g.first_tag.p;
  r.ov_tag = (const unsigned *)g.ov_tag.p;
  r.ov_start = (const double *)g.ov_start.p;
  r.ov_end = (const double *)g.ov_end.p;
  r.ov_fill = (const unsigned *)g.ov_fill.p;
  r.max_fill = sv.ov_count;
  r.ov_cap = (int)g.last_ov_cap;
  r.n = n;
  r.first_person = first_person;
  r.chunk_rows = (unsigned long long *)g.partial.p;
  r.n_rows = sv.n_rows;
  r.n_chunks = (n + ROW_CHUNK - 1) / ROW_CHUNK;
  r.rowlog = (double *)g.rowlog.p;
  r.capacity = capacity;
  if (r.n_chunks == 0) return MSIM_OK;
  if (count_first) {
    rowpass_count_kernel<<<(unsigned)r.n_chunks, 256, 0, g.stream>>>(r);
    CK(cudaGetLastError());
    rowpass_scan_kernel<<<1, 1024, 0, g.stream>>>(r);
    CK(cudaGetLastError());
    g.launches += 2;
  }
  rowpass_final_kernel<<<(unsigned)r.n_chunks, 256, 0, g.stream>>>(r);
  CK(cudaGetLastError());
  g.launches++;
  return MSIM_OK;
}

int person_enqueue_staged(const msim_person_shard *sh, int64_t capacity, int64_t ov_cap_override) {
  int rc;
  if (sh->n >= (int64_t)4294967296ll) return fail(MSIM_ERR_ARG, "a shard holds at most 2^32 - 1 persons");
  StagedParams p;
  memset(&p, 0, sizeof p);
  p.first_person = sh->first_person;
  p.n = sh->n;
  p.outputs = sh->outputs & 7u;
  p.consts = person_r::make_consts();
  p.geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate);
  size_t smem = sizeof(WarpQueues) * SWARPS;
  if (p.outputs & OUT_ROWLOG) smem += sizeof(WarpRows) * SWARPS;
  if (p.outputs & OUT_REPORT) {
    smem += (size_t)(p.geom.total - p.geom.off_dstart) * 4;  // the counters; FP sums go to per-block global slices
  }
  int per_sm = 0;
  typedef void (*StagedKernel)(const StagedParams);
  // kernels specialised for the output combinations without a report; the general one with (person_staged.cuh)
  static const StagedKernel kernels[8] = {person_staged_kernel<0>, person_staged_kernel<1>, person_staged_kernel<OUT_RUNTIME>,
                                          person_staged_kernel<OUT_RUNTIME>, person_staged_kernel<4>, person_staged_kernel<5>,
                                          person_staged_kernel<OUT_RUNTIME>, person_staged_kernel<OUT_RUNTIME>};
  const StagedKernel kernel = kernels[p.outputs & 7u];
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SBLOCK, smem));
  if (per_sm < 1) return fail(MSIM_ERR_CUDA, "staged kernel does not fit on an SM");
  const int64_t tiles = (sh->n + 31) / 32;
  int64_t gmax = (int64_t)per_sm * g.sm_count;
  if (gmax > JUMP_TAB * (int64_t)JUMP_TAB / SBLOCK) gmax = JUMP_TAB * (int64_t)JUMP_TAB / SBLOCK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(gmax, (tiles + SWARPS - 1) / SWARPS));
  uint32_t state[6];
  seed_to_u32(sh->seed, state);
  stream_start(state, sh->first_person, (int64_t)grid * SBLOCK, &p.st);
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  SmallView sv = small_view();
  p.error = sv.error;
  const int cells = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  if ((rc = ensure(g.dense, sizeof(double) * (cells + N_COUNTS)))) return rc;
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * (cells + N_COUNTS), g.stream));
  p.dense = (double *)g.dense.p;
  p.counts = p.dense + cells;
  g.last_cells = cells;
  int64_t ov_cap = 0;
  if (p.outputs & OUT_ROWLOG) {
    const size_t n1 = (size_t)std::max<int64_t>(sh->n, 1);
    // rows beyond a person's first: 0.232 per person at the reference's parameters, at most 4; each chunk of
    // ROW_CHUNK persons owns a region planned with a wide margin (a chunk that needs more repeats the run)
    const int64_t n_chunks = (sh->n + ROW_CHUNK - 1) / ROW_CHUNK;
    ov_cap = ov_cap_override > 0 ? ov_cap_override
                                 : std::min<int64_t>((int64_t)(g.plan_extra_rows * 1.6 * ROW_CHUNK) + (g.plan_extra_rows >= 0.2 ? 96 : 1),
                                                     (int64_t)ROW_CHUNK * (person_r::MAX_ROWS - 1));
    const size_t ov_n = (size_t)std::max<int64_t>(n_chunks, 1) * (size_t)ov_cap;
    if ((rc = ensure(g.first_end, sizeof(double) * n1))) return rc;
    if ((rc = ensure(g.first_tag, sizeof(unsigned short) * n1))) return rc;
    if ((rc = ensure(g.ov_fill, sizeof(unsigned) * (size_t)std::max<int64_t>(n_chunks, 1)))) return rc;
    if ((rc = ensure(g.ov_tag, sizeof(unsigned) * ov_n))) return rc;
    if ((rc = ensure(g.ov_start, sizeof(double) * ov_n))) return rc;
    if ((rc = ensure(g.ov_end, sizeof(double) * ov_n))) return rc;
    CK(cudaMemsetAsync(g.ov_fill.p, 0, sizeof(unsigned) * (size_t)std::max<int64_t>(n_chunks, 1), g.stream));
    if ((rc = ensure(g.partial, sizeof(unsigned long long) * (size_t)((sh->n + ROW_CHUNK - 1) / ROW_CHUNK + 1)))) return rc;
    if ((rc = ensure(g.rowlog, sizeof(double) * 5 * (size_t)std::max<int64_t>(capacity, 1)))) return rc;
    p.first_end = (double *)g.first_end.p;
    p.first_tag = (unsigned short *)g.first_tag.p;
    p.ov_tag = (unsigned *)g.ov_tag.p;
    p.ov_start = (double *)g.ov_start.p;
    p.ov_end = (double *)g.ov_end.p;
    p.ov_fill = (unsigned *)g.ov_fill.p;
    p.ov_cap = (int)ov_cap;
  }
  g.last_ov_cap = ov_cap;
  const int slice = p.geom.off_dstart;  // the FP sums (pt, ut)
  if (p.outputs & OUT_REPORT) {
    const size_t bytes = sizeof(double) * (size_t)grid * slice;
    if ((rc = ensure(g.slices, bytes))) return rc;
    CK(cudaMemsetAsync(g.slices.p, 0, bytes, g.stream));
    p.slices = (double *)g.slices.p;
  }
  CK(cudaEventRecord(g.ev0, g.stream));
  if (p.n > 0) {
    kernel<<<grid, SBLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches++;
    if (p.outputs & OUT_REPORT) {
      fold_slices_kernel<<<(slice + 255) / 256, 256, 0, g.stream>>>(p.slices, grid, slice, p.dense);
      CK(cudaGetLastError());
      g.launches++;
    }
    if (p.outputs & OUT_ROWLOG) {
      if ((rc = staged_rowpasses(p.n, p.first_person, capacity, true))) return rc;
    }
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  g.last_geom = p.geom;
  g.last_outputs = p.outputs;
  g.last_capacity = capacity;
  g.last_n = p.n;
  g.last_first = p.first_person;
  g.last_max_rows = person_r::MAX_ROWS;
  g.last_staged = true;
  g.have_last = true;
  return MSIM_OK;
}

int person_enqueue(const msim_person_shard *sh, int64_t capacity_override, int64_t ov_cap_override = 0) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!sh || sh->n < 0 || sh->first_person < 0) return fail(MSIM_ERR_ARG, "bad shard");
  if (!mrg_seed_ok(sh->seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  const int64_t capacity = capacity_override > 0 ? capacity_override : rowlog_capacity_guess(sh->n, person_r::MAX_ROWS, g.plan_rows);
  g.want_capacity = capacity;
  g.last_simple_names = false;
  if (g.person_kernel == 0) return person_enqueue_staged(sh, capacity, ov_cap_override);
  CohortParams<person_r::Consts> p;
  memset(&p, 0, sizeof p);
  p.first_person = sh->first_person;
  p.n = sh->n;
  p.outputs = sh->outputs & 7u;
  p.consts = person_r::make_consts();
  p.geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate);
  uint32_t state[6];
  seed_to_u32(sh->seed, state);
  int grid;
  size_t smem;
  if ((rc = cohort_prepare<PersonTraits, false>(p, state, &grid, &smem))) return rc;
  CK(cudaEventRecord(g.ev0, g.stream));
  if ((rc = cohort_fire<PersonTraits, false>(p, grid, smem, capacity))) return rc;
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  return MSIM_OK;
}

int person_run(const msim_person_shard *sh) {
  int rc;
  if ((rc = person_enqueue(sh, 0))) return rc;
  int64_t rows = 0;
  if ((rc = finish_run(&rows))) return rc;
  if ((sh->outputs & OUT_ROWLOG) && g.last_staged && g.last_ov_rows > g.last_ov_cap) {
    // rare: more rows beyond the first than planned for — run again with the exact size
    if ((rc = person_enqueue(sh, rows > g.last_capacity ? rows : 0, g.last_ov_rows))) return rc;
    if ((rc = finish_run(&rows))) return rc;
  }
  if ((sh->outputs & OUT_ROWLOG) && rows > g.last_capacity) {  // rare: redo the expand pass with the exact size
    if ((rc = reexpand(rows))) return rc;
  }
  return MSIM_OK;
}

// ---- sequential-stream models

template <class Traits>
int sequential_run(int64_t n, const typename Traits::Consts &consts, int max_draws, unsigned outputs,
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
    mt_generate_kernel<<<1, 256, 0, g.stream>>>((uint32_t *)g.mt_words.p, (int)(n_batches - 1));
    CK(cudaGetLastError());
    g.launches++;
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

// ---- Sick-Sicker (SummaryReport)

// One attempt in the stream regime with `n_pos` start positions planned for; *short_plan is set when the chain
// of persons ran off the planned positions before person n-1 (the caller retries with the maximum).
int sick_stream_attempt(int64_t n, const SickParams &base, int64_t n_pos, int max_draws, bool *short_plan, int64_t *end_word_out,
                        int64_t *mt_end_out) {
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
    mt_generate_kernel<<<1, 256, 0, g.stream>>>((uint32_t *)g.mt_words.p, (int)(n_batches - 1));
    CK(cudaGetLastError());
    g.launches++;
  }
  const int stride = 256;
  const int64_t n_chunks = (n_pos + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
  if ((rc = ensure(g.len, (size_t)n_pos))) return rc;
  if ((rc = ensure(g.aux, (size_t)n_pos))) return rc;
  if ((rc = ensure(g.starts, sizeof(int64_t) * (size_t)n))) return rc;
  if ((rc = ensure(g.chain_exit, (size_t)n_chunks * stride))) return rc;
  if ((rc = ensure(g.chain_count, sizeof(int) * (size_t)n_chunks * stride))) return rc;
  if ((rc = ensure(g.chain_entry, (size_t)n_chunks))) return rc;
  if ((rc = ensure(g.chain_base, sizeof(int64_t) * (size_t)n_chunks))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  SmallView sv = small_view();
  int lgrid = (int)std::min<int64_t>((n_pos + BLOCK - 1) / BLOCK, (int64_t)g.sm_count * 8);
  seq_len_kernel<SickTraits><<<lgrid, BLOCK, 0, g.stream>>>((const uint32_t *)g.mt_words.p, mt_end, mti0, n_pos, base.par,
                                                            (unsigned char *)g.len.p, (unsigned char *)g.aux.p);
  CK(cudaGetLastError());
  chain_chunk_kernel<<<(unsigned)n_chunks, 256, 0, g.stream>>>((const unsigned char *)g.len.p, n_pos, max_draws, stride,
                                                               (unsigned char *)g.chain_exit.p, (int *)g.chain_count.p);
  CK(cudaGetLastError());
  chain_scan_kernel<<<1, 256, 0, g.stream>>>((const unsigned char *)g.chain_exit.p, (const int *)g.chain_count.p, n_chunks, n,
                                             stride, (unsigned char *)g.chain_entry.p, (int64_t *)g.chain_base.p);
  CK(cudaGetLastError());
  chain_emit_kernel<<<(unsigned)n_chunks, 64, 0, g.stream>>>((const unsigned char *)g.len.p, n_pos, mti0,
                                                             (const unsigned char *)g.chain_entry.p, (const int64_t *)g.chain_base.p,
                                                             n, (int64_t *)g.starts.p, sv.end_word, sv.error);
  CK(cudaGetLastError());
  g.launches += 4;
  CK(cudaStreamSynchronize(g.stream));
  double small[16];
  CK(cudaMemcpy(small, g.small.p, sizeof small, cudaMemcpyDeviceToHost));
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
  CK(cudaEventRecord(g.ev0, g.stream));
  if (n > 0 && !substreams) {
    if (!g.rng.mt_seeded)
      return fail(MSIM_ERR_STATE, "Mersenne-Twister state not set: call msim_mt_set_seed or msim_mt_set_state first");
    const int max_draws = 250;
    bool short_plan = false;
    int64_t mt_end = 0;
    // 36.4 draws per person at the README's rates; plan for 48 and fall back to the per-person maximum
    int64_t n_pos = std::min<int64_t>(48 * n + 8192, (int64_t)max_draws * n);
    if ((rc = sick_stream_attempt(n, p, n_pos, max_draws, &short_plan, &end_word, &mt_end))) return rc;
    if (short_plan && n_pos < (int64_t)max_draws * n) {
      n_pos = (int64_t)max_draws * n;
      if ((rc = sick_stream_attempt(n, p, n_pos, max_draws, &short_plan, &end_word, &mt_end))) return rc;
    }
    if (short_plan) return fail(MSIM_ERR_STREAM, "a person drew more uniforms than the sequential-stream plan allows (250)");
    CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
    p.mt_words = (const uint32_t *)g.mt_words.p;
    p.mt_end = mt_end;
    p.starts = (const int64_t *)g.starts.p;
    gather_code_kernel<<<g.sm_count * 4, 256, 0, g.stream>>>((const unsigned char *)g.aux.p, p.starts, (int64_t)g.rng.mt[0], n,
                                                             p.code_out);
    CK(cudaGetLastError());
    carry_scan_kernel<<<1, 1024, 0, g.stream>>>(p.code_out, n, 0, (unsigned char *)g.code_in.p);
    CK(cudaGetLastError());
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(sick_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sick_kernel<true><<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches += 3;
  } else if (n > 0) {
    // rng = new Rng(): takes the package seed, which then moves on one stream; person i draws from substream
    // first_person + i + 1 (the regime the oracle's `substreams` flag selects)
    HostStream hs = g.rng.new_stream();
    stream_start(hs.Cg.s, first_person, (int64_t)grid * BLOCK, &p.st);
    sick_code_kernel<<<grid, BLOCK, 0, g.stream>>>(p);
    CK(cudaGetLastError());
    carry_scan_kernel<<<1, 1024, 0, g.stream>>>(p.code_out, n, 0, (unsigned char *)g.code_in.p);
    CK(cudaGetLastError());
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(sick_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sick_kernel<false><<<grid, BLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches += 3;
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  g.have_last = false;
  if ((rc = finish_run(nullptr))) return rc;
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
    T.n_interior = t.n_knots;
    T.intercept = t.intercept != 0;
    T.cure = t.cure;
    T.b0 = t.Boundary_knots[0];
    T.b1 = t.Boundary_knots[1];
    T.df = T.intercept + 3 + T.n_interior;
    T.nknots = T.n_interior + 8;
    T.ncoef = T.nknots - 4;
    for (int k = 0; k < 4; k++) {
      T.knots[k] = T.b0;
      T.knots[T.nknots - k - 1] = T.b1;
    }
    for (int k = 0; k < T.n_interior; k++) T.knots[k + 4] = t.knots[k];
    // q_matrix = q_const, transposed if it has fewer columns than rows (splines.cpp:197-198)
    const bool tr = t.q_cols < t.q_rows;
    const int rows = tr ? t.q_cols : t.q_rows, cols = tr ? t.q_rows : t.q_cols;
    if (rows != t.n_gamma || cols != T.df || rows > GSM_MAX_NS) return fail(MSIM_ERR_ARG, "gsm term: q_const does not match gamma / knots");
    T.n_ns = rows;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) T.q[r][c] = tr ? t.q_const[c * t.q_cols + r] : t.q_const[r * t.q_cols + c];
    for (int r = 0; r < rows; r++) T.gamma[r] = t.gamma[r];
    gsm_term_finish(T);
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

int seed_from_ints(const int32_t in[6], double seed[6]) {
  for (int i = 0; i < 6; i++) {
    if (in[i] < 0) return fail(MSIM_ERR_SEED, "negative seed component");
    seed[i] = in[i];
  }
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  return MSIM_OK;
}

}  // namespace

// ====================================================================== C ABI

extern "C" {

const char *msim_last_error(void) { return g_err.c_str(); }
const char *msim_version(void) { return "microsimulation_b200 0.1 (sm_100a)"; }

int msim_init(int device) {
  if (g.inited && device != g.device) return fail(MSIM_ERR_STATE, "already bound to another device");
  g.device = device;
  return ensure_init();
}

void msim_shutdown(void) {
  if (!g.inited) return;
  cudaStreamSynchronize(g.stream);
  DevBuf *bufs[] = {&g.rowlog, &g.prov_start, &g.prov_end, &g.prov_meta, &g.tile_count, &g.partial, &g.dense, &g.small, &g.mt_words, &g.len, &g.starts, &g.chain_exit,
                    &g.chain_count, &g.chain_entry, &g.chain_base, &g.calib_out, &g.calib_par, &g.uniforms,
                    &g.first_end, &g.first_tag, &g.ov_fill, &g.ov_tag, &g.ov_start, &g.ov_end,
                    &g.aux, &g.code_in, &g.code_out, &g.sgeom, &g.indiv, &g.gsm_model, &g.gsm_data, &g.gsm_in, &g.gsm_out, &g.slices};
  for (DevBuf *b : bufs) {
    if (b->p) cudaFree(b->p);
    b->p = nullptr;
    b->bytes = 0;
  }
  cudaFree(g.d_lo);
  cudaFree(g.d_hi);
  if (g.pin_cache) cudaFreeHost(g.pin_cache);
  g.pin_cache = nullptr;
  g.pin_cache_bytes = 0;
  cudaEventDestroy(g.ev0);
  cudaEventDestroy(g.ev1);
  cudaStreamDestroy(g.own_stream);
  g.stream = g.own_stream = nullptr;
  g.inited = false;
  g.have_last = false;
}

int64_t msim_launch_count(void) { return g.launches; }

int msim_set_stream(void *cuda_stream, int use_own) {
  int rc;
  if ((rc = ensure_init())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  g.stream = use_own ? g.own_stream : (cudaStream_t)cuda_stream;
  return MSIM_OK;
}

int msim_set_rowlog_plan(double rows_per_person, double extra_rows_per_person) {
  if (!(rows_per_person > 0) || !(extra_rows_per_person >= 0)) return fail(MSIM_ERR_ARG, "bad plan");
  g.plan_rows = rows_per_person;
  g.plan_extra_rows = extra_rows_per_person;
  return MSIM_OK;
}

int msim_set_person_schedule(int which) {
  if (which != 0 && which != 1) return fail(MSIM_ERR_ARG, "schedule must be 0 (staged) or 1 (generic)");
  g.person_kernel = which;
  return MSIM_OK;
}

void msim_table_free(msim_table *t) {
  if (!t) return;
  host_free(t->data);
  t->data = nullptr;
  t->nrow = 0;
}
void msim_report_free(msim_report *r) {
  if (!r) return;
  msim_table_free(&r->pt);
  msim_table_free(&r->ut);
  msim_table_free(&r->events);
  msim_table_free(&r->prev);
  msim_table_free(&r->costs);
  msim_table_free(&r->indiv);
}

// ---- RNG state

void msim_r_create_current_stream(void) {
  g.rng.default_stream = g.rng.new_stream();
  g.rng.have_default = true;
}
void msim_r_remove_current_stream(void) { g.rng.have_default = false; }

int msim_r_set_user_random_seed(const double seed[6]) {
  if (!g.rng.have_default) return fail(MSIM_ERR_STATE, "no current stream: call msim_r_create_current_stream");
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed (the reference ignores it silently)");
  uint32_t s[6];
  seed_to_u32(seed, s);
  memcpy(g.rng.next_seed, s, sizeof s);  // Rng::SetPackageSeed
  g.rng.default_stream.set_all(s);       // default_stream->SetSeed
  return MSIM_OK;
}
int msim_r_get_user_random_seed(double seed[6]) {
  if (!g.rng.have_default) return fail(MSIM_ERR_STATE, "no current stream");
  for (int i = 0; i < 6; i++) seed[i] = (double)g.rng.default_stream.Cg.s[i];
  return MSIM_OK;
}
int msim_r_next_rng_substream(void) {
  if (!g.rng.have_default) return fail(MSIM_ERR_STATE, "no current stream");
  g.rng.default_stream.reset_next_substream();
  return MSIM_OK;
}
int msim_r_rng_advance_substream(double seed[6], const int *n) {
  HostStream r = g.rng.new_stream();  // `RngStream r;` consumes a package stream (microsimulation.cc:66)
  if (mrg_seed_ok(seed)) {            // SetSeed keeps the constructed state when the seed is illegal
    uint32_t s[6];
    seed_to_u32(seed, s);
    r.set_all(s);
  }
  r.advance_substream(0, *n);
  for (int i = 0; i < 6; i++) seed[i] = (double)r.Cg.s[i];
  return MSIM_OK;
}

int msim_mt_set_state(const uint32_t state[625]) {
  if (state[0] < 1 || state[0] > 624) return fail(MSIM_ERR_ARG, "MT position must be in 1..624");
  memcpy(g.rng.mt, state, sizeof g.rng.mt);
  g.rng.mt_seeded = true;
  return MSIM_OK;
}
int msim_mt_get_state(uint32_t state[625]) {
  if (!g.rng.mt_seeded) return fail(MSIM_ERR_STATE, "Mersenne-Twister state not set");
  memcpy(state, g.rng.mt, sizeof g.rng.mt);
  return MSIM_OK;
}
int msim_mt_set_seed(uint32_t seed) {
  g.rng.mt_set_seed(seed);
  return MSIM_OK;
}

// ---- model entry points

int msim_callPersonSimulation(const int32_t inseed[6], int32_t n, msim_table *out) {
  int rc;
  if (!out || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  msim_person_shard sh;
  memset(&sh, 0, sizeof sh);
  if ((rc = seed_from_ints(inseed, sh.seed))) return rc;
  // RngStream::SetPackageSeed(seed); rng["NH"] = new Rng(); rng["S"] = new Rng()  (person-r.cc:204-207)
  seed_to_u32(sh.seed, g.rng.next_seed);
  (void)g.rng.new_stream();
  (void)g.rng.new_stream();
  sh.first_person = 0;
  sh.n = n;
  sh.outputs = MSIM_OUT_ROWLOG;
  if ((rc = person_run(&sh))) return rc;
  return fetch_rowlog(ROWLOG_COLS_PERSON, out);
}

int msim_callSimplePerson(int32_t n, msim_table *out) {
  int rc;
  if (!out || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  simple_example::Consts c = {0};
  ReportGeom geom = ages_geom(simple_example::N_STATE, simple_example::N_EVENT, 0.0);
  if ((rc = sequential_run<SimpleTraits>(n, c, simple_example::MAX_DRAWS, OUT_ROWLOG, geom, simple_example::MAX_ROWS, 1.6)))
    return rc;
  return fetch_rowlog(ROWLOG_COLS_SIMPLE, out);
}

int msim_callSimplePerson2(int32_t n, msim_report *out) {
  int rc;
  if (!out || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  simple_example::Consts c = {0};
  ReportGeom geom = ages_geom(simple_example::N_STATE, simple_example::N_EVENT, 0.0);
  if ((rc = sequential_run<SimpleTraits>(n, c, simple_example::MAX_DRAWS, OUT_REPORT, geom, simple_example::MAX_ROWS, 1.6)))
    return rc;
  return fetch_report(out);
}

int msim_callIllnessDeath(int32_t n, double cure, double zsd, msim_report *out) {
  int rc;
  if (!out || n < 0 || !(zsd >= 0)) return fail(MSIM_ERR_ARG, "bad arguments");
  illness_death::Consts c = illness_consts(cure, zsd);
  ReportGeom geom = ages_geom(illness_death::N_STATE, illness_death::N_EVENT, 0.0);
  int max_draws = (zsd == 0.0) ? 5 : 7;  // rnorm(0, 0) draws nothing
  if ((rc = sequential_run<IllnessDeathTraits>(n, c, max_draws, OUT_REPORT, geom, illness_death::MAX_ROWS, 1.3))) return rc;
  return fetch_report(out);
}

int msim_calibration_sweep(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets, const double *runpar,
                           msim_calibration *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!out || !runpar || n < 0 || n_sets < 1 || n_sets > 65535 || first_person < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  if ((rc = ensure(g.calib_par, sizeof(double) * 6 * n_sets))) return rc;
  if ((rc = ensure(g.calib_out, sizeof(double) * 48 * n_sets))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemcpyAsync(g.calib_par.p, runpar, sizeof(double) * 6 * n_sets, cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemsetAsync(g.calib_out.p, 0, sizeof(double) * 48 * n_sets, g.stream));
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  CalibParams p;
  memset(&p, 0, sizeof p);
  int64_t tiles = (n + BLOCK - 1) / BLOCK;
  // enough blocks per set to fill the GPU even for few sets, no more than the work
  int64_t want = std::max<int64_t>(1, ((int64_t)g.sm_count * 8 + n_sets - 1) / n_sets);
  int grid_x = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, tiles), 2048));
  uint32_t state[6];
  seed_to_u32(seed, state);
  stream_start(state, first_person, (int64_t)grid_x * BLOCK, &p.st);
  p.n = n;
  p.n_sets = n_sets;
  p.runpar = (const double *)g.calib_par.p;
  p.out = (double *)g.calib_out.p;
  p.error = small_view().error;
  CK(cudaEventRecord(g.ev0, g.stream));
  if (n > 0) {
    calib_kernel<<<dim3(grid_x, n_sets), BLOCK, 0, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches++;
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  if ((rc = finish_run(nullptr))) return rc;
  std::vector<double> h(48 * (size_t)n_sets);
  CK(cudaMemcpy(h.data(), g.calib_out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  for (int s = 0; s < n_sets; s++) {
    msim_calibration *o = out + s;
    memset(o, 0, sizeof *o);
    const double *r = &h[48 * (size_t)s];
    for (int st = 0; st < 4; st++) {
      double tot = 0;
      for (int k = 0; k < 10; k++) {
        o->occupancy[st * 10 + k] = r[st * 10 + k];
        tot += r[st * 10 + k];
      }
      o->present[st] = tot > 0;  // report[stage] exists once any Count saw that stage (calibperson-r.cc:166-169)
    }
    o->n_time_at_risk = (int32_t)r[44];
    for (int k = 0; k < o->n_time_at_risk; k++) o->time_at_risk[k] = r[40 + k];
  }
  return MSIM_OK;
}

int msim_callCalibrationSimulation(int32_t n, const double runpar[6], msim_calibration *out) {
  if (!out || !runpar || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  // rng = new Rng(): takes the package seed, which then moves on one stream (calibperson-r.cc:182)
  HostStream s = g.rng.new_stream();
  double seed[6];
  for (int i = 0; i < 6; i++) seed[i] = (double)s.Cg.s[i];
  return msim_calibration_sweep(seed, 0, n, 1, runpar, out);
}

int msim_callSim_SickSicker(int32_t n, const double param[17], int indivp, int substreams, int64_t first_person,
                            msim_report *out) {
  return sick_sicker_run(n, param, indivp, substreams, first_person, out);
}

int msim_discounted(int mode, int64_t n, const double *y, const double *a, const double *b, double rate, double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n < 0 || (mode != 0 && mode != 1) || (n > 0 && (!y || !a || !out || (mode == 0 && !b)))) return fail(MSIM_ERR_ARG, "bad arguments");
  if (n == 0) return MSIM_OK;
  const size_t nn = (size_t)n;
  if ((rc = ensure(g.gsm_in, sizeof(double) * 3 * nn))) return rc;
  if ((rc = ensure(g.gsm_out, sizeof(double) * nn))) return rc;
  double *d = (double *)g.gsm_in.p;
  CK(cudaMemcpyAsync(d, y, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemcpyAsync(d + nn, a, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  if (mode == 0) CK(cudaMemcpyAsync(d + 2 * nn, b, sizeof(double) * nn, cudaMemcpyHostToDevice, g.stream));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)g.sm_count * 8));
  discount_kernel<<<grid, 256, 0, g.stream>>>(mode, n, d, d + nn, d + 2 * nn, rate, (double *)g.gsm_out.p);
  CK(cudaGetLastError());
  g.launches++;
  CK(cudaMemcpyAsync(out, g.gsm_out.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MSIM_OK;
}

int msim_gsm_randU(const msim_gsm *model, int64_t n, const double *u, const double *tentry, const int32_t *index,
                   double scale, double *out) {
  return gsm_run(model, 0, n, u, tentry, index, scale, out);
}
int msim_gsm_randU0(const msim_gsm *model, int64_t n, const double *u, const int32_t *index, double scale, double *out) {
  return gsm_run(model, 1, n, u, nullptr, index, scale, out);
}
int msim_gsm_eta(const msim_gsm *model, int64_t n, const double *y, const int32_t *index, double *out) {
  return gsm_run(model, 2, n, y, nullptr, index, 1.0, out);
}

// ---- sharded / device-resident forms

int msim_person_enqueue(const msim_person_shard *shard, msim_person_device_result *res) {
  int rc;
  if ((rc = person_enqueue(shard, 0))) return rc;
  if (res) {
    memset(res, 0, sizeof *res);
    res->n_rows = -1;  // known after msim_sync()
    res->capacity = g.last_capacity;
    res->d_rowlog = (shard->outputs & OUT_ROWLOG) ? (const double *)g.rowlog.p : nullptr;
    res->d_dense = (shard->outputs & OUT_REPORT) ? (const double *)g.dense.p : nullptr;
    res->d_counts = (const double *)g.dense.p + g.last_cells;
    res->dims.n_state = g.last_geom.n_state;
    res->dims.n_event = g.last_geom.n_event;
    res->dims.n_bin = g.last_geom.n_bin;
    res->dims.n_doubles = g.last_cells + N_COUNTS;
  }
  return MSIM_OK;
}

int msim_sync(void) {
  if (!g.inited) return fail(MSIM_ERR_STATE, "not initialised");
  return finish_run(nullptr);
}

int msim_person_run_device(const msim_person_shard *shard, msim_person_device_result *res) {
  int rc;
  if ((rc = person_run(shard))) return rc;
  if (res) {
    memset(res, 0, sizeof *res);
    res->n_rows = g.last_n_rows;
    res->capacity = g.last_capacity;
    res->d_rowlog = (shard->outputs & OUT_ROWLOG) ? (const double *)g.rowlog.p : nullptr;
    res->d_dense = (shard->outputs & OUT_REPORT) ? (const double *)g.dense.p : nullptr;
    res->d_counts = (const double *)g.dense.p + g.last_cells;
    res->dims.n_state = g.last_geom.n_state;
    res->dims.n_event = g.last_geom.n_event;
    res->dims.n_bin = g.last_geom.n_bin;
    res->dims.n_doubles = g.last_cells + N_COUNTS;
  }
  return MSIM_OK;
}

// CUDA events on the library's own stream, for callers that time a region of enqueued work
// (torch.cuda.Event only sees torch's streams).
int msim_timer_start(void) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!g.tev0) {
    CK(cudaEventCreate(&g.tev0));
    CK(cudaEventCreate(&g.tev1));
  }
  CK(cudaEventRecord(g.tev0, g.stream));
  return MSIM_OK;
}
int msim_timer_stop(float *ms) {
  if (!g.tev0) return fail(MSIM_ERR_STATE, "timer not started");
  CK(cudaEventRecord(g.tev1, g.stream));
  CK(cudaEventSynchronize(g.tev1));
  CK(cudaEventElapsedTime(ms, g.tev0, g.tev1));
  return MSIM_OK;
}

int msim_last_kernel_ms(float *ms) {
  if (!g.have_last && g.last_ms == 0.f) return fail(MSIM_ERR_STATE, "no run yet");
  *ms = g.last_ms;
  return MSIM_OK;
}

int msim_person_fetch_rowlog(msim_table *out) {
  if (!g.have_last || !(g.last_outputs & OUT_ROWLOG)) return fail(MSIM_ERR_STATE, "last run produced no row log");
  if (g.last_n_rows > g.last_capacity) return fail(MSIM_ERR_STATE, "row log overflowed its capacity; use msim_person_run_device");
  return fetch_rowlog(g.last_simple_names ? ROWLOG_COLS_SIMPLE : ROWLOG_COLS_PERSON, out);
}

int msim_person_fetch_counts(double counts[9]) {
  if (!g.have_last) return fail(MSIM_ERR_STATE, "no run yet");
  CK(cudaMemcpy(counts, (const double *)g.dense.p + g.last_cells, sizeof(double) * N_COUNTS, cudaMemcpyDeviceToHost));
  return MSIM_OK;
}

int msim_dense_to_report(const double *host_dense, const msim_dense_dims *dims, double discount_rate, msim_report *out) {
  if (!host_dense || !dims || !out) return fail(MSIM_ERR_ARG, "bad arguments");
  ReportGeom G = geom_from_dims(dims, discount_rate);
  if (G.total + N_COUNTS != dims->n_doubles) return fail(MSIM_ERR_ARG, "dims.n_doubles does not match the geometry");
  dense_to_report_impl(host_dense, G, out);
  return MSIM_OK;
}

int msim_dense_device_ptr(double **d_ptr, int64_t *n_doubles) {
  if (!g.have_last) return fail(MSIM_ERR_STATE, "no run yet");
  *d_ptr = (double *)g.dense.p;
  *n_doubles = g.last_cells + N_COUNTS;
  return MSIM_OK;
}

int msim_dense_fetch(double *host_dense, int64_t n_doubles) {
  if (!g.have_last) return fail(MSIM_ERR_STATE, "no run yet");
  if (n_doubles != g.last_cells + N_COUNTS) return fail(MSIM_ERR_ARG, "size mismatch");
  CK(cudaMemcpy(host_dense, g.dense.p, sizeof(double) * n_doubles, cudaMemcpyDeviceToHost));
  return MSIM_OK;
}

// ---- parity hooks

int msim_rng_uniforms(const double seed[6], int64_t first_substream, int64_t n_substreams, int32_t draws,
                      double *uniforms_host) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!uniforms_host || n_substreams < 0 || draws < 1 || first_substream < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  if (n_substreams == 0) return MSIM_OK;
  size_t bytes = sizeof(double) * (size_t)n_substreams * draws;
  if ((rc = ensure(g.uniforms, bytes))) return rc;
  UniformsParams p;
  int grid = (int)std::min<int64_t>((n_substreams + BLOCK - 1) / BLOCK, (int64_t)g.sm_count * 8);
  uint32_t state[6];
  seed_to_u32(seed, state);
  // substream j of this hook = j jumps from `seed`; stream_start applies first+1 jumps
  Mrg32k3a b;
  memcpy(b.s, state, sizeof b.s);
  b.jump(mrg_jump_pow(mrg_substream_jump(), (uint64_t)first_substream));
  memcpy(p.st.base, b.s, sizeof b.s);
  p.st.lo = g.d_lo;
  p.st.hi = g.d_hi;
  p.st.stride = mrg_jump_pow(mrg_substream_jump(), (uint64_t)grid * BLOCK);
  p.n_substreams = n_substreams;
  p.draws = draws;
  p.out = (double *)g.uniforms.p;
  rng_uniforms_kernel<<<grid, BLOCK, 0, g.stream>>>(p);
  CK(cudaGetLastError());
  g.launches++;
  CK(cudaMemcpyAsync(uniforms_host, g.uniforms.p, bytes, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MSIM_OK;
}

int msim_mt_uniforms(int64_t n, double *uniforms_host) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!uniforms_host || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
  if (!g.rng.mt_seeded) return fail(MSIM_ERR_STATE, "Mersenne-Twister state not set");
  if (n == 0) return MSIM_OK;
  const int64_t mti0 = g.rng.mt[0];
  const int64_t n_batches = (mti0 + n + 623) / 624;
  if ((rc = ensure(g.mt_words, sizeof(uint32_t) * (size_t)n_batches * 624))) return rc;
  if ((rc = ensure(g.uniforms, sizeof(double) * (size_t)n))) return rc;
  CK(cudaMemcpyAsync(g.mt_words.p, &g.rng.mt[1], sizeof(uint32_t) * 624, cudaMemcpyHostToDevice, g.stream));
  if (n_batches > 1) {
    mt_generate_kernel<<<1, 256, 0, g.stream>>>((uint32_t *)g.mt_words.p, (int)(n_batches - 1));
    CK(cudaGetLastError());
    g.launches++;
  }
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)g.sm_count * 8);
  mt_uniforms_kernel<<<grid, 256, 0, g.stream>>>((const uint32_t *)g.mt_words.p, mti0, n, (double *)g.uniforms.p);
  CK(cudaGetLastError());
  g.launches++;
  CK(cudaMemcpyAsync(uniforms_host, g.uniforms.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MSIM_OK;
}

}  // extern "C"
