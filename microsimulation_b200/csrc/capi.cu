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
#include "host_cohort.h"
#include "host_gsm.h"
#include "host_lookups.h"
#include "host_person.h"
#include "host_runtime.h"
#include "host_summary.h"
#include "calib_shared.cuh"
#include "microbench.cuh"

using namespace msim;

using namespace msim_host;

namespace {

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
                    &g.chain_count, &g.chain_entry, &g.chain_base, &g.chain_gexit, &g.chain_gcount, &g.chain_gentry, &g.chain_gbase, &g.calib_out, &g.calib_partial, &g.calib_par, &g.uniforms,
                    &g.first_end, &g.first_tag, &g.ov_fill, &g.ov_tag, &g.ov_start, &g.ov_end,
                    &g.aux, &g.code_in, &g.code_out, &g.sgeom, &g.indiv, &g.gsm_model, &g.gsm_data, &g.gsm_in, &g.gsm_out, &g.slices, &g.fold_sync, &g.carry_tiles};
  peer_close();
  if (g.d_mt_table) cudaFree(g.d_mt_table);
  g.d_mt_table = nullptr;
  g.mt_table_polys = 0;
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
  if (!(rows_per_person > 0) || !(extra_rows_per_person >= 0) || rows_per_person > 1e6 || extra_rows_per_person > 1e6)
    return fail(MSIM_ERR_ARG, "bad plan");
  g.plan_rows = rows_per_person;
  g.plan_extra_rows = extra_rows_per_person;
  return MSIM_OK;
}

int msim_set_wave_share(double share) {
  if (share > 0 && !(share >= 0.40 && share <= 0.65)) return fail(MSIM_ERR_ARG, "wave share: 0.40 to 0.65, or <= 0 to adapt");
  g.wave_share_fixed = share > 0 ? share : 0;
  if (!(share > 0))
    for (double &f : g.wave_share) f = 0.5;
  return MSIM_OK;
}

int msim_last_wave_share(double *share) {
  if (!share) return fail(MSIM_ERR_ARG, "null pointer");
  *share = g.have_last && g.last_staged ? g.last_wave_share : 0.0;
  return MSIM_OK;
}

int msim_set_report_options(const msim_report_options *opt) {
  if (!opt) {
    g.have_report_opts = false;
    return MSIM_OK;
  }
  ReportGeom G;
  if (!make_report_geom_partition(&G, 1, 1, opt->start, opt->finish, opt->delta, opt->max_time, 0.0, opt->start_report_age, opt->output_utilities))
    return fail(MSIM_ERR_ARG, "report options: the partition must ascend (delta > 0, finish >= start, maxTime > finish) and hold at most 127 points");
  if (!(opt->start_report_age >= 0)) return fail(MSIM_ERR_ARG, "report options: startReportAge must be >= 0");
  g.report_opts = *opt;
  g.have_report_opts = true;
  return MSIM_OK;
}

int msim_set_mt_parallel(int on) {
  g.mt_parallel = on != 0;
  return (on && g.inited && !g.d_mt_table) ? 1 : MSIM_OK;  // 1: asked for, but mt_jump.bin was not found
}

int msim_peer_export(int32_t world, void *handle_out) { return peer_export(world, handle_out); }
int msim_peer_import(int32_t rank, int32_t world, const void *handles) { return peer_import(rank, world, handles); }
int msim_peer_merge(int mode) {
  if (mode < 0 || mode > 2) return fail(MSIM_ERR_ARG, "peer merge: mode 0 (off), 1 (merged when the launch ends) or 2 (merged one launch later)");
  if (mode && (!g.peers.local || !g.peers.mapped[g.peers.rank])) return fail(MSIM_ERR_STATE, "peer merge: msim_peer_export and msim_peer_import first");
  if (g.peers.mode == 2 && mode != 2 && g.peers.epoch != g.peers.collected) return fail(MSIM_ERR_STATE, "peer merge: msim_peer_flush first");
  g.peers.mode = mode;
  return MSIM_OK;
}
int msim_peer_flush(void) { return peer_flush(); }
int msim_peer_result_ptr(double **d_vector, int64_t *n_doubles) {
  if (!d_vector || !n_doubles) return fail(MSIM_ERR_ARG, "null argument");
  if (g.peers.mode != 2 || !g.peers.out) return fail(MSIM_ERR_STATE, "peer merge: no run in mode 2 yet");
  *d_vector = g.peers.out;
  *n_doubles = g.peers.out_doubles;
  return MSIM_OK;
}
int msim_peer_close(void) { return peer_close(); }

int msim_set_person_schedule(int which) {
  if (which < 0 || which > 3)
    return fail(MSIM_ERR_ARG, "schedule must be 0 (default), 1 (generic / binary heap), 2 (test: 0 with the repeat paths forced) or 3 (as 0, calibration through the event queue)");
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
  if (!seed) return fail(MSIM_ERR_ARG, "null seed");
  if (!g.rng.have_default) return fail(MSIM_ERR_STATE, "no current stream: call msim_r_create_current_stream");
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed (the reference ignores it silently)");
  uint32_t s[6];
  seed_to_u32(seed, s);
  memcpy(g.rng.next_seed, s, sizeof s);  // Rng::SetPackageSeed
  g.rng.default_stream.set_all(s);       // default_stream->SetSeed
  return MSIM_OK;
}
int msim_r_get_user_random_seed(double seed[6]) {
  if (!seed) return fail(MSIM_ERR_ARG, "null seed");
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
  if (!seed || !n) return fail(MSIM_ERR_ARG, "null argument");
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

int msim_next_rng_stream(double seed[6]) {
  if (!seed || !mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  Mrg32k3a t;
  seed_to_u32(seed, t.s);
  t.jump(mrg_stream_jump());
  for (int i = 0; i < 6; i++) seed[i] = (double)t.s[i];
  return MSIM_OK;
}

double *msim_user_unif_rand(void) {
  static double rn;
  if (!g.rng.have_default) {
    fail(MSIM_ERR_STATE, "user_unif_rand(): no stream created yet");
    return nullptr;
  }
  rn = g.rng.default_stream.Cg.u01();
  return &rn;
}

int msim_mt_set_state(const uint32_t state[625]) {
  if (!state) return fail(MSIM_ERR_ARG, "null state");
  if (state[0] < 1 || state[0] > 624) return fail(MSIM_ERR_ARG, "MT position must be in 1..624");
  memcpy(g.rng.mt, state, sizeof g.rng.mt);
  g.rng.mt_seeded = true;
  return MSIM_OK;
}
int msim_mt_get_state(uint32_t state[625]) {
  if (!state) return fail(MSIM_ERR_ARG, "null state");
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
  if (!out || !inseed || n < 0) return fail(MSIM_ERR_ARG, "bad arguments");
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

// sets [set_lo, set_hi) of the n_sets parameter vectors on persons [first_person, first_person + n): the device
// vector [n_sets][48] (rows of the other sets zero), complete on return
static int calibration_run(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets, const double *runpar,
                           int32_t set_lo, int32_t set_hi) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!seed || !runpar || n < 0 || n_sets < 1 || n_sets > 65535 || first_person < 0 || set_lo < 0 || set_hi > n_sets || set_lo > set_hi)
    return fail(MSIM_ERR_ARG, "bad arguments");
  if (!mrg_seed_ok(seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  const int n_here = set_hi - set_lo;
  if ((rc = ensure(g.calib_par, sizeof(double) * 6 * n_sets))) return rc;
  if ((rc = ensure(g.calib_out, sizeof(double) * 48 * n_sets))) return rc;
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemcpyAsync(g.calib_par.p, runpar, sizeof(double) * 6 * n_sets, cudaMemcpyHostToDevice, g.stream));
  uint32_t state[6];
  seed_to_u32(seed, state);
  int *tie_flag = (int *)((double *)g.small.p + 13);
  g.last_calib_sets = n_sets;
  const int64_t tiles = (n + BLOCK - 1) / BLOCK;
  // Three kernels, same results.  SHARED (calib_shared.cuh): the persons' draws are computed once and every parameter
  // set runs over them, without an event queue; ORDERED / HEAP (calib_kernel): every (person, set) pair through
  // the event loop, pending actions in an ordered array / in heap.h's binary heap.  SHARED and ORDERED are exact
  // unless two of a person's event times are equal, which they report; such a sweep is repeated with HEAP.
  // Schedule 0: SHARED where the parameters allow it (calib::straight_ok), else ORDERED; 3: ORDERED; 1: HEAP;
  // 2 (tests): as 0 with the repeat taken unconditionally.
  enum { K_SHARED, K_ORDERED, K_HEAP };
  bool straight = n_here > 0;
  for (int s = set_lo; s < set_hi; s++) straight = straight && calib::straight_ok(runpar + 6 * (size_t)s);
  const int first = g.person_kernel == 1 ? K_HEAP : (g.person_kernel == 3 || !straight) ? K_ORDERED : K_SHARED;
  for (int kind = first;;) {
    CK(cudaMemsetAsync(g.calib_out.p, 0, sizeof(double) * 48 * n_sets, g.stream));
    CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
    CK(cudaEventRecord(g.ev0, g.stream));
    if (n > 0 && n_here > 0) {
      int grid_x;
      if (kind == K_SHARED) {
        // a block = up to BLOCK parameter sets x a share of the person tiles; 4 blocks fit an SM (shared memory)
        const int n_chunks = (n_here + BLOCK - 1) / BLOCK;
        CalibSharedParams p;
        memset(&p, 0, sizeof p);
        static int per_sm = 0;
        if (!per_sm) {
          CK(cudaFuncSetAttribute(calib_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_SMEM));
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, calib_shared_kernel, BLOCK, CS_SMEM));
          if (per_sm < 1) per_sm = 1;
        }
        // one wave of resident blocks, every block the same number of person tiles (give or take one)
        const int64_t slots = std::max<int64_t>(1, (int64_t)g.sm_count * per_sm / n_chunks);
        const int64_t tiles_per_block = (tiles + slots - 1) / slots;
        grid_x = (int)std::max<int64_t>(1, (tiles + tiles_per_block - 1) / tiles_per_block);
        stream_start(state, first_person, (int64_t)grid_x * BLOCK, &p.st);
        p.n = n;
        p.runpar = (const double *)g.calib_par.p;
        p.out = (double *)g.calib_out.p;
        p.set0 = set_lo;
        p.n_here = n_here;
        p.spb = (n_here + n_chunks - 1) / n_chunks;
        p.tie = tie_flag;
        if ((rc = ensure(g.calib_partial, sizeof(double) * 4 * (size_t)grid_x * n_here))) return rc;
        p.partial = (double *)g.calib_partial.p;
        calib_shared_kernel<<<dim3(grid_x, n_chunks), BLOCK, CS_SMEM, g.stream>>>(p);
      } else {
        CalibParams p;
        memset(&p, 0, sizeof p);
        // enough blocks per set to fill the GPU even for few sets, no more than the work
        const int64_t want = std::max<int64_t>(1, ((int64_t)g.sm_count * 8 + n_here - 1) / n_here);
        grid_x = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, tiles), 2048));
        stream_start(state, first_person, (int64_t)grid_x * BLOCK, &p.st);
        p.n = n;
        p.n_sets = n_sets;
        p.set0 = set_lo;
        p.runpar = (const double *)g.calib_par.p;
        p.out = (double *)g.calib_out.p;
        p.error = small_view().error;
        p.tie = tie_flag;
        if ((rc = ensure(g.calib_partial, sizeof(double) * 4 * (size_t)grid_x * n_here))) return rc;
        p.partial = (double *)g.calib_partial.p;
        if (kind == K_ORDERED) calib_kernel<true><<<dim3(grid_x, n_here), BLOCK, 0, g.stream>>>(p);
        else calib_kernel<false><<<dim3(grid_x, n_here), BLOCK, 0, g.stream>>>(p);
      }
      CK(cudaGetLastError());
      calib_fold_kernel<<<n_here, 32, 0, g.stream>>>((const double *)g.calib_partial.p, grid_x, set_lo, (double *)g.calib_out.p);
      CK(cudaGetLastError());
      g.launches += 2;
    }
    CK(cudaEventRecord(g.ev1, g.stream));
    g.timing_pending = true;
    if ((rc = finish_run(nullptr))) return rc;
    if (kind == K_HEAP) break;
    const int tie = *(const int *)(g.last_small + 13);  // finish_run has just read the flag block
    if (!tie && g.person_kernel != 2) break;
    kind = K_HEAP;
  }
  return MSIM_OK;
}

int msim_calibration_unpack(const double *host_vector, int32_t n_sets, msim_calibration *out) {
  if (!host_vector || !out || n_sets < 1) return fail(MSIM_ERR_ARG, "bad arguments");
  for (int s = 0; s < n_sets; s++) {
    msim_calibration *o = out + s;
    memset(o, 0, sizeof *o);
    const double *r = host_vector + 48 * (size_t)s;
    for (int st = 0; st < 4; st++) {
      double tot = 0;
      for (int k = 0; k < 10; k++) {
        o->occupancy[st * 10 + k] = r[st * 10 + k];
        tot += r[st * 10 + k];
      }
      o->present[st] = tot > 0;  // report[stage] exists once any Count saw that stage (calibperson-r.cc:166-169)
    }
    int nt = 0;
    while (nt < 4 && r[44 + nt] > 0) nt++;  // indices are reached in order (the loop at calibperson-r.cc:124-135 breaks)
    o->n_time_at_risk = nt;
    for (int k = 0; k < nt; k++) o->time_at_risk[k] = r[40 + k];
  }
  return MSIM_OK;
}

int msim_calibration_sweep_device(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets, const double *runpar,
                                  int32_t set_lo, int32_t set_hi, double **d_vector, int64_t *n_doubles) {
  int rc;
  if (!d_vector || !n_doubles) return fail(MSIM_ERR_ARG, "null argument");
  if ((rc = calibration_run(seed, first_person, n, n_sets, runpar, set_lo, set_hi))) return rc;
  *d_vector = (double *)g.calib_out.p;
  *n_doubles = 48 * (int64_t)n_sets;
  return MSIM_OK;
}

int msim_calibration_fetch(double *host_vector, int64_t n_doubles) {
  if (!host_vector) return fail(MSIM_ERR_ARG, "null argument");
  if (g.last_calib_sets < 1 || n_doubles != 48 * (int64_t)g.last_calib_sets) return fail(MSIM_ERR_STATE, "no calibration run of that size");
  CK(cudaMemcpyAsync(host_vector, g.calib_out.p, sizeof(double) * n_doubles, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return MSIM_OK;
}

int msim_calibration_sweep(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets, const double *runpar,
                           msim_calibration *out) {
  int rc;
  if (!out) return fail(MSIM_ERR_ARG, "bad arguments");
  if ((rc = calibration_run(seed, first_person, n, n_sets, runpar, 0, n_sets))) return rc;
  std::vector<double> h(48 * (size_t)n_sets);
  CK(cudaMemcpy(h.data(), g.calib_out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  return msim_calibration_unpack(h.data(), n_sets, out);
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

int msim_callSim_CancerManagement(int32_t n, double odShape, double discountRate, double partitionBy, int indivp,
                                  int64_t first_person, msim_report *out) {
  return cancer_management_run(n, odShape, discountRate, partitionBy, indivp, first_person, out);
}

int msim_cost_report(double discountRate, double startReportAge, double start, double finish, double delta, int32_t n_state,
                     int64_t n, const int32_t *state, const double *time, const double *cost, msim_table *out, double *current) {
  return cost_report_run(discountRate, startReportAge, start, finish, delta, n_state, n, state, time, cost, out, current);
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

int msim_rpexp(int mode, int32_t n_t, const double *h, const double *t, int64_t n, const double *a, const double *b, double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n_t < 1 || !h || !t || n < 0 || mode < 0 || mode > 4 || (n > 0 && (!a || !out))) return fail(MSIM_ERR_ARG, "bad arguments");
  std::vector<double> Hi(n_t);
  rpexp_cumulate(n_t, h, t, Hi.data());
  std::vector<const double *> ptr;
  if ((rc = upload_vectors({Hi, std::vector<double>(h, h + n_t), std::vector<double>(t, t + n_t)}, &ptr))) return rc;
  LookupParams p;
  memset(&p, 0, sizeof p);
  p.family = 0;
  p.mode = mode;
  p.pexp = Rpexp{n_t, ptr[0], ptr[1], ptr[2]};
  return lookup_run(p, n, n, a, mode == 0 ? b : nullptr, nullptr, out);
}

int msim_rpexp_gamma(int mode, int32_t n_t, double theta, const double *mu, const double *t, int64_t n, const double *a,
                     const double *Z, const double *from, double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n_t < 1 || !mu || !t || !(theta > 0) || n < 0 || mode < 0 || mode > 3 || (n > 0 && (!a || !Z || !out)))
    return fail(MSIM_ERR_ARG, "bad arguments");
  std::vector<double> H0(n_t);
  rpexp_gamma_cumulate(n_t, theta, mu, t, H0.data());
  std::vector<const double *> ptr;
  if ((rc = upload_vectors({H0, std::vector<double>(mu, mu + n_t), std::vector<double>(t, t + n_t)}, &ptr))) return rc;
  LookupParams p;
  memset(&p, 0, sizeof p);
  p.family = 1;
  p.mode = mode;
  p.pgam = RpexpGamma{n_t, theta, ptr[0], ptr[1], ptr[2]};
  return lookup_run(p, n, n, a, Z, mode == 0 ? from : nullptr, out);
}

int msim_interpolate(int mode, int32_t n_pts, const double *x, const double *y, int64_t n, const double *xfind, double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n_pts < 2 || !x || !y || n < 0 || mode < 0 || mode > 3 || (n > 0 && (!xfind || !out))) return fail(MSIM_ERR_ARG, "bad arguments");
  std::vector<double> slope(n_pts - 1);
  for (int i = 0; i < n_pts - 1; i++) slope[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);  // NumericInterpolate::prepare
  std::vector<const double *> ptr;
  if ((rc = upload_vectors({std::vector<double>(x, x + n_pts), std::vector<double>(y, y + n_pts), slope}, &ptr))) return rc;
  LookupParams p;
  memset(&p, 0, sizeof p);
  p.family = 2;
  p.mode = mode;
  p.interp = Interp{n_pts, ptr[0], ptr[1], ptr[2]};
  return lookup_run(p, n, n, xfind, nullptr, nullptr, out);
}

int msim_table_lookup(int32_t n_axes, int64_t n_rows, const double *rows, int64_t n, const double *keys, double *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (n_axes < 1 || n_axes > TABLE_MAX_AXES || n_rows < 1 || !rows || n < 0 || (n > 0 && (!keys || !out)))
    return fail(MSIM_ERR_ARG, "bad arguments");
  // axes = sorted distinct key values per column; outcomes on the dense grid, 0 where no row names the combination
  std::vector<std::vector<double> > vecs(n_axes);
  for (int a = 0; a < n_axes; a++) {
    for (int64_t r = 0; r < n_rows; r++) vecs[a].push_back(rows[r * (n_axes + 1) + a]);
    std::sort(vecs[a].begin(), vecs[a].end());
    vecs[a].erase(std::unique(vecs[a].begin(), vecs[a].end()), vecs[a].end());
  }
  size_t cells = 1;
  for (int a = 0; a < n_axes; a++) cells *= vecs[a].size();
  if (cells > ((size_t)1 << 26)) return fail(MSIM_ERR_ARG, "table grid too large");
  std::vector<double> values(cells, 0.0);
  for (int64_t r = 0; r < n_rows; r++) {  // later rows overwrite earlier ones, as map::operator[] = does
    size_t flat = 0;
    for (int a = 0; a < n_axes; a++)
      flat = flat * vecs[a].size() +
             (size_t)(std::lower_bound(vecs[a].begin(), vecs[a].end(), rows[r * (n_axes + 1) + a]) - vecs[a].begin());
    values[flat] = rows[r * (n_axes + 1) + n_axes];
  }
  vecs.push_back(values);
  std::vector<const double *> ptr;
  if ((rc = upload_vectors(vecs, &ptr))) return rc;
  LookupParams p;
  memset(&p, 0, sizeof p);
  p.family = 3;
  p.table.n_axes = n_axes;
  for (int a = 0; a < n_axes; a++) {
    p.table.n[a] = (int)vecs[a].size();
    p.table.axis[a] = ptr[a];
  }
  p.table.values = ptr[n_axes];
  return lookup_run(p, n, n * n_axes, keys, nullptr, nullptr, out);
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
  int rc;
  if ((rc = finish_run(nullptr))) return rc;
  return rowlog_overflow_check();
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
  if (!ms) return fail(MSIM_ERR_ARG, "null argument");
  if (!g.tev0) return fail(MSIM_ERR_STATE, "timer not started");
  CK(cudaEventRecord(g.tev1, g.stream));
  CK(cudaEventSynchronize(g.tev1));
  CK(cudaEventElapsedTime(ms, g.tev0, g.tev1));
  return MSIM_OK;
}

int msim_last_kernel_ms(float *ms) {
  if (!ms) return fail(MSIM_ERR_ARG, "null argument");
  if (!g.have_last && g.last_ms == 0.f) return fail(MSIM_ERR_STATE, "no run yet");
  *ms = g.last_ms;
  return MSIM_OK;
}

int msim_person_fetch_rowlog(msim_table *out) {
  if (!g.have_last || !(g.last_outputs & OUT_ROWLOG)) return fail(MSIM_ERR_STATE, "last run produced no row log");
  if (g.last_n_rows > g.last_capacity) return fail(MSIM_ERR_STATE, "row log overflowed its capacity; use msim_person_run_device");
  int rc;
  if ((rc = rowlog_overflow_check())) return rc;
  return fetch_rowlog(g.last_simple_names ? ROWLOG_COLS_SIMPLE : ROWLOG_COLS_PERSON, out);
}

int msim_person_fetch_counts(double counts[9]) {
  if (!counts) return fail(MSIM_ERR_ARG, "null argument");
  if (!g.have_last) return fail(MSIM_ERR_STATE, "no run yet");
  CK(cudaMemcpy(counts, (const double *)g.dense.p + g.last_cells, sizeof(double) * N_COUNTS, cudaMemcpyDeviceToHost));
  return MSIM_OK;
}

int msim_person_fetch_indiv(double *out, int64_t n, double means[7]) {
  if (!g.have_last || !g.last_staged || (g.last_outputs & (OUT_REPORT | OUT_INDIV)) != (OUT_REPORT | OUT_INDIV))
    return fail(MSIM_ERR_STATE, "the last run kept no individual totals (MSIM_OUT_REPORT | MSIM_OUT_INDIV)");
  if (n != g.last_n) return fail(MSIM_ERR_ARG, "size mismatch");
  std::vector<double> h((size_t)std::max<int64_t>(n, 1));
  if (n > 0) CK(cudaMemcpy(h.data(), g.indiv.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  if (out && n > 0) memcpy(out, h.data(), sizeof(double) * (size_t)n);
  if (means) {  // Means::operator+= in person order (microsimulation.h:476-481)
    long double sum = 0, sumsq = 0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++)
      if (h[(size_t)i] == h[(size_t)i]) {
        cnt++;
        sum += (long double)h[(size_t)i];
        sumsq += (long double)h[(size_t)i] * (long double)h[(size_t)i];
      }
    const double mean = (double)(sum / cnt);
    const double var = (double)((long double)cnt / (cnt - 1) * (sumsq / cnt - (long double)mean * mean));
    means[0] = (double)cnt;
    means[1] = mean;
    means[2] = var;
    means[3] = sqrt(var);
    means[4] = sqrt(var) / sqrt((double)cnt);
    means[5] = (double)sum;
    means[6] = (double)sumsq;
  }
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
  if (!d_ptr || !n_doubles) return fail(MSIM_ERR_ARG, "null argument");
  if (!g.have_last) return fail(MSIM_ERR_STATE, "no run yet");
  *d_ptr = (double *)g.dense.p;
  *n_doubles = g.last_cells + N_COUNTS;
  return MSIM_OK;
}

int msim_dense_fetch(double *host_dense, int64_t n_doubles) {
  if (!host_dense) return fail(MSIM_ERR_ARG, "null argument");
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
  if (!seed || !uniforms_host || n_substreams < 0 || draws < 1 || first_substream < 0) return fail(MSIM_ERR_ARG, "bad arguments");
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

// Instruction-weight microbenchmarks (csrc/microbench.cuh): one launch of building block `op` on n_threads threads,
// `reps` repetitions each; meant to be run under ncu (scripts/inst_floor.py).  Returns the kernel's device time.
int msim_microbench(int op, int64_t n_threads, int reps, float *ms) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (op < 0 || op >= MB_N_OPS || n_threads < 256 || n_threads > (int64_t)JUMP_TAB * JUMP_TAB || reps < 1 || !ms)
    return fail(MSIM_ERR_ARG, "bad arguments");
  const int grid = (int)(n_threads / 256);
  if ((rc = ensure(g.uniforms, sizeof(double) * (size_t)grid * 256))) return rc;
  MicroParams p;
  memset(&p, 0, sizeof p);
  uint32_t state[6] = {12345, 12345, 12345, 12345, 12345, 12345};
  stream_start(state, 0, (int64_t)grid * 256, &p.st);
  p.geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, 0.03);
  p.reps = reps;
  p.sink = (double *)g.uniforms.p;
  typedef void (*K)(const MicroParams);
  static const K kernels[MB_N_OPS] = {microbench_kernel<MB_EMPTY>, microbench_kernel<MB_JUMP>,     microbench_kernel<MB_U01>,
                                      microbench_kernel<MB_WEIBULL>, microbench_kernel<MB_EXP_RAND>, microbench_kernel<MB_HEAP>,
                                      microbench_kernel<MB_REPORT>, microbench_kernel<MB_EXP>};
  CK(cudaEventRecord(g.ev0, g.stream));
  kernels[op]<<<grid, 256, 0, g.stream>>>(p);
  CK(cudaGetLastError());
  g.launches++;
  CK(cudaEventRecord(g.ev1, g.stream));
  CK(cudaEventSynchronize(g.ev1));
  CK(cudaEventElapsedTime(ms, g.ev0, g.ev1));
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
    if ((rc = mt_generate((uint32_t *)g.mt_words.p, n_batches * 624))) return rc;
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
