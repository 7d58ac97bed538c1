/* msim.h — C ABI of the B200 engine for the microsimulation hot path.
 *
 * This is the drop-in boundary: an R-free restatement of the entry points the
 * reference registers with R in src/init.c:32-62 (its .Call and .C tables).
 * Every function below names the reference interface it replaces.  The R glue
 * a maintainer adds on top (SEXP <-> these structs) is shown in
 * INTEGRATION.md.  No torch, CUDA or C++ types appear in any signature.
 *
 * Conventions
 *   - every function that can fail returns MSIM_OK (0) or a negative code and
 *     records a message readable through msim_last_error(); nothing throws
 *     across this boundary;
 *   - one call at a time per process (the reference is single-threaded with
 *     static state, src/ssim.cc:35-79, 154-156);
 *   - tables are column-major (each column is one R numeric vector), live in
 *     page-locked host memory owned by the library, and are released with
 *     msim_table_free / msim_report_free;
 *   - there is NO CPU fallback: without a CUDA device every simulation entry
 *     returns MSIM_ERR_NO_DEVICE.
 */
#ifndef MSIM_H
#define MSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSIM_OK 0
#define MSIM_ERR_NO_DEVICE (-1)   /* no CUDA device / driver */
#define MSIM_ERR_CUDA (-2)        /* a CUDA call failed (message has the detail) */
#define MSIM_ERR_ARG (-3)         /* invalid argument */
#define MSIM_ERR_HEAP (-4)        /* a person needed more pending events than the device heap holds */
#define MSIM_ERR_SEED (-5)        /* illegal MRG32k3a seed (RngStream.cpp:234-268) */
#define MSIM_ERR_STREAM (-6)      /* sequential-stream model drew more uniforms per person than planned for */
#define MSIM_ERR_STATE (-7)       /* call order error (e.g. no current stream) */
#define MSIM_ERR_MODEL (-8)       /* a model reported a (state, event) pair outside the set it declared reachable */
#define MSIM_ERR_PEER (-9)        /* in-launch merge across GPUs: a rank did not deliver its vector in time */

/* ---- tables --------------------------------------------------------- */

typedef struct msim_table {
  int64_t nrow;
  int32_t ncol;
  const char *const *colnames; /* ncol static strings, the reference's R column names */
  double *data;                /* column j = data[j*nrow .. (j+1)*nrow) */
} msim_table;

/* EventReport::wrap() (microsimulation.h:974-986) / SummaryReport::asList()
 * (microsimulation.h:1125-1145): sparse, key-sorted tables.
 *   pt     : Key, age, pt            ut    : Key, age, utility
 *   events : Key, event, age, number prev  : Key, age, number
 *   costs  : Key, age, cost          (SummaryReport only, else nrow = 0)
 *   indiv  : utilities, costs        (SummaryReport with indivp, else nrow = 0)  */
typedef struct msim_report {
  msim_table pt, ut, events, prev, costs, indiv;
} msim_report;

/* callCalibrationSimulation's named list (calibperson-r.cc:176-200):
 * occupancy is 10 x 4 column-major, columns DiseaseFree, Precursor,
 * PreClinical, Clinical; `present[j]` says whether the reference's list would
 * contain column j at all (it only has stages that were ever counted).
 * TimeAtRisk has n_time_at_risk (<= 4) valid entries. */
typedef struct msim_calibration {
  double occupancy[40];
  int32_t present[4];
  double time_at_risk[4];
  int32_t n_time_at_risk;
} msim_calibration;

void msim_table_free(msim_table *t);
void msim_report_free(msim_report *r);

/* ---- library state --------------------------------------------------- */

const char *msim_last_error(void);
const char *msim_version(void);
/* Bind the calling process to a CUDA device (default: device 0 on first use). */
int msim_init(int device);
void msim_shutdown(void);
/* The CUDA stream all work is enqueued on: by default one the library creates;
 * use_own = 0 makes it `cuda_stream` (a cudaStream_t, e.g. the stream a
 * caller's NCCL all-reduce runs on, so that the merge needs no host
 * synchronisation), use_own = 1 returns to the library's own. */
int msim_set_stream(void *cuda_stream, int use_own);
/* Merging a sharded person-r run across the GPUs of ONE node inside the
 * simulation launch (csrc/person_staged.cuh, staged_peer_merge; replaces the
 * all-reduce a caller would otherwise enqueue after every step — SURVEY.md
 * 8(e), "one all-reduce of the accumulator vector").  One process per GPU:
 *   msim_peer_export(world, handle)          allocate this rank's mailbox, get
 *                                            its CUDA IPC handle (64 bytes);
 *   (exchange the handles, e.g. torch.distributed.all_gather_object)
 *   msim_peer_import(rank, world, handles)   map the other ranks' mailboxes
 *                                            (handles: world x 64 bytes, in
 *                                            rank order), then a host barrier;
 *   msim_peer_merge(1)                       from now on every person-r run
 *                                            with MSIM_OUT_REPORT leaves the SUM
 *                                            over all ranks in its dense vector
 *                                            (msim_dense_device_ptr), added in
 *                                            rank order: bit-identical on all
 *                                            ranks.  Every rank must make the
 *                                            same sequence of such runs; a rank
 *                                            that stays away for 4 s makes the
 *                                            others fail with MSIM_ERR_PEER.
 *   msim_peer_merge(2)                       the same, one launch later: run k
 *                                            adds up what the ranks sent in run
 *                                            k - 1 (nobody waits for a slower
 *                                            rank) and leaves it in the vector
 *                                            msim_peer_result_ptr names; the
 *                                            dense vector keeps this rank's own
 *                                            sums.  msim_peer_flush() adds up
 *                                            the last run's (one small launch)
 *                                            — after it the vector holds the
 *                                            latest run's merged result.
 *   msim_peer_merge(0) / msim_peer_close()   stop merging / release. */
#define MSIM_PEER_HANDLE_BYTES 64
int msim_peer_export(int32_t world, void *handle_out);
int msim_peer_import(int32_t rank, int32_t world, const void *handles);
int msim_peer_merge(int mode);
int msim_peer_flush(void);
int msim_peer_result_ptr(double **d_vector, int64_t *n_doubles);
int msim_peer_close(void);
/* Number of kernels this library has launched since it was loaded. */
int64_t msim_launch_count(void);

/* Which GPU schedule the person-r entry points use: 0 (default) = staged
 * (persons are regrouped between stages so that warps stay full,
 * csrc/person_staged.cuh), 1 = generic (one person per thread straight
 * through, the kernel every other model uses).  The calibration and Sick-Sicker
 * models keep their pending actions in an ordered array under 0 (a run in which
 * two pending actions compared equal is repeated with the heap) and in heap.h's
 * binary heap under 1.  The calibration sweep under 0 computes every person's
 * draws once for all parameter sets and needs no event queue
 * (csrc/calib_shared.cuh; sweeps it cannot run, and runs in which two event
 * times of a person were equal, go through the event queue); 3 = as 0, but
 * the calibration sweep always goes through the event queue (ordered array).
 * Same results either way; the switch exists so tests can check exactly
 * that.  2 = 0 with the repeat taken unconditionally (tests). */
int msim_set_person_schedule(int which);
/* R's Mersenne-Twister stream is generated in parallel segments started by jump-ahead when the table of jump
 * polynomials (mt_jump.bin, next to the library; written at build time) is present and the stream is long; 0
 * switches that off (one block generates the whole stream).  Same words either way; the switch exists so tests can
 * check exactly that.  Returns 1 if switched on without a table. */
int msim_set_mt_parallel(int on);
/* Sizing of the person-r row-log buffers: expected rows per person (default
 * 1.25; the model logs 1.232) and expected rows beyond a person's first
 * (default 0.26).  A run that needs more is repeated with the exact sizes, so
 * the plan only affects speed and memory, never results. */
int msim_set_rowlog_plan(double rows_per_person, double extra_rows_per_person);
/* Staged person-r schedule, scheduling only (results never depend on it): with two
 * thread blocks per SM, the share of the persons the blocks that arrive first on
 * their SMs simulate.  The warp schedulers favour those blocks; by default
 * (share <= 0) the library moves the share, from the finishing times of earlier
 * launches, to where both halves of the grid end together.  0.40 <= share <= 0.65
 * fixes it.  msim_last_wave_share: what the last launch used (0: it was not split). */
int msim_set_wave_share(double share);
int msim_last_wave_share(double *share);

/* ---- RNG state (replaces the .C table, src/init.c:11-30) ------------- */

/* src/microsimulation.cc:33-42 — default RngStream that serves unif_rand()
 * under RNGkind("user"); created in .onLoad, removed in .onUnload. */
void msim_r_create_current_stream(void);
void msim_r_remove_current_stream(void);
/* src/microsimulation.cc:44-51: package seed AND default-stream seed. */
int msim_r_set_user_random_seed(const double seed[6]);
/* src/microsimulation.cc:53-59 */
int msim_r_get_user_random_seed(double seed[6]);
/* src/microsimulation.cc:61-63 */
int msim_r_next_rng_substream(void);
/* src/microsimulation.cc:65-75: seed <- seed advanced by *n substreams
 * (side effect kept: it constructs an RngStream, advancing the package seed). */
int msim_r_rng_advance_substream(double seed[6], const int *n);
/* seed <- the start of the NEXT stream (A^(2^127) applied, RngStream.cpp:337-355): what base R's
 * parallel::nextRNGStream() does to the six seed words, which the reference's R code uses to give every mclapply
 * chunk its own stream (R/calibSim.R:21-24) and inside the RNGStream R class (R/rcpp_hello_world.R:793, 807). */
int msim_next_rng_stream(double seed[6]);
/* src/microsimulation.cc:77-85 — the hook R resolves BY NAME (`user_unif_rand`)
 * under RNGkind("user"): one U01 draw of the default stream, on the host (it
 * serves R-level runif() between simulations; the models never call it).
 * NULL if no stream exists, as in the reference. */
double *msim_user_unif_rand(void);

/* R's own generator, which the wrappers select with
 * RNGkind("Mersenne-Twister"); set.seed(12345) before the .Call
 * (R/rcpp_hello_world.R:325-336, 346-353, 374-381).  R glue copies
 * .Random.seed in before the call and back out after it (Rcpp::RNGScope,
 * illness-death.cpp:75). State = R's 625 words: position, then mt[624]. */
int msim_mt_set_state(const uint32_t state[625]);
int msim_mt_get_state(uint32_t state[625]);
int msim_mt_set_seed(uint32_t seed); /* == RNGkind("Mersenne-Twister"); set.seed(seed) */

/* ---- the five model entry points (replace the .Call table) ----------- */

/* callPersonSimulation(inseed, parms$n)   src/person-r.cc:195-226
 * Row log columns (std::map key order): endTime, event, id, startTime, state. */
int msim_callPersonSimulation(const int32_t inseed[6], int32_t n, msim_table *out);

/* callSimplePerson(parms$n)               src/simple-example.cc:72-83
 * Columns: endtime, event, id, startTime, state.  Draws from the MT state. */
int msim_callSimplePerson(int32_t n, msim_table *out);

/* callSimplePerson2(parms$n)              src/simple-example2.cc:62-83 */
int msim_callSimplePerson2(int32_t n, msim_report *out);

/* callIllnessDeath(parms$n, $cure, $zsd)  src/illness-death.cpp:73-94 */
int msim_callIllnessDeath(int32_t n, double cure, double zsd, msim_report *out);

/* callCalibrationSimulation(parms$n, $runpar)  src/calibperson-r.cc:176-200
 * Uses the package seed last set with msim_r_set_user_random_seed. */
int msim_callCalibrationSimulation(int32_t n, const double runpar[6], msim_calibration *out);

/* callSim(n, param, indivp) of the reference's documented SummaryReport
 * example, the continuous-time Sick-Sicker model (README.org:148-301; an
 * inline Rcpp function there, not a registered symbol).  param = r_HS1, r_HD,
 * r_S1H, r_S1S2, r_S1D, r_S2D, c_H, c_S1, c_S2, c_Trt, u_H, u_S1, u_S2, u_Trt,
 * discountRate, partitionBy, Trt.  substreams = 0: draws from R's
 * Mersenne-Twister state, person after person (the README's runs);
 * substreams = 1: person i draws from substream first_person+i+1 of a stream
 * made from the package seed.  out: pt, ut, events, prev, costs and indiv
 * (n rows, or one row of totals when indivp = 0), SummaryReport::asList. */
int msim_callSim_SickSicker(int32_t n, const double param[17], int indivp, int substreams, int64_t first_person,
                            msim_report *out);

/* simulations(n, discountRate, odShape, partitionBy) of the reference's dev-script example with costs
 * (test/test_microsimulation.R:1006-1093, and :782-865 for the id-indexed calls): a SummaryReport over
 * setPartition(0, 100, partitionBy) with SummaryReport::add for every interval and SummaryReport::addPointCost
 * (microsimulation.h:1105-1112) at cancer onset (15000) and at every two-yearly management visit (1000).  Person i
 * draws from substream first_person+i+1 of a stream made from the package seed.  out as for msim_callSim_SickSicker. */
int msim_callSim_CancerManagement(int32_t n, double odShape, double discountRate, double partitionBy, int indivp,
                                  int64_t first_person, msim_report *out);

/* CostReport<State>(discountRate, size, startReportAge) + setPartition(start, finish, delta) + add(state, time, cost)
 * for n point costs (inst/include/microsimulation.h:1198-1296), all booked to one individual (no individualReset
 * between them): out = wrap() — Key, age, cost for every cell a cost was booked to — and *current = the individual's
 * running discounted total (discounted from startReportAge on when that is not 0, microsimulation.h:1273-1277). */
int msim_cost_report(double discountRate, double startReportAge, double start, double finish, double delta, int32_t n_state,
                     int64_t n, const int32_t *state, const double *time, const double *cost, msim_table *out, double *current);

/* discountedInterval(y, start, end, rate) (mode 0; b = end) and discountedPoint(y, time, rate) (mode 1; b unused)
 * of inst/include/microsimulation.h:811-836, element-wise on the GPU: the device functions models use for
 * their own discounted totals; parity hook. */
int msim_discounted(int mode, int64_t n, const double *y, const double *a, const double *b, double rate, double *out);

/* ---- survival-time utilities and lookup tables --------------------------
 * The building blocks of the external FHCRC screening model, element-wise on the GPU.
 *
 * Rpexp(h, t, n_t) (inst/include/microsimulation.h:527-617), piecewise constant hazards h[i] from t[i]:
 *   mode 0  rand(u = a[i], from = b[i])  (b may be NULL: 0); an all-zero hazard tail is MSIM_ERR_ARG (Rcpp::stop)
 *   mode 1  H(a[i])   2  h(a[i])   3  p(a[i]) (lower tail)   4  d(a[i])                                        */
int msim_rpexp(int mode, int32_t n_t, const double *h, const double *t, int64_t n, const double *a, const double *b, double *out);
/* Rpexp_gamma(theta, mu, t, n_t) (microsimulation.h:626-704), gamma frailty Z[i]:
 *   mode 0  rand(u = a[i], Z[i], from[i])  (from may be NULL)   1  H(a[i], Z[i])   2  h(a[i], Z[i])   3  p(a[i], Z[i]) */
int msim_rpexp_gamma(int mode, int32_t n_t, double theta, const double *mu, const double *t, int64_t n, const double *a,
                     const double *Z, const double *from, double *out);
/* NumericInterpolate over points (x, y) (inst/include/rcpp_table.h:52-111):
 *   mode 0  approx(xfind)   1  operator()(xfind) (step)   2  invert(yfind)   3  invert_decreasing(yfind)          */
int msim_interpolate(int mode, int32_t n_pts, const double *x, const double *y, int64_t n, const double *xfind, double *out);
/* Table<key..., Outcome> built from the rows of a data frame (rcpp_table.h:136-324): rows is n_rows x (n_axes + 1)
 * row-major, keys first, outcome last; keys is n x n_axes.  On every axis the largest key <= the argument is taken
 * (the smallest if there is none); combinations the data frame does not hold give 0.                              */
int msim_table_lookup(int32_t n_axes, int64_t n_rows, const double *rows, int64_t n, const double *keys, double *out);

/* ---- generalised survival models (src/gsm.cpp, src/splines.cpp) ------- */

/* The fields of the Rcpp::List that gsm::gsm(Rcpp::List) reads (gsm.cpp:60-92). */
typedef struct msim_gsm_term {
  int32_t n_gamma;             /* length of gamma = rows of the natural-spline basis */
  int32_t n_knots;             /* interior knots */
  int32_t intercept, cure;
  int32_t q_rows, q_cols;      /* q_const as given (either orientation, splines.cpp:197-198), row-major */
  const double *gamma, *knots, *q_const;
  double Boundary_knots[2];
  const double *x;             /* [n_index] covariate value per covariate pattern */
} msim_gsm_term;

typedef struct msim_gsm {
  int32_t link;                /* 0 = "PH" (the only link the reference implements) */
  int32_t log_time;
  int32_t n_index;             /* covariate patterns: length of etap, etap0 and every term's x */
  int32_t n_terms;
  double tmin, tmax, inflate, t0;
  const double *etap, *etap0;
  const msim_gsm_term *terms;
} msim_gsm;

/* out[i] = gsm::randU(u[i], tentry[i], index[i], scale)   (gsm.cpp:38-47): the
 * time t with S(t | t > tentry) = u, by Brent's method on the spline scale, one
 * inversion per GPU thread.  tentry and index may be NULL (all 0). */
int msim_gsm_randU(const msim_gsm *model, int64_t n, const double *u, const double *tentry, const int32_t *index,
                   double scale, double *out);
/* out[i] = gsm::randU0(u[i], index[i], scale)             (gsm.cpp:48-58) */
int msim_gsm_randU0(const msim_gsm *model, int64_t n, const double *u, const int32_t *index, double scale, double *out);
/* out[i] = gsm::eta(y[i]) for covariate pattern index[i]   (gsm.cpp:15-21); parity hook */
int msim_gsm_eta(const msim_gsm *model, int64_t n, const double *y, const int32_t *index, double *out);

/* ---- sharded / device-resident forms (multi-GPU, benchmarks) --------- */

#define MSIM_OUT_ROWLOG 1u   /* life-history rows */
#define MSIM_OUT_REPORT 2u   /* EventReport attached to every handled event */
#define MSIM_OUT_COUNTS 4u   /* per-kind event counts + sum of end times */
#define MSIM_OUT_INDIV 8u    /* with MSIM_OUT_REPORT: every individual's discounted total (EventReport::individualReset) */

/* The EventReport attached to person-r runs (and to plugin models): what its constructor and setPartition take
 * (inst/include/microsimulation.h:873-904).
 *   setPartition(start, finish, delta, maxTime): points start, start+delta, ... <= finish (by repeated addition, as the
 *     reference builds them), then maxTime; at most 127 points
 *   startReportAge: an individual's discounted total (addBrief / individualReset, :905-932) counts from this age
 *   outputUtilities = 0: wrap() has no "ut" table
 * NULL restores the default: the built models' partition {0, 1, ..., 100, 1e6} (illness-death.cpp:81-85),
 * startReportAge 0, utilities on.  Times below `start` have no bin in the reference (lower_bound returns end());
 * here they are booked to the first bin. */
typedef struct msim_report_options {
  double start, finish, delta, max_time;
  double start_report_age;
  int32_t output_utilities;
} msim_report_options;
int msim_set_report_options(const msim_report_options *opt);

/* Dense report geometry (states x events x age bins; the partition {0,1,..,100,1e6} of every built model unless
 * msim_set_report_options chose another). */
#define MSIM_N_COUNTS 9 /* per-kind event counts [0..7] and sum(endTime) [8] */

typedef struct msim_dense_dims {
  int32_t n_state, n_event, n_bin;
  int64_t n_doubles; /* length of the accumulator vector: the dense report cells (none if no
                        report was requested) followed by the MSIM_N_COUNTS count cells */
} msim_dense_dims;

typedef struct msim_person_shard {
  double seed[6];        /* package seed (stream "NH" = first stream made from it) */
  int64_t first_person;  /* person i uses substream i+1 of that stream */
  int64_t n;
  uint32_t outputs;      /* MSIM_OUT_* */
  double discount_rate;  /* EventReport utility discounting (0 = none) */
} msim_person_shard;

typedef struct msim_person_device_result {
  int64_t n_rows;        /* rows written to the device row log */
  int64_t capacity;      /* rows the device row log can hold */
  const double *d_rowlog;   /* DEVICE pointer, 5 columns, column stride = capacity */
  const double *d_dense;    /* DEVICE pointer, dims.n_doubles accumulators (or NULL) */
  const double *d_counts;   /* DEVICE pointer: 8 per-kind counts, then sum(endTime) */
  msim_dense_dims dims;
} msim_person_device_result;

/* Run persons [first_person, first_person+n) of callPersonSimulation on the
 * current device and leave the results in device memory owned by the library
 * (valid until the next call).  Synchronous on return. */
int msim_person_run_device(const msim_person_shard *shard, msim_person_device_result *res);
/* Same, but asynchronous: only enqueues work on the library stream. */
int msim_person_enqueue(const msim_person_shard *shard, msim_person_device_result *res);
int msim_sync(void);
/* Average device time (ms) of the last enqueue/run, measured with CUDA events
 * on the library's stream around the simulation kernel. */
int msim_last_kernel_ms(float *ms);
/* CUDA events on the library's stream around a region of enqueued work:
 * start records, stop records + waits and returns the elapsed device time. */
int msim_timer_start(void);
int msim_timer_stop(float *ms);
/* Copy results of the last run to the host. */
int msim_person_fetch_rowlog(msim_table *out);
int msim_person_fetch_counts(double counts[9]);
/* After a run with MSIM_OUT_REPORT | MSIM_OUT_INDIV: out[i] = individual i's discounted total at individualReset
 * (wrap_indiv(); NaN = NA_REAL where the individual's last event came before startReportAge), and means[0..6] = what
 * wrap_means() tabulates (the reference's Means class, :468-496): n, mean, var, sd, se, sum, sumsq over the
 * individuals that count, summed in person order in long double as the reference does.  Either pointer may be NULL. */
int msim_person_fetch_indiv(double *out, int64_t n, double means[7]);
/* Accumulator vector (e.g. after an all-reduce done by the caller on the
 * device buffer) -> sparse report tables in the reference's layout.  Pure host
 * code: needs no device. */
int msim_dense_to_report(const double *host_dense, const msim_dense_dims *dims, double discount_rate,
                         msim_report *out);
/* Writable device pointer to the accumulator vector of the last run (report
 * cells + counts), for ONE in-place NCCL all-reduce(sum, f64) by the host layer:
 * every cell is a sum, integer counts are exact in f64. */
int msim_dense_device_ptr(double **d_ptr, int64_t *n_doubles);
int msim_dense_fetch(double *host_dense, int64_t n_doubles);

/* Calibration sweep: n_sets parameter vectors (row-major n_sets x 6), every
 * set run on persons [first_person, first_person+n) with the same seed
 * (common random numbers, R/calibSim.R:12-14). out has n_sets entries. */
int msim_calibration_sweep(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets,
                           const double *runpar, msim_calibration *out);

/* The sweep in pieces, for sharding over GPUs (BASELINE configs[4]): parameter sets [set_lo, set_hi) of the n_sets
 * given, on persons [first_person, first_person + n), results left on the device as ONE vector of 48 doubles per
 * set — occupancy[4][10], TimeAtRisk[4], and for k < 4 the number of thread blocks that added to TimeAtRisk[k] —
 * in which every cell is a sum: shards of the person range and shards of the sets both merge with a single
 * all-reduce(sum, f64) on *d_vector (rows of sets outside [set_lo, set_hi) are zero).  msim_calibration_fetch
 * copies the (merged) vector to the host, msim_calibration_unpack (pure host code) turns it into the reference's
 * per-set lists.  R/calibSim.R:10-40 does the same with mclapply chunks and Reduce("+"). */
int msim_calibration_sweep_device(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets, const double *runpar,
                                  int32_t set_lo, int32_t set_hi, double **d_vector, int64_t *n_doubles);
int msim_calibration_fetch(double *host_vector, int64_t n_doubles);
int msim_calibration_unpack(const double *host_vector, int32_t n_sets, msim_calibration *out);

/* K1 parity hook: uniforms[j*draws + k] = k-th uniform of substream
 * first_substream+j of the stream whose state is `seed` (substream 0 = seed). */
int msim_rng_uniforms(const double seed[6], int64_t first_substream, int64_t n_substreams, int32_t draws,
                      double *uniforms_host);
/* Measurement hook: one launch of building block `op` of the person-r path (0 empty loop, 1 substream jump, 2 U01,
 * 3 Weibull arithmetic (log + pow), 4 exp_rand, 5 event queue insert + pop, 6 EventReport::add, 7 exp) on n_threads
 * threads (a multiple of 256), `reps` repetitions per thread, every warp full and converged.  Run under ncu to read
 * the block's instruction count (scripts/inst_floor.py -> profiles/traffic.json "instruction_floor"); *ms is the
 * kernel's device time. */
int msim_microbench(int op, int64_t n_threads, int reps, float *ms);
/* MT19937 parity hook: the first n uniforms R's unif_rand() would return from
 * the current MT state (state is NOT advanced). */
int msim_mt_uniforms(int64_t n, double *uniforms_host);

#ifdef __cplusplus
}
#endif

#endif /* MSIM_H */
