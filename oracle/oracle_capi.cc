// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// C entry points of the CPU oracle.  They mirror include/msim.h one for one
// (same structs, `oracle_` prefix) so a parity test calls both sides with the
// same arguments.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

#include "cprocess_lite.h"
#include "models.h"
#include "gsm_lite.h"
#include "lookups_lite.h"
#include "reports_lite.h"
#include "../examples/plugins/screening_skeleton.h"

using namespace orc;

namespace {

const char *const ROWLOG_COLS_PERSON[5] = {"endTime", "event", "id", "startTime", "state"};
const char *const ROWLOG_COLS_SIMPLE[5] = {"endtime", "event", "id", "startTime", "state"};
const char *const PT_COLS[3] = {"Key", "age", "pt"};
const char *const UT_COLS[3] = {"Key", "age", "utility"};
const char *const EV_COLS[4] = {"Key", "event", "age", "number"};
const char *const PREV_COLS[3] = {"Key", "age", "number"};
const char *const COST_COLS[3] = {"Key", "age", "cost"};
const char *const INDIV_COLS[2] = {"utilities", "costs"};

void fill_rowlog(const RowLog &log, const char *const *cols, msim_table *t) {
  int64_t n = (int64_t)log.id.size();
  t->nrow = n;
  t->ncol = 5;
  t->colnames = cols;
  t->data = (double *)malloc(sizeof(double) * 5 * (n ? n : 1));
  const std::vector<double> *src[5] = {&log.endTime, &log.event, &log.id, &log.startTime, &log.state};
  for (int j = 0; j < 5; j++)
    if (n) memcpy(t->data + j * n, src[j]->data(), sizeof(double) * n);
}

// row-major rows -> column-major table
void fill_table(const std::vector<double> &rows, int64_t n, int ncol, const char *const *cols, msim_table *t) {
  t->nrow = n;
  t->ncol = ncol;
  t->colnames = cols;
  t->data = (double *)malloc(sizeof(double) * ncol * (n ? n : 1));
  for (int64_t i = 0; i < n; i++)
    for (int j = 0; j < ncol; j++) t->data[j * n + i] = rows[i * ncol + j];
}

void fill_report(const ReportTables &r, msim_report *out) {
  fill_table(r.pt, r.n_pt, 3, PT_COLS, &out->pt);
  fill_table(r.ut, r.n_ut, 3, UT_COLS, &out->ut);
  fill_table(r.events, r.n_events, 4, EV_COLS, &out->events);
  fill_table(r.prev, r.n_prev, 3, PREV_COLS, &out->prev);
  fill_table(r.costs, r.n_costs, 3, COST_COLS, &out->costs);
  fill_table(r.indiv, r.n_indiv, 2, INDIV_COLS, &out->indiv);
}

}  // namespace

extern "C" {

const char *oracle_flavour(void) {
#ifdef MSIM_ORACLE_USE_REFERENCE
  return "reference";  // reference's ssim.cc + RngStream.cpp + heap.h compiled where they lie
#else
  return "port";  // restated engine + integer RngStream
#endif
}

void oracle_table_free(msim_table *t) {
  free(t->data);
  t->data = 0;
  t->nrow = 0;
}
void oracle_report_free(msim_report *r) {
  oracle_table_free(&r->pt);
  oracle_table_free(&r->ut);
  oracle_table_free(&r->events);
  oracle_table_free(&r->prev);
  oracle_table_free(&r->costs);
  oracle_table_free(&r->indiv);
}

uint64_t oracle_events_handled(void) { return g_events_handled; }
uint64_t oracle_draws_taken(void) { return rlite::draws_taken(); }
void oracle_reset_counters(void) {
  g_events_handled = 0;
  rlite::reset_draw_counter();
}

// ---- RNG state
void oracle_r_create_current_stream(void) { ssim::r_create_current_stream(); }
void oracle_r_remove_current_stream(void) { ssim::r_remove_current_stream(); }
int oracle_r_set_user_random_seed(const double seed[6]) {
  double s[6];
  memcpy(s, seed, sizeof s);
  ssim::r_set_user_random_seed(s);
  return 0;
}
int oracle_r_get_user_random_seed(double seed[6]) {
  ssim::r_get_user_random_seed(seed);
  return 0;
}
int oracle_r_next_rng_substream(void) {
  ssim::r_next_rng_substream();
  return 0;
}
int oracle_r_rng_advance_substream(double seed[6], const int *n) {
  int nn = *n;
  ssim::r_rng_advance_substream(seed, &nn);
  return 0;
}
// RngStream::AdvanceSubstream(e, c) on a stream seeded with `seed` (RngStream.cpp:454-458)
int oracle_rng_advance_e(double seed[6], int e, int c) {
  ssim::RngStream g;
  if (!g.SetSeed(seed)) return MSIM_ERR_SEED;
  g.AdvanceSubstream(e, c);
  g.GetState(seed);
  return 0;
}
int oracle_mt_set_seed(uint32_t seed) {
  rlite::RNGkind(rlite::MERSENNE_TWISTER);
  rlite::set_seed(seed);
  return 0;
}

// uniforms[j*draws+k]: k-th uniform of substream first_substream+j of the
// stream whose start state is `seed` (substream 0 = seed itself)
int oracle_rng_uniforms(const double seed[6], int64_t first_substream, int64_t n_substreams, int32_t draws,
                        double *uniforms) {
  ssim::RngStream g;
  if (!g.SetSeed(seed)) return MSIM_ERR_SEED;
  if (first_substream > 0) g.AdvanceSubstream(0, (int32_t)first_substream);
  for (int64_t j = 0; j < n_substreams; j++) {
    for (int k = 0; k < draws; k++) uniforms[j * draws + k] = g.RandU01();
    g.ResetNextSubstream();
  }
  return 0;
}

int oracle_mt_uniforms(int64_t n, double *uniforms) {
  for (int64_t i = 0; i < n; i++) uniforms[i] = unif_rand();
  return 0;
}

// nmath samplers driven by the current generator, for sampler-level parity
// which: 0 runif(a,b) 1 rweibull(a,b) 2 rexp(a) 3 rnorm(a,b) 4 exp_rand 5 norm_rand
int oracle_sample(int which, double a, double b, int64_t n, double *out) {
  for (int64_t i = 0; i < n; i++) {
    switch (which) {
      case 0: out[i] = R::runif(a, b); break;
      case 1: out[i] = R::rweibull(a, b); break;
      case 2: out[i] = R::rexp(a); break;
      case 3: out[i] = R::rnorm(a, b); break;
      case 4: out[i] = exp_rand(); break;
      case 5: out[i] = norm_rand(); break;
      default: return MSIM_ERR_ARG;
    }
  }
  return 0;
}
double oracle_qnorm(double p) { return R::qnorm(p, 0.0, 1.0, 1, 0); }
double oracle_gammafn(double x) { return R::gammafn(x); }
double oracle_discountedInterval(double y, double start, double end, double rate) {
  return ssim::discountedInterval(y, start, end, rate);
}

// ---- models
int oracle_callPersonSimulation(const int32_t inseed[6], int32_t n, msim_table *out) {
  double seed[6];
  for (int i = 0; i < 6; i++) seed[i] = inseed[i];
  RowLog log;
  run_person_r(seed, 0, n, true, &log, 0, 0, 0.0);
  fill_rowlog(log, ROWLOG_COLS_PERSON, out);
  return 0;
}

// sharded form: counts[0..7] events by kind, counts[8] = sum(endTime)
int oracle_person_run(const msim_person_shard *sh, msim_table *rowlog, double counts[9], msim_report *report) {
  RowLog log;
  ReportTables rt;
  run_person_r(sh->seed, sh->first_person, sh->n, rowlog != 0, &log, counts, report ? &rt : 0, sh->discount_rate);
  if (rowlog) fill_rowlog(log, ROWLOG_COLS_PERSON, rowlog);
  if (report) fill_report(rt, report);
  return 0;
}

// the same with the report's options (msim_report_options), every individual's discounted total and their means
int oracle_person_run_options(const msim_person_shard *sh, const msim_report_options *opt, double counts[9], msim_report *report,
                              double *indiv, double means[7]) {
  RowLog log;
  ReportTables rt;
  PersonReportOptions o;
  if (opt) {
    o.start = opt->start; o.finish = opt->finish; o.delta = opt->delta; o.max_time = opt->max_time;
    o.start_report_age = opt->start_report_age; o.output_utilities = opt->output_utilities;
  }
  run_person_r(sh->seed, sh->first_person, sh->n, false, &log, counts, &rt, sh->discount_rate, opt ? &o : 0, indiv, means);
  fill_report(rt, report);
  return 0;
}

int oracle_callSimplePerson(int32_t n, msim_table *out) {
  RowLog log;
  run_simple_person(n, &log);
  fill_rowlog(log, ROWLOG_COLS_SIMPLE, out);
  return 0;
}
int oracle_callSimplePerson2(int32_t n, msim_report *out) {
  ReportTables rt;
  run_simple_person2(n, &rt);
  fill_report(rt, out);
  return 0;
}
int oracle_callIllnessDeath(int32_t n, double cure, double zsd, msim_report *out) {
  ReportTables rt;
  run_illness_death(n, cure, zsd, &rt);
  fill_report(rt, out);
  return 0;
}
int oracle_callCalibrationSimulation(int32_t n, const double runpar[6], msim_calibration *out) {
  run_calibration(0, n, runpar, out);
  return 0;
}
int oracle_calibration_sweep(const double seed[6], int64_t first_person, int64_t n, int32_t n_sets,
                             const double *runpar, msim_calibration *out) {
  for (int s = 0; s < n_sets; s++) {
    double sd[6];
    memcpy(sd, seed, sizeof sd);
    if (!ssim::RngStream::SetPackageSeed(sd)) return MSIM_ERR_SEED;
    run_calibration(first_person, n, runpar + 6 * s, out + s);
  }
  return 0;
}

// Sick-Sicker (README.org:214-301). par = the 17 SickSickerParam fields in
// declaration order. trace (optional) receives 5 doubles per handled event:
// id, state, kind, previous, now — the README's debug print.
int oracle_sick_sicker(int64_t n, const double par[17], int indivp, int substreams, int64_t first_person,
                       msim_report *out, double *trace, int64_t trace_cap, int64_t *trace_len) {
  SickSickerParam p;
  p.r_HS1 = par[0]; p.r_HD = par[1]; p.r_S1H = par[2]; p.r_S1S2 = par[3]; p.r_S1D = par[4]; p.r_S2D = par[5];
  p.c_H = par[6]; p.c_S1 = par[7]; p.c_S2 = par[8]; p.c_Trt = par[9];
  p.u_H = par[10]; p.u_S1 = par[11]; p.u_S2 = par[12]; p.u_Trt = par[13];
  p.discountRate = par[14]; p.partitionBy = par[15]; p.Trt = (int)par[16];
  ReportTables rt;
  std::vector<double> tr;
  run_sick_sicker(n, &p, indivp != 0, substreams != 0, first_person, &rt, trace ? &tr : 0);
  fill_report(rt, out);
  if (trace) {
    int64_t m = (int64_t)tr.size() < trace_cap ? (int64_t)tr.size() : trace_cap;
    memcpy(trace, tr.data(), sizeof(double) * m);
    *trace_len = (int64_t)tr.size();
  }
  return 0;
}

int oracle_cancer_management(int64_t n, double odShape, double discountRate, double partitionBy, int indivp, int64_t first_person,
                             msim_report *out) {
  ReportTables rt;
  run_cancer_management(n, odShape, discountRate, partitionBy, indivp != 0, first_person, &rt);
  fill_report(rt, out);
  return 0;
}

// doc examples: return event times; tick_tock's finish time is the last one
int64_t oracle_tick_tock(double *times, int64_t cap) {
  std::vector<double> t;
  run_tick_tock(&t);
  for (size_t i = 0; i < t.size() && (int64_t)i < cap; i++) times[i] = t[i];
  return (int64_t)t.size();
}
int64_t oracle_doc_illness_death(int n, double *times, int64_t cap, double *qaly_cost) {
  std::vector<double> t, qc;
  run_doc_illness_death(n, &t, &qc);
  for (size_t i = 0; i < t.size() && (int64_t)i < cap; i++) times[i] = t[i];
  for (size_t i = 0; i < qc.size(); i++) qaly_cost[i] = qc[i];
  return (int64_t)t.size();
}
int oracle_priority_demo(double order[4]) {
  std::vector<double> o;
  run_priority_demo(&o);
  for (size_t i = 0; i < o.size() && i < 4; i++) order[i] = o[i];
  return (int)o.size();
}

// ---- generalised survival model (gsm_lite)
static orc::gsm make_gsm(const msim_gsm *m) {  // gsm::gsm(Rcpp::List), gsm.cpp:60-92
  orc::gsm g;
  g.tmin = m->tmin / m->inflate;
  g.tmax = m->tmax * m->inflate;
  g.etap.assign(m->etap, m->etap + m->n_index);
  g.etap0.assign(m->etap0, m->etap0 + m->n_index);
  for (int i = 0; i < m->n_terms; i++) {
    const msim_gsm_term &t = m->terms[i];
    orc::gsm_term term;
    term.gamma.assign(t.gamma, t.gamma + t.n_gamma);
    orc::vec knots(t.knots, t.knots + t.n_knots), B(t.Boundary_knots, t.Boundary_knots + 2);
    std::vector<orc::vec> q(t.q_rows, orc::vec(t.q_cols));
    for (int r = 0; r < t.q_rows; r++)
      for (int c = 0; c < t.q_cols; c++) q[r][c] = t.q_const[r * t.q_cols + c];
    term.ns1 = orc::ns(B, knots, q, t.intercept, t.cure);
    term.x.assign(t.x, t.x + m->n_index);
    g.terms.push_back(term);
  }
  g.log_time = m->log_time != 0;
  g.target = g.target0 = 0.0;
  g.index = 0;
  g.t0 = m->t0;
  return g;
}
int oracle_gsm_randU(const msim_gsm *m, int64_t n, const double *u, const double *tentry, const int32_t *index, double scale,
                     double *out) {
  orc::gsm g = make_gsm(m);
  for (int64_t i = 0; i < n; i++) out[i] = g.randU(u[i], tentry ? tentry[i] : 0.0, index ? index[i] : 0, scale);
  return 0;
}
int oracle_gsm_randU0(const msim_gsm *m, int64_t n, const double *u, const int32_t *index, double scale, double *out) {
  orc::gsm g = make_gsm(m);
  for (int64_t i = 0; i < n; i++) out[i] = g.randU0(u[i], index ? index[i] : 0, scale);
  return 0;
}
int oracle_gsm_eta(const msim_gsm *m, int64_t n, const double *y, const int32_t *index, double *out) {
  orc::gsm g = make_gsm(m);
  for (int64_t i = 0; i < n; i++) {
    g.index = index ? index[i] : 0;
    out[i] = g.eta(y[i]);
  }
  return 0;
}
// ns::eval(x, der) of term `term`: n_gamma values per x (scipy cross-check of the spline code)
int oracle_ns_eval(const msim_gsm *m, int term, int64_t n, const double *x, int der, double *out) {
  orc::gsm g = make_gsm(m);
  for (int64_t i = 0; i < n; i++) {
    orc::vec v = g.terms[term].ns1.eval(x[i], der);
    for (size_t k = 0; k < v.size(); k++) out[i * v.size() + k] = v[k];
  }
  return 0;
}
// bs::eval(x, der) of term `term`: df values per x
int oracle_bs_eval(const msim_gsm *m, int term, int64_t n, const double *x, int der, double *out) {
  orc::gsm g = make_gsm(m);
  for (int64_t i = 0; i < n; i++) {
    orc::vec v = g.terms[term].ns1.bs::eval(x[i], der);
    for (size_t k = 0; k < v.size(); k++) out[i * v.size() + k] = v[k];
  }
  return 0;
}

// CostReport (reports_lite.h): n point costs; rows (state, age, cost) in key order; returns the row count
int oracle_cost_report(double rate, double start_age, double p_start, double p_finish, double p_delta, int64_t n, const int *state,
                       const double *time, const double *cost, double *out_rows, int64_t cap, double *current) {
  ssim::CostReport<short> rep(rate, start_age);
  rep.setPartition(p_start, p_finish, p_delta);
  for (int64_t i = 0; i < n; i++) rep.add((short)state[i], time[i], cost[i]);
  *current = rep.current;
  int64_t k = 0;
  const std::map<std::pair<short, double>, double> sorted(rep._table.begin(), rep._table.end());  // wrap_map order
  for (auto it = sorted.begin(); it != sorted.end() && k < cap; ++it, ++k) {
    out_rows[3 * k] = it->first.first;
    out_rows[3 * k + 1] = it->first.second;
    out_rows[3 * k + 2] = it->second;
  }
  return (int)k;
}

// ---- the screening-model skeleton (examples/plugins/screening_skeleton.h): Rpexp + Table + NumericInterpolate + gsm +
// EventReport in one model, restated on the oracle's building blocks.  Person i draws from substream
// first_person + i + 1 of a stream made from the package seed.
namespace scr {
using namespace ssim;
// R::rnormPos (src/microsimulation.cc:103-107): a normal draw conditioned on being non-negative, by rejection
static double rnormPos(double mean, double sd) {
  double x;
  while ((x = R::rnorm(mean, sd)) < 0.0) {
  }
  return x;
}
struct Model {
  const screening_inputs *in;
  orc::Rpexp *mortality;
  orc::Table *biopsy;
  orc::NumericInterpolate *sensitivity, *survival;
  orc::gsm *sojourn;
  EventReport<short, short> *report;
};
class Person : public cProcess {
 public:
  Model *M;
  int state, grade;
  bool dx, screen_dx;
  double age_o, beta0, beta1, beta2, eps, age_dx;
  Person(Model *m) : M(m) {}
  void init() {
    const screening_inputs &p = *M->in;
    state = SCR_Healthy;
    dx = screen_dx = false;
    age_dx = 0.0;
    age_o = 35.0 + sqrt(-2.0 * log(R::runif(0.0, 1.0)) / p.g0);
    grade = (R::runif(0.0, 1.0) < 0.006 * (age_o - 35.0)) ? 1 : 0;
    beta0 = R::rnorm(p.mubeta0, p.sebeta0);
    beta1 = rnormPos(p.mubeta1, p.sebeta1);
    beta2 = rnormPos(p.mubeta2[grade], p.sebeta2[grade]);
    eps = R::rnorm(0.0, p.tau2);
    scheduleAt(M->mortality->rand(R::runif(0.0, 1.0)), SCR_toOtherDeath);
    scheduleAt(age_o, SCR_toLocalised);
    if (p.screen > 0) scheduleAt(50.0, SCR_toScreen);
  }
  void handleMessage(const cMessage *msg) {
    const screening_inputs &p = *M->in;
    M->report->add((short)state, msg->kind, previous(), now());
    switch (msg->kind) {
      case SCR_toOtherDeath:
      case SCR_toCancerDeath:
        Sim::stop_simulation();
        break;
      case SCR_toLocalised:
        state = SCR_Localised;
        scheduleAt(now() + M->sojourn->randU(R::runif(0.0, 1.0), 0.0, grade), SCR_toClinicalDiagnosis);
        break;
      case SCR_toScreen: {
        const double psa = exp(std::min(log(20.0), beta0 + beta1 * (now() - 35.0) + beta2 * std::max(0.0, now() - age_o) + eps));
        if (psa >= p.psa_threshold) {
          std::vector<double> key(2);
          key[0] = psa;
          key[1] = now();
          if (R::runif(0.0, 1.0) < p.compliance * (*M->biopsy)(key)) scheduleAt(now(), SCR_toBiopsy);
        }
        if (now() + p.screen <= 70.0) scheduleAt(now() + p.screen, SCR_toScreen);
        break;
      }
      case SCR_toBiopsy:
        if (state == SCR_Localised && R::runif(0.0, 1.0) < (*M->sensitivity)(p.cohort + now())) scheduleAt(now(), SCR_toScreenDiagnosis);
        break;
      case SCR_toClinicalDiagnosis:
      case SCR_toScreenDiagnosis:
        if (!dx) {
          dx = true;
          screen_dx = msg->kind == SCR_toScreenDiagnosis;
          age_dx = now();
          state = SCR_Diagnosed;
          RemoveKind(SCR_toScreen);
          RemoveKind(SCR_toBiopsy);
          RemoveKind(SCR_toClinicalDiagnosis);
          RemoveKind(SCR_toScreenDiagnosis);
          const double u = R::runif(0.0, 1.0);
          if (u >= p.surv_value[p.n_surv - 1]) scheduleAt(now() + M->survival->invert_decreasing(u), SCR_toCancerDeath);
        }
        break;
      default:
        break;
    }
  }
};
}  // namespace scr

int oracle_screening_skeleton(const screening_inputs *in, int64_t first_person, int64_t n, double discount_rate, msim_report *report,
                              double *values) {
  using namespace scr;
  orc::Rpexp mortality(in->mort_rate, in->mort_age, in->n_mort);
  orc::Table biopsy(2);
  for (int i = 0; i < in->n_biopsy; i++) {
    std::vector<double> key(in->biopsy_rows + 3 * i, in->biopsy_rows + 3 * i + 2);
    biopsy.insert(key, in->biopsy_rows[3 * i + 2]);
  }
  orc::NumericInterpolate sensitivity(in->sens_year, in->sens_value, in->n_sens), survival(in->surv_time, in->surv_value, in->n_surv);
  orc::gsm sojourn = make_gsm(in->gsm);
  ssim::EventReport<short, short> rep(discount_rate);
  std::vector<double> ages;
  for (int a = 0; a <= 100; a++) ages.push_back(a);
  ages.push_back(1.0e+6);
  rep.clear();
  rep.setPartition(ages);
  Model M = {in, &mortality, &biopsy, &sensitivity, &survival, &sojourn, &rep};
  rlite::RNGkind(rlite::USER_UNIF);
  rlite::set_user_unif_rand(ssim::user_unif_rand);
  ssim::Rng *rng = new ssim::Rng();
  rng->set();
  if (first_person > 0) rng->AdvanceSubstream(0, (int32_t)first_person);
  Person person(&M);
  for (int64_t i = 0; i < n; i++) {
    rng->nextSubstream();
    ssim::Sim::create_process(&person);
    ssim::Sim::run_simulation();
    values[SCR_N_VALUES * i + 0] = ssim::now();
    values[SCR_N_VALUES * i + 1] = person.age_dx;
    values[SCR_N_VALUES * i + 2] = person.screen_dx ? 1.0 : 0.0;
    values[SCR_N_VALUES * i + 3] = person.grade;
    ssim::Sim::clear();
  }
  delete rng;
  ReportTables rt;
  flatten(rep, &rt);
  fill_report(rt, report);
  return 0;
}

// ---- survival-time utilities and lookup tables (lookups_lite.h); modes as in include/msim.h
int oracle_rpexp(int mode, int n_t, const double *h, const double *t, int64_t n, const double *a, const double *b, double *out) {
  orc::Rpexp r(h, t, n_t);
  for (int64_t i = 0; i < n; i++) {
    try {
      out[i] = mode == 0 ? r.rand(a[i], b ? b[i] : 0.0) : mode == 1 ? r.H(a[i]) : mode == 2 ? r.h(a[i]) : mode == 3 ? r.p(a[i]) : r.d(a[i]);
    } catch (std::exception &) {
      return -3;  // Rcpp::stop in the reference
    }
  }
  return 0;
}
int oracle_rpexp_gamma(int mode, int n_t, double theta, const double *mu, const double *t, int64_t n, const double *a, const double *Z,
                       const double *from, double *out) {
  orc::Rpexp_gamma r(theta, mu, t, (size_t)n_t);
  for (int64_t i = 0; i < n; i++)
    out[i] = mode == 0 ? r.rand(a[i], Z[i], from ? from[i] : 0.0) : mode == 1 ? r.H(a[i], Z[i]) : mode == 2 ? r.h(a[i], Z[i]) : r.p(a[i], Z[i]);
  return 0;
}
int oracle_interpolate(int mode, int n_pts, const double *x, const double *y, int64_t n, const double *xfind, double *out) {
  orc::NumericInterpolate f(x, y, n_pts);
  for (int64_t i = 0; i < n; i++)
    out[i] = mode == 0 ? f.approx(xfind[i]) : mode == 1 ? f(xfind[i]) : mode == 2 ? f.invert(xfind[i]) : f.invert_decreasing(xfind[i]);
  return 0;
}
// the table as the rows of a data frame: n_rows x (n_axes keys + outcome), row-major
int oracle_table_lookup(int n_axes, int64_t n_rows, const double *rows, int64_t n, const double *keys, double *out) {
  orc::Table T(n_axes);
  for (int64_t r = 0; r < n_rows; r++) {
    std::vector<double> key(rows + r * (n_axes + 1), rows + r * (n_axes + 1) + n_axes);
    T.insert(key, rows[r * (n_axes + 1) + n_axes]);
  }
  for (int64_t i = 0; i < n; i++) out[i] = T(std::vector<double>(keys + i * n_axes, keys + (i + 1) * n_axes));
  return 0;
}

}  // extern "C"
