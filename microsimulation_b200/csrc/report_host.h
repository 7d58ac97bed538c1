// Host-side finalisation shared by the C ABI (capi.cu) and the CPU emulation
// used in tests: dense accumulators -> the reference's sparse report tables,
// and model constants evaluated with the host's libm.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../include/msim.h"
#include "models.cuh"
#include "report.cuh"
#include "summary.cuh"

namespace msim {

const char *const ROWLOG_COLS_PERSON[5] = {"endTime", "event", "id", "startTime", "state"};
const char *const ROWLOG_COLS_SIMPLE[5] = {"endtime", "event", "id", "startTime", "state"};
const char *const PT_COLS[3] = {"Key", "age", "pt"};
const char *const UT_COLS[3] = {"Key", "age", "utility"};
const char *const EV_COLS[4] = {"Key", "event", "age", "number"};
const char *const PREV_COLS[3] = {"Key", "age", "number"};
const char *const COST_COLS[3] = {"Key", "age", "cost"};
const char *const INDIV_COLS[2] = {"utilities", "costs"};

inline void table_from_rows(const std::vector<double> &rows, int ncol, const char *const *cols, msim_table *t) {
  int64_t n = (int64_t)(rows.size() / ncol);
  t->nrow = n;
  t->ncol = ncol;
  t->colnames = cols;
  t->data = (double *)malloc(sizeof(double) * ncol * (n ? n : 1));
  for (int64_t i = 0; i < n; i++)
    for (int j = 0; j < ncol; j++) t->data[j * n + i] = rows[i * ncol + j];
}

// R's gammafn for 0 < x <= 10: Chebyshev series on
// (0,1) plus the recurrence.  Only used for illness-death.cpp:17-19's constants.
inline double host_gammafn(double x) {
  static const double gamcs[22] = {
      +.8571195590989331421920062399942e-2,  +.4415381324841006757191315771652e-2,
      +.5685043681599363378632664588789e-1,  -.4219835396418560501012500186624e-2,
      +.1326808181212460220584006796352e-2,  -.1893024529798880432523947023886e-3,
      +.3606925327441245256578082217225e-4,  -.6056761904460864218485548290365e-5,
      +.1055829546302283344731823509093e-5,  -.1811967365542384048291855891166e-6,
      +.3117724964715322277790254593169e-7,  -.5354219639019687140874081024347e-8,
      +.9193275519859588946887786825940e-9,  -.1577941280288339761767423273953e-9,
      +.2707980622934954543266540433089e-10, -.4646818653825730144081661058933e-11,
      +.7973350192007419656460767175359e-12, -.1368078209830916025799499172309e-12,
      +.2347319486563800657233471771688e-13, -.4027432614949066932766570534699e-14,
      +.6910051747372100912138336975257e-15, -.1185584500221992907052387126192e-15};
  int n = (int)x;
  double y = x - n;
  --n;
  double twox = (y * 2 - 1) * 2, b0 = 0, b1 = 0, b2 = 0;
  for (int i = 1; i <= 22; i++) {
    b2 = b1;
    b1 = b0;
    b0 = twox * b1 - b2 + gamcs[22 - i];
  }
  double value = .9375 + (b0 - b2) * 0.5;
  if (n == 0) return value;
  if (n < 0) return value / x;
  for (int i = 1; i <= n; i++) value *= (y + i);
  return value;
}

// ---------------------------------------------------------------- reports

// dense accumulators -> the reference's sparse, key-sorted tables
// (EventReport::wrap, microsimulation.h:974-986; row order = std::map order,
// microsimulation.h:1797-1860)
inline void dense_to_report_impl(const double *d, const ReportGeom &G, msim_report *out) {
  std::vector<double> pt, ut, ev, prev, none;
  const int nb = G.n_bin;
  bool any_event = false;
  for (int i = 0; i < G.n_state * G.n_event * nb; i++)
    if (d[G.off_events + i] != 0) {
      any_event = true;
      break;
    }
  if (any_event) {
    for (int s = 0; s < G.n_state; s++) {
      double starts = 0, ends = 0;  // running sums of dstart and of the intervals ended so far
      for (int b = 0; b < nb; b++) {
        double ended = 0;  // intervals in state s that end in bin b
        for (int e = 0; e < G.n_event; e++) ended += d[G.off_events + (s * G.n_event + e) * nb + b];
        starts += d[G.off_dstart + s * nb + b];
        ends += ended;
        const double full = starts - ends;  // intervals that cover bin b completely
        const double pv = full + ended - d[G.off_noedge + s * nb + b];
        const double part = d[G.off_pt + s * nb + b];
        const double age = report_part(G, b);
        if (full > 0 || ended > 0 || part != 0) {  // the bins EventReport::add visited
          double width = 0, dwidth = 0;
          if (full > 0) {
            double nxt = report_part(G, b + 1);
            width = nxt - age;
            dwidth = report_discounted(G, age, nxt, 1.0);
          }
          double ptv = part + full * width;
          double utv = (G.discount_rate == 0.0) ? ptv : d[G.off_ut + s * nb + b] + full * dwidth;
          pt.insert(pt.end(), {(double)s, age, ptv});
          ut.insert(ut.end(), {(double)s, age, utv});
        }
        if (pv > 0) prev.insert(prev.end(), {(double)s, age, pv});
      }
      for (int e = 0; e < G.n_event; e++)
        for (int b = 0; b < nb; b++) {
          double c = d[G.off_events + (s * G.n_event + e) * nb + b];
          if (c > 0) ev.insert(ev.end(), {(double)s, (double)e, report_part(G, b), c});
        }
    }
  }
  if (!G.output_utilities) ut.clear();  // EventReport(..., outputUtilities = false): wrap() has no "ut" (microsimulation.h:981-985)
  table_from_rows(pt, 3, PT_COLS, &out->pt);
  table_from_rows(ut, 3, UT_COLS, &out->ut);
  table_from_rows(ev, 4, EV_COLS, &out->events);
  table_from_rows(prev, 3, PREV_COLS, &out->prev);
  table_from_rows(none, 3, COST_COLS, &out->costs);
  table_from_rows(none, 2, INDIV_COLS, &out->indiv);
}

// SummaryReport::asList (microsimulation.h:1125-1145): pt, ut, events, prev, costs (sparse, key-sorted) and the
// per-individual discounted utilities and costs
inline void dense_to_summary_impl(const double *d, const SummaryGeom &G, const double *indiv, int64_t n_indiv,
                                  msim_report *out) {
  std::vector<double> pt, ut, ev, prev, cost, ind;
  const int nb = G.n_bin;
  bool any_event = false;
  for (int i = 0; i < G.n_state * G.n_event * nb; i++)
    if (d[G.off_events + i] != 0) {
      any_event = true;
      break;
    }
  if (any_event) {
    for (int s = 0; s < G.n_state; s++) {
      double starts = 0, ends = 0, wu = 0, wc = 0;
      for (int b = 0; b < nb; b++) {
        double ended = 0;
        for (int e = 0; e < G.n_event; e++) ended += d[G.off_events + (s * G.n_event + e) * nb + b];
        starts += d[G.off_dstart + s * nb + b];
        ends += ended;
        wu += d[G.off_wu0 + s * nb + b] - d[G.off_wu1 + s * nb + b];
        wc += d[G.off_wc0 + s * nb + b] - d[G.off_wc1 + s * nb + b];
        const double full = starts - ends;
        const double pv = full + ended - d[G.off_noedge + s * nb + b];
        const double part = d[G.off_pt + s * nb + b];
        const double age = G.part[b];
        const bool visited = full > 0 || ended > 0 || part != 0;
        const double pcn = d[G.off_pcn + s * nb + b];
        double utv = d[G.off_ut + s * nb + b], cv = d[G.off_cost + s * nb + b];
        if (visited) {
          double width = 0;
          if (full > 0 && b + 1 < nb) {
            const double nxt = G.part[b + 1];
            width = nxt - age;
            // weights of the intervals covering the bin completely, times the bin's discounted width
            utv += wu * (G.u_rate == 0.0 ? width : G.u_inv_alpha * (G.u_decay[b] - G.u_decay[b + 1]));
            cv += wc * (G.c_rate == 0.0 ? width : G.c_inv_alpha * (G.c_decay[b] - G.c_decay[b + 1]));
          }
          pt.insert(pt.end(), {(double)s, age, part + full * width});
          ut.insert(ut.end(), {(double)s, age, utv});
        }
        if (visited || pcn > 0) cost.insert(cost.end(), {(double)s, age, cv});
        if (pv > 0) prev.insert(prev.end(), {(double)s, age, pv});
      }
      for (int e = 0; e < G.n_event; e++)
        for (int b = 0; b < nb; b++) {
          double c = d[G.off_events + (s * G.n_event + e) * nb + b];
          if (c > 0) ev.insert(ev.end(), {(double)s, (double)e, G.part[b], c});
        }
    }
    for (int64_t i = 0; i < n_indiv; i++) ind.insert(ind.end(), {indiv[2 * i], indiv[2 * i + 1]});
  }
  table_from_rows(pt, 3, PT_COLS, &out->pt);
  table_from_rows(ut, 3, UT_COLS, &out->ut);
  table_from_rows(ev, 4, EV_COLS, &out->events);
  table_from_rows(prev, 3, PREV_COLS, &out->prev);
  table_from_rows(cost, 3, COST_COLS, &out->costs);
  table_from_rows(ind, 2, INDIV_COLS, &out->indiv);
}

inline ReportGeom ages_geom(int n_state, int n_event, double discount_rate) {
  // partition {0,1,...,100,1e6} (illness-death.cpp:81-85, simple-example2.cc:69-74)
  return make_report_geom(n_state, n_event, 102, 0.0, 1.0, 1.0e+6, discount_rate);
}

inline illness_death::Consts illness_consts(double cure, double zsd) {
  illness_death::Consts c;
  c.cure = cure;
  c.zsd = zsd;
  // b_weibull(mean, a, rr) = mean/gammafn(1+1/a)*pow(rr,-1/a)   (illness-death.cpp:17-19)
  volatile double one = 1.0;
  c.b_other = 80.0 / host_gammafn(1.0 + 1.0 / 4.0) * pow(one, -1.0 / 4.0);
  c.b_cancer_mean = 80.0 / host_gammafn(1.0 + 1.0 / 3.0);
  return c;
}


}  // namespace msim
