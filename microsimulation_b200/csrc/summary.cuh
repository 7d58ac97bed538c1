// Dense, O(1)-per-event restatement of the reference's SummaryReport
// (inst/include/microsimulation.h:1010-1190): EventReport's person-time,
// events and prevalence plus utilities and costs that the MODEL sets before each
// interval (setUtility / setCost), separate discount rates for the two, point
// costs, and per-individual discounted totals.
//
// Cells per (state s, bin b), all doubles in one vector (report.cuh explains
// the counting scheme; the weighted parts are new):
//   pt, ut, cost          partial overlaps (ut, cost weighted and discounted)
//   wu0, wu1, wc0, wc1    weighted difference arrays: an interval with utility u
//                         that covers bins [f0, hi) completely adds u to wu0[f0]
//                         and to wu1[hi]; sum_{j<=b} (wu0 - wu1) is then the total
//                         utility weight on bin b, to be multiplied by the bin's
//                         discounted width once, at finalisation (same for costs)
//   dstart, noedge        as in report.cuh
//   pcn                   number of point costs booked to the bin (a point cost
//                         creates a cost row even where no interval passed)
//   events[s][e][b]
// Individual totals: indiv[2*i] += discounted utility, indiv[2*i+1] += discounted
// cost of person i (index 0 only when indivp is false).
#pragma once
#include <math.h>
#include <stdint.h>

#include "report.cuh"

namespace msim {

struct SummaryGeom {
  int n_state, n_event, n_bin;
  double part[REPORT_MAX_BINS];      // partition points, ascending; the last is the catch-all (1e100)
  double u_decay[REPORT_MAX_BINS];   // exp(-part[i]*u_alpha)
  double c_decay[REPORT_MAX_BINS];   // exp(-part[i]*c_alpha)
  double u_rate, u_alpha, u_inv_alpha;
  double c_rate, c_alpha, c_inv_alpha;
  int off_pt, off_ut, off_cost, off_wu0, off_wu1, off_wc0, off_wc1, off_dstart, off_noedge, off_pcn, off_events, total;
  int n_float;  // cells [0, n_float) are FP sums, the rest counters
};

// setPartition(start, finish, delta, maxTime): points by REPEATED addition, as
// the reference does (microsimulation.h:1047-1052), then maxTime
inline bool make_summary_geom(SummaryGeom *g, int n_state, int n_event, double start, double finish, double delta,
                              double max_time, double u_rate, double c_rate) {
  int nb = 0;
  if (!(delta > 0)) return false;
  for (double t = start; t <= finish; t += delta) {
    if (nb >= REPORT_MAX_BINS - 1) return false;
    g->part[nb++] = t;
  }
  g->part[nb++] = max_time;
  for (int i = nb; i < REPORT_MAX_BINS; i++) g->part[i] = max_time;
  g->n_state = n_state;
  g->n_event = n_event;
  g->n_bin = nb;
  g->u_rate = u_rate;
  g->c_rate = c_rate;
  g->u_alpha = log(1.0 + u_rate);
  g->c_alpha = log(1.0 + c_rate);
  g->u_inv_alpha = u_rate != 0.0 ? 1.0 / g->u_alpha : 0.0;
  g->c_inv_alpha = c_rate != 0.0 ? 1.0 / g->c_alpha : 0.0;
  for (int i = 0; i < REPORT_MAX_BINS; i++) {
    g->u_decay[i] = u_rate != 0.0 ? exp(-g->part[i] * g->u_alpha) : 1.0;
    g->c_decay[i] = c_rate != 0.0 ? exp(-g->part[i] * g->c_alpha) : 1.0;
  }
  const int sb = n_state * nb;
  int o = 0;
  g->off_pt = o; o += sb;
  g->off_ut = o; o += sb;
  g->off_cost = o; o += sb;
  g->off_wu0 = o; o += sb;
  g->off_wu1 = o; o += sb;
  g->off_wc0 = o; o += sb;
  g->off_wc1 = o; o += sb;
  g->n_float = o;
  g->off_dstart = o; o += sb;
  g->off_noedge = o; o += sb;
  g->off_pcn = o; o += sb;
  g->off_events = o; o += n_state * n_event * nb;
  g->total = o;
  return true;
}

// largest partition point <= x (binary search; partitions need not be regular)
MSIM_HD int summary_bin(const SummaryGeom &g, double x) {
  int lo = 0, hi = g.n_bin - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (g.part[mid] <= x) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// discountedUtilityInterval / discountedCostInterval (microsimulation.h:1146-1162), exponentials passed in
MSIM_HD double summary_discounted(double w, double rate, double inv_alpha, double a, double b, double ea, double eb) {
  if (a == b || w == 0.0) return 0.0;
  if (rate == 0.0) return w * (b - a);
  return w * inv_alpha * (ea - eb);
}

// SummaryReport::add(state, event, lhs, rhs, index) with the report's current
// utility u and cost c.  Returns the interval's discounted utility and cost
// through *du, *dc (the caller owns the per-individual totals).
template <class Acc>
MSIM_DEV void summary_add(const SummaryGeom &g, Acc &acc, int state, int event, double lhs, double rhs, double u, double c,
                          double *du, double *dc) {
  const int nb = g.n_bin;
  const int lo = summary_bin(g, lhs), hi = summary_bin(g, rhs);
  const bool same = (g.u_rate == g.c_rate);
  const double eul = g.u_rate != 0.0 ? exp(-lhs * g.u_alpha) : 1.0, eur = g.u_rate != 0.0 ? exp(-rhs * g.u_alpha) : 1.0;
  const double ecl = same ? eul : (g.c_rate != 0.0 ? exp(-lhs * g.c_alpha) : 1.0);
  const double ecr = same ? eur : (g.c_rate != 0.0 ? exp(-rhs * g.c_alpha) : 1.0);
  *du = summary_discounted(u, g.u_rate, g.u_inv_alpha, lhs, rhs, eul, eur);
  *dc = summary_discounted(c, g.c_rate, g.c_inv_alpha, lhs, rhs, ecl, ecr);
  const int row = state * nb;
  const double plo = g.part[lo];
  const bool on_edge = lhs <= plo;
  acc.addi(g.off_events + (state * g.n_event + event) * nb + hi, 1);
  int f0 = on_edge ? lo : lo + 1;
  if (f0 > hi) f0 = hi;
  acc.addc(g.off_dstart + row + f0);
  if (f0 < hi) {
    if (u != 0.0) {
      acc.addf(g.off_wu0 + row + f0, u);
      acc.addf(g.off_wu1 + row + hi, u);
    }
    if (c != 0.0) {
      acc.addf(g.off_wc0 + row + f0, c);
      acc.addf(g.off_wc1 + row + hi, c);
    }
  }
  if (lo == hi) {
    const double a = on_edge ? plo : lhs;
    acc.addf(g.off_pt + row + hi, rhs - a);
    acc.addf(g.off_ut + row + hi, summary_discounted(u, g.u_rate, g.u_inv_alpha, a, rhs, on_edge ? g.u_decay[lo] : eul, eur));
    acc.addf(g.off_cost + row + hi, summary_discounted(c, g.c_rate, g.c_inv_alpha, a, rhs, on_edge ? g.c_decay[lo] : ecl, ecr));
    if (!(on_edge && plo < rhs)) acc.addi(g.off_noedge + row + hi, 1);
    return;
  }
  if (!on_edge) {
    const double nxt = g.part[lo + 1];
    acc.addf(g.off_pt + row + lo, nxt - lhs);
    acc.addf(g.off_ut + row + lo, summary_discounted(u, g.u_rate, g.u_inv_alpha, lhs, nxt, eul, g.u_decay[lo + 1]));
    acc.addf(g.off_cost + row + lo, summary_discounted(c, g.c_rate, g.c_inv_alpha, lhs, nxt, ecl, g.c_decay[lo + 1]));
  }
  const double phi = g.part[hi];
  acc.addf(g.off_pt + row + hi, rhs - phi);
  acc.addf(g.off_ut + row + hi, summary_discounted(u, g.u_rate, g.u_inv_alpha, phi, rhs, g.u_decay[hi], eur));
  acc.addf(g.off_cost + row + hi, summary_discounted(c, g.c_rate, g.c_inv_alpha, phi, rhs, g.c_decay[hi], ecr));
  if (!(phi < rhs)) acc.addi(g.off_noedge + row + hi, 1);
}

// SummaryReport::addPointCost(state, cost) at time `now` (microsimulation.h:1109-1116); returns the discounted cost
template <class Acc>
MSIM_DEV double summary_point_cost(const SummaryGeom &g, Acc &acc, int state, double now, double cost) {
  double v = 0.0;
  if (cost != 0.0) v = (g.c_rate == 0.0) ? cost : cost * exp(-now * g.c_alpha);  // discountedCost, :1163-1172
  const int b = summary_bin(g, now);
  acc.addf(g.off_cost + state * g.n_bin + b, v);
  acc.addi(g.off_pcn + state * g.n_bin + b, 1);
  return v;
}

// CostReport::add(state, time, cost) (microsimulation.h:1265-1278): a point cost discounted as cost/(1+rate)^time
// (pow, not the exp form SummaryReport uses) into the bin of `time`.  Returns what the reference adds to the
// individual's running total `current` (discounted from startReportAge on).  Uses the cost cells of a SummaryGeom.
template <class Acc>
MSIM_DEV double cost_report_add(const SummaryGeom &g, Acc &acc, int state, double time, double cost, double rate,
                                double startReportAge) {
  const double c = (rate == 0.0) ? cost : cost / pow(1 + rate, time);
  const int b = summary_bin(g, time);
  acc.addf(g.off_cost + state * g.n_bin + b, c);
  acc.addi(g.off_pcn + state * g.n_bin + b, 1);
  if (startReportAge == 0.0) return c;
  if (time >= startReportAge) return (rate == 0.0) ? cost : cost / pow(1 + rate, time - startReportAge);
  return 0.0;
}

// The free functions models use for their own running totals (microsimulation.h:811-836; e.g. the
// illness-death doc example, inst/doc/examples.org:300-336): integral of y*(1+rate)^(-u) over [start, end],
// and y/(1+rate)^time.  Written with pow, as the reference writes them.
MSIM_HD double discountedInterval(double y, double start, double end, double discountRate) {
  if (discountRate == 0.0) return y * (end - start);
  return y * (pow(1.0 + discountRate, -start) - pow(1.0 + discountRate, -end)) / log(1.0 + discountRate);
}
MSIM_HD double discountedInterval(double start, double end, double discountRate) {
  if (discountRate == 0.0) return end - start;
  return (pow(1.0 + discountRate, -start) - pow(1.0 + discountRate, -end)) / log(1.0 + discountRate);
}
MSIM_HD double discountedPoint(double y, double time, double discountRate) {
  return discountRate <= 0.0 ? y : y * pow(1.0 + discountRate, -time);
}
MSIM_HD double discountedPoint(double time, double discountRate) {
  return discountRate <= 0.0 ? 1.0 : pow(1.0 + discountRate, -time);
}

}  // namespace msim
