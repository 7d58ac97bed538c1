// Dense, O(1)-per-event restatement of the reference's EventReport::add
// (inst/include/microsimulation.h:933-960) and of the interval part of
// SummaryReport::add (:1070-1107).
//
// The reference walks every age bin between the previous event and this one
// (~70 hash-map updates per person).  Here an interval [lhs, rhs] touches a
// fixed number of cells:
//   events[state][event][hi]            += 1                (hi = bin of rhs)
//   pt / ut partial sums at bins lo, hi += the two partial overlaps
//   touch[lo], touch[hi]                += 1   (the reference emits a pt/ut row
//                                               for every bin it visited, even
//                                               when it added 0.0)
//   dfull[lo+1] += 1, dfull[hi] -= 1           difference array: bins fully
//                                               covered; each adds exactly one
//                                               bin width (and its discounted
//                                               width) at finalisation
//   dprev[p0] += 1, dprev[p1+1] -= 1           difference array over the bins
//                                               whose left edge lies in [lhs, rhs)
// All cells live in one vector of doubles per report ("dense accumulators");
// counts are integer-valued doubles (< 2^53), so sums are exact and
// order-independent, and ONE all-reduce(sum, f64) merges shards.  Finalisation
// (prefix sums, widths, sparse rows in std::map order) happens once on the
// host: dense_to_report() in capi.cu.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mrg32k3a.cuh"

namespace msim {

constexpr int REPORT_MAX_BINS = 128;

struct ReportGeom {
  int n_state, n_event, n_bin;  // n_bin = number of partition points (bins are [part[i], part[i+1]))
  double part_start, part_delta;  // partition = start + i*delta for i < n_bin-1, then part_last
  double part_last;
  double discount_rate, alpha;  // alpha = log(1 + discount_rate), computed on the host
  // offsets (in doubles) into the dense vector
  int off_pt, off_ut, off_touch, off_dfull, off_dprev, off_events, total;
};

inline ReportGeom make_report_geom(int n_state, int n_event, int n_bin, double start, double delta, double last,
                                   double discount_rate) {
  ReportGeom g;
  g.n_state = n_state;
  g.n_event = n_event;
  g.n_bin = n_bin;
  g.part_start = start;
  g.part_delta = delta;
  g.part_last = last;
  g.discount_rate = discount_rate;
  g.alpha = log(1.0 + discount_rate);
  int sb = n_state * n_bin;
  g.off_pt = 0;
  g.off_ut = g.off_pt + sb;
  g.off_touch = g.off_ut + sb;
  g.off_dfull = g.off_touch + sb;
  g.off_dprev = g.off_dfull + n_state * (n_bin + 1);
  g.off_events = g.off_dprev + n_state * (n_bin + 1);
  g.total = g.off_events + n_state * n_event * n_bin;
  return g;
}

MSIM_HD double report_part(const ReportGeom &g, int i) {
  return (i >= g.n_bin - 1) ? g.part_last : g.part_start + i * g.part_delta;
}

// index of the largest partition point <= x (set<greater>::lower_bound, :937-938)
MSIM_HD int report_bin(const ReportGeom &g, double x) {
  int nb = g.n_bin;
  int i = (int)floor((x - g.part_start) / g.part_delta);
  if (i < 0) i = 0;
  if (i > nb - 2) i = nb - 2;
  while (i + 1 < nb && report_part(g, i + 1) <= x) i++;
  while (i > 0 && report_part(g, i) > x) i--;
  return i;
}

// EventReport::discountedUtilities (microsimulation.h:917-928)
MSIM_HD double report_discounted(const ReportGeom &g, double a, double b, double utility) {
  if (g.discount_rate == 0.0) return utility * (b - a);
  if (a == b) return 0.0;
  return utility / g.alpha * (exp(-a * g.alpha) - exp(-b * g.alpha));
}

// Acc supplies addf(index, double) and addi(index, int) on the dense vector
// (shared-memory atomics on the device).
template <class Acc>
MSIM_DEV void report_add(const ReportGeom &g, Acc &acc, int state, int event, double lhs, double rhs) {
  const int nb = g.n_bin;
  const bool disc = g.discount_rate != 0.0;  // rate 0: ut == pt, filled in at finalisation
  int lo = report_bin(g, lhs);
  int hi = report_bin(g, rhs);
  acc.addi(g.off_events + (state * g.n_event + event) * nb + hi, 1);
  const int row = state * nb;
  const int drow = state * (nb + 1);
  double plo = report_part(g, lo);
  if (lo == hi) {
    double a = lhs > plo ? lhs : plo;
    acc.addi(g.off_touch + row + lo, 1);
    acc.addf(g.off_pt + row + lo, rhs - a);
    if (disc) acc.addf(g.off_ut + row + lo, report_discounted(g, a, rhs, 1.0));
    if (lhs <= plo && plo < rhs) {  // cadlag prevalence at the bin's left edge
      acc.addi(g.off_dprev + drow + lo, 1);
      acc.addi(g.off_dprev + drow + lo + 1, -1);
    }
    return;
  }
  double a = lhs > plo ? lhs : plo;
  double nxt = report_part(g, lo + 1);
  double phi = report_part(g, hi);
  acc.addi(g.off_touch + row + lo, 1);
  acc.addf(g.off_pt + row + lo, nxt - a);
  if (disc) acc.addf(g.off_ut + row + lo, report_discounted(g, a, nxt, 1.0));
  acc.addi(g.off_touch + row + hi, 1);
  acc.addf(g.off_pt + row + hi, rhs - phi);
  if (disc) acc.addf(g.off_ut + row + hi, report_discounted(g, phi, rhs, 1.0));
  if (hi > lo + 1) {
    acc.addi(g.off_dfull + drow + lo + 1, 1);
    acc.addi(g.off_dfull + drow + hi, -1);
  }
  int p0 = (lhs <= plo) ? lo : lo + 1;
  int p1 = (phi < rhs) ? hi : hi - 1;
  if (p0 <= p1) {
    acc.addi(g.off_dprev + drow + p0, 1);
    acc.addi(g.off_dprev + drow + p1 + 1, -1);
  }
}

}  // namespace msim
