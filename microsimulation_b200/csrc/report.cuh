// Dense, O(1)-per-event restatement of the reference's EventReport::add
// (inst/include/microsimulation.h:933-960) and of the interval part of
// SummaryReport::add (:1070-1107).
//
// The reference walks every age bin between the previous event and this one
// (~70 hash-map updates per person).  Here an interval [lhs, rhs] in state s
// ending with event e touches a fixed, small number of cells.  With
// lo = bin(lhs), hi = bin(rhs), plo/phi the left edges of those bins,
// on_edge = (lhs <= plo)  (the interval starts exactly on a partition point,
// e.g. every person's first interval starts at age 0):
//
//   events[s][e][hi] += 1
//   dstart[s][f0]    += 1     f0 = first bin the interval covers COMPLETELY
//                             (lo if on_edge else lo+1), clipped to hi.
//        The number of intervals covering bin b completely is then
//        full[b] = sum_{j<=b} dstart[s][j] - sum_{j<=b} sum_e events[s][e][j]
//        (each interval stops covering whole bins at its own hi), and every
//        such interval adds exactly one bin width (and its discounted width).
//   pt/ut[s][lo]     += the partial overlap with bin lo   (only if !on_edge, lo < hi)
//   pt/ut[s][hi]     += the partial overlap with bin hi
//   noedge[s][hi]    += 1     only in the RARE case that the interval does not
//                             contain hi's left edge (phi < lhs, or phi == rhs);
//        prevalence (cadlag, lhs <= edge < rhs) is then
//        prev[b] = full[b] + sum_e events[s][e][b] - noedge[s][b].
//   The reference emits a pt/ut row for every bin an interval visited, even
//   with value 0.0: bin b has a row iff full[b] > 0, or an interval ended in it
//   (events > 0), or it received a lo-side partial (always > 0).
//
// So a person's first interval costs 2 counter updates + 2 FP sums; exp() is
// needed once per event (exp(-alpha*rhs); the left edges come from a table and
// exp(-alpha*lhs) is the previous event's value).  All cells live in ONE
// vector of doubles ("dense accumulators"); counters are integer-valued
// doubles (< 2^53), so their sums are exact and order-independent, and ONE
// all-reduce(sum, f64) merges shards.  Finalisation (prefix sums, widths,
// sparse rows in std::map order) happens once on the host: report_host.h.
#pragma once
#include <math.h>
#include <stdint.h>

#include "fastmath.cuh"
#include "mrg32k3a.cuh"

namespace msim {

constexpr int REPORT_MAX_BINS = 128;

struct ReportGeom {
  int n_state, n_event, n_bin;    // n_bin = number of partition points (bins are [part[i], part[i+1]))
  double part_start, part_delta;  // partition = start + i*delta for i < n_bin-1, then part_last
  double part_inv_delta;          // 1/delta: first guess of a bin index (corrected by comparison)
  double part_last;
  int unit;  // 1: partition {0, 1, 2, ..., n_bin-2, part_last}: a bin index is just the integer part
  int bin0;  // the bin of time 0, where every person's first interval starts (cProcess::startTime)
  // EventReport(discountRate, outputUtilities, size, startReportAge, indiv) (microsimulation.h:873-877): every
  // individual's discounted total counts from startReportAge on (addBrief, :929-932); sra_gain = exp(alpha * sra)
  double start_report_age, sra_gain;
  int output_utilities;
  double part[REPORT_MAX_BINS];  // the partition points (setPartition), ascending; entries >= n_bin repeat the last
  double discount_rate, alpha, inv_alpha;  // alpha = log(1 + discount_rate), inv_alpha = 1/alpha (host libm)
  // offsets (in doubles) into the dense vector; FP sums first, then counters
  int off_pt, off_ut, off_dstart, off_noedge, off_events, total;
  double nb2;  // (double)(n_bin - 2), for report_bin (an integer-to-double conversion per call otherwise)
  // exp(-alpha*part[i]) for every partition point, evaluated once on the host
  double edge_decay[REPORT_MAX_BINS];
};

inline ReportGeom make_report_geom(int n_state, int n_event, int n_bin, double start, double delta, double last,
                                   double discount_rate) {
  ReportGeom g;
  g.n_state = n_state;
  g.n_event = n_event;
  g.n_bin = n_bin;
  g.nb2 = (double)(n_bin - 2);
  g.part_start = start;
  g.part_delta = delta;
  g.part_inv_delta = 1.0 / delta;
  g.part_last = last;
  g.unit = (start == 0.0 && delta == 1.0 && last >= (double)(n_bin - 1)) ? 1 : 0;
  g.discount_rate = discount_rate;
  g.alpha = log(1.0 + discount_rate);
  g.inv_alpha = discount_rate != 0.0 ? 1.0 / g.alpha : 0.0;
  int sb = n_state * n_bin;
  g.off_pt = 0;
  g.off_ut = g.off_pt + sb;
  g.off_dstart = g.off_ut + sb;
  g.off_noedge = g.off_dstart + sb;
  g.off_events = g.off_noedge + sb;
  g.total = g.off_events + n_state * n_event * n_bin;
  for (int i = 0; i < REPORT_MAX_BINS; i++) {
    double p = (i >= n_bin - 1) ? last : start + i * delta;
    g.edge_decay[i] = (discount_rate != 0.0) ? fm_exp(-p * g.alpha) : 1.0;  // the function the events use (report_decay)
  }
  for (int i = 0; i < REPORT_MAX_BINS; i++) g.part[i] = (i >= n_bin - 1) ? last : start + i * delta;
  g.bin0 = 0;
  while (g.bin0 + 1 < n_bin && g.part[g.bin0 + 1] <= 0.0) g.bin0++;
  g.start_report_age = 0.0;
  g.sra_gain = 1.0;
  g.output_utilities = 1;
  return g;
}

// EventReport::setPartition(start, finish, delta, maxTime) (microsimulation.h:879-884): the points by REPEATED
// addition, as the reference builds them (start + i*delta can differ in the last bit), then maxTime; plus the
// constructor's startReportAge and outputUtilities.  Returns false if the partition does not fit the tables.
inline bool make_report_geom_partition(ReportGeom *out, int n_state, int n_event, double start, double finish, double delta,
                                       double max_time, double discount_rate, double start_report_age, int output_utilities) {
  if (!(delta > 0) || !(finish >= start) || !(max_time > finish)) return false;
  double pts[REPORT_MAX_BINS];
  int n = 0;
  for (double t = start; t <= finish; t += delta) {
    if (n >= REPORT_MAX_BINS - 1) return false;
    pts[n++] = t;
  }
  pts[n++] = max_time;
  ReportGeom g = make_report_geom(n_state, n_event, n, start, delta, max_time, discount_rate);
  bool unit = (start == 0.0 && delta == 1.0);
  for (int i = 0; i < REPORT_MAX_BINS; i++) {
    g.part[i] = pts[i < n ? i : n - 1];
    g.edge_decay[i] = (discount_rate != 0.0) ? fm_exp(-g.part[i] * g.alpha) : 1.0;
  }
  g.unit = unit ? 1 : 0;
  g.bin0 = 0;
  while (g.bin0 + 1 < n && g.part[g.bin0 + 1] <= 0.0) g.bin0++;
  g.start_report_age = start_report_age;
  g.sra_gain = (discount_rate != 0.0) ? fm_exp(start_report_age * g.alpha) : 1.0;
  g.output_utilities = output_utilities ? 1 : 0;
  *out = g;
  return true;
}

// partition point i.  Always from the table (one shared-memory load in the kernels): computing it for the built
// models' unit partition — (double)i, or `last` for the final point — is an integer conversion, a compare and a
// select, three times per event: 4.77 against 4.63 ms per 1e8 persons in the person-r kernel.
MSIM_HD double report_part(const ReportGeom &g, int i) { return g.part[i]; }

// (The (int)x below stays a conversion instruction: x + 1.5 * 2^52, the low word and a comparison — floor without the
// conversion — measured slower, 4.63 against 4.61 ms per 1e8 persons; the conversions of constants did pay: nb2, part[].)
// index of the largest partition point <= x (set<greater>::lower_bound, :937-938).  UNIT: the caller knows that
// g.unit is set (a kernel instantiated for the built models' partition: the general search, never executed there, is
// then not in its code at all — 230 instructions of the person-r kernel's loop, 2% of its time).
template <bool UNIT = false>
MSIM_HD int report_bin(const ReportGeom &g, double x) {
  const int nb = g.n_bin;
  if (UNIT || g.unit) {  // every built model's partition {0, 1, ..., 100, 1e6}: exact without a search
    if (x >= g.part_last) return nb - 1;
    return (x > 0.0) ? ((x < g.nb2) ? (int)x : nb - 2) : 0;
  }
  if (UNIT) return 0;  // (not reached)
  double y = (x - g.part_start) * g.part_inv_delta;
  int i = (y > 0.0) ? ((y < g.nb2) ? (int)y : nb - 2) : 0;
  while (i + 1 < nb && g.part[i + 1] <= x) i++;
  while (i > 0 && g.part[i] > x) i--;
  return i;
}

// exp(-alpha*t): the one transcendental an event needs when utilities are discounted
MSIM_HD double report_decay(const ReportGeom &g, double t) { return fm_exp(-t * g.alpha); }

// EventReport::discountedUtilities (microsimulation.h:917-928) with the two
// exponentials passed in: utility/alpha*(exp(-a*alpha) - exp(-b*alpha))
MSIM_HD double report_discounted_e(const ReportGeom &g, double a, double b, double ea, double eb) {
  if (a == b) return 0.0;
  return g.inv_alpha * (ea - eb);
}
MSIM_HD double report_discounted(const ReportGeom &g, double a, double b, double utility) {
  if (g.discount_rate == 0.0) return utility * (b - a);
  if (a == b) return 0.0;
  return utility / g.alpha * (exp(-a * g.alpha) - exp(-b * g.alpha));
}

// Which cells of the dense report a model can reach: a kernel that keeps its block-level tables in shared memory
// only needs rows for the states a person can be in and for the (state, event) pairs that can occur (person-r:
// 4 of 8 states, 11 of 64 pairs — 8 KB of counters instead of 33 KB).  Built on the host from the model's list;
// a pair outside the list raises an error on the device instead of being lost.
constexpr int REPORT_MAX_STATES = 16, REPORT_MAX_PAIRS = 256, REPORT_MAX_USED_PAIRS = 32;
struct ReportCompact {
  signed char state_slot[REPORT_MAX_STATES];  // state -> row of the compact tables, -1 = not reachable
  signed char pair_slot[REPORT_MAX_PAIRS];    // state * n_event + event -> row of the compact event table
  unsigned char slot_state[REPORT_MAX_STATES];
  unsigned char pair_state[REPORT_MAX_USED_PAIRS], pair_event[REPORT_MAX_USED_PAIRS];
  int n_states, n_pairs;
};

// pairs = {state, event, state, event, ...}; returns false if the geometry is too large for the tables
inline bool make_report_compact(const ReportGeom &g, const int *pairs, int n_pairs, ReportCompact *out) {
  ReportCompact c;
  for (int i = 0; i < REPORT_MAX_STATES; i++) c.state_slot[i] = -1, c.slot_state[i] = 0;
  for (int i = 0; i < REPORT_MAX_PAIRS; i++) c.pair_slot[i] = -1;
  for (int i = 0; i < REPORT_MAX_USED_PAIRS; i++) c.pair_state[i] = c.pair_event[i] = 0;
  c.n_states = c.n_pairs = 0;
  if (g.n_state > REPORT_MAX_STATES || g.n_state * g.n_event > REPORT_MAX_PAIRS || n_pairs > REPORT_MAX_USED_PAIRS) return false;
  for (int k = 0; k < n_pairs; k++) {
    const int st = pairs[2 * k], ev = pairs[2 * k + 1];
    if (st < 0 || st >= g.n_state || ev < 0 || ev >= g.n_event) return false;
    if (c.state_slot[st] < 0) {
      c.slot_state[c.n_states] = (unsigned char)st;
      c.state_slot[st] = (signed char)c.n_states++;
    }
    if (c.pair_slot[st * g.n_event + ev] < 0) {
      c.pair_state[c.n_pairs] = (unsigned char)st;
      c.pair_event[c.n_pairs] = (unsigned char)ev;
      c.pair_slot[st * g.n_event + ev] = (signed char)c.n_pairs++;
    }
  }
  *out = c;
  return true;
}

// What a process remembers between two of its report() calls: the previous
// interval's right end, its bin and exp(-alpha*rhs) are this interval's left
// end (previousEventTime), so neither the bin search nor the exponential is
// repeated.  `valid` is cleared at the start of every person.
struct ReportCursor {
  double t, e;
  int bin;
  bool valid;
};

// A person's first interval starts at startTime = 0 (microsimulation.h:430): its bin is known and
// exp(-alpha * 0) = 1 exactly, so the cursor can start out valid and the first event needs no second bin search
// and no second exponential.
MSIM_HD void report_cursor_start(const ReportGeom &g, ReportCursor &cur) {
  cur.t = 0.0;
  cur.e = 1.0;
  cur.bin = g.bin0;
  cur.valid = true;
}

// The cells one interval touches (see the header comment): what report_add
// adds, separated from HOW it is added so that a warp can add its lanes'
// intervals together (person_staged.cuh).
struct ReportCells {
  int state, event, bin_hi, bin_f0;  // the same cells by coordinates, for accumulators with their own layout
  int i_events, i_dstart;  // counters to increment
  double e_lhs, e_rhs;     // exp(-alpha * lhs), exp(-alpha * rhs) (1 without discounting): for addBrief
  bool noedge;             // also increment noedge[s][hi]
  int row_hi, row_lo;      // state*n_bin + bin of the partial overlaps; row_lo < 0: none
  double pt_hi, ut_hi, pt_lo, ut_lo;
};

// UNIT: instantiated for the built models' partition (report_bin), by a caller that has made sure the interval starts
// where the process's previous one ended (cur.valid && cur.t == lhs: true of every interval of a model that reports
// from previous() to now()), so that the left end's bin search and exponential are not in its code either.
template <bool UNIT = false>
MSIM_HD ReportCells report_cells(const ReportGeom &g, ReportCursor &cur, int state, int event, double lhs, double rhs) {
  const int nb = g.n_bin;
  const bool disc = g.discount_rate != 0.0;  // rate 0: ut == pt, filled in at finalisation
  ReportCells c;
  int lo;
  double elhs = 1.0;
  if (UNIT || (cur.valid && cur.t == lhs)) {
    lo = cur.bin;
    elhs = cur.e;
  } else {
    lo = report_bin<UNIT>(g, lhs);
    if (disc) elhs = report_decay(g, lhs);
  }
  const int hi = report_bin<UNIT>(g, rhs);
  const double erhs = disc ? report_decay(g, rhs) : 1.0;
  cur.t = rhs;
  cur.e = erhs;
  cur.bin = hi;
  cur.valid = true;
  c.e_lhs = elhs;
  c.e_rhs = erhs;

  const int row = state * nb;
  const double plo = report_part(g, lo);
  const bool on_edge = lhs <= plo;
  c.i_events = g.off_events + (state * g.n_event + event) * nb + hi;
  int f0 = on_edge ? lo : lo + 1;
  if (f0 > hi) f0 = hi;
  c.state = state;
  c.event = event;
  c.bin_hi = hi;
  c.bin_f0 = f0;
  c.i_dstart = g.off_dstart + row + f0;
  c.row_hi = row + hi;
  c.row_lo = -1;
  c.pt_lo = c.ut_lo = 0.0;
  c.ut_hi = 0.0;
  if (lo == hi) {
    const double a = on_edge ? plo : lhs;
    c.pt_hi = rhs - a;
    if (disc) c.ut_hi = report_discounted_e(g, a, rhs, on_edge ? g.edge_decay[lo] : elhs, erhs);
    c.noedge = !(on_edge && plo < rhs);
    return c;
  }
  if (!on_edge) {
    const double nxt = report_part(g, lo + 1);
    c.row_lo = row + lo;
    c.pt_lo = nxt - lhs;
    if (disc) c.ut_lo = report_discounted_e(g, lhs, nxt, elhs, g.edge_decay[lo + 1]);
  }
  const double phi = report_part(g, hi);
  c.pt_hi = rhs - phi;
  if (disc) c.ut_hi = report_discounted_e(g, phi, rhs, g.edge_decay[hi], erhs);
  c.noedge = !(phi < rhs);
  return c;
}

// EventReport::addBrief (microsimulation.h:929-932): what the interval [lhs, rhs] adds to the individual's running
// discounted total, which counts from startReportAge on: discountedUtilities(max(lhs - sra, 0), rhs - sra) with
// utility 1.  exp(-alpha (t - sra)) = exp(-alpha t) exp(alpha sra): the interval's own exponentials are reused.
MSIM_HD double report_brief(const ReportGeom &g, const ReportCells &c, double lhs, double rhs) {
  const double sra = g.start_report_age;
  if (!(rhs >= sra)) return 0.0;
  const double a = (lhs > sra) ? lhs - sra : 0.0, b = rhs - sra;
  if (g.discount_rate == 0.0) return b - a;
  if (a == b) return 0.0;
  const double ea = (lhs > sra) ? c.e_lhs * g.sra_gain : 1.0;
  return g.inv_alpha * (ea - c.e_rhs * g.sra_gain);
}

// Acc supplies addf(index, double) for FP sums, addi(index, int) for counters
// and addc(index) for the `dstart` counter (+1; consecutive calls usually name
// the same cell, which an accumulator may exploit) on the dense vector.
template <class Acc>
MSIM_DEV void report_add(const ReportGeom &g, Acc &acc, ReportCursor &cur, int state, int event, double lhs, double rhs) {
  const ReportCells c = report_cells(g, cur, state, event, lhs, rhs);
  const bool disc = g.discount_rate != 0.0;
  acc.addi(c.i_events, 1);
  acc.addc(c.i_dstart);
  if (c.row_lo >= 0) {
    acc.addf(g.off_pt + c.row_lo, c.pt_lo);
    if (disc) acc.addf(g.off_ut + c.row_lo, c.ut_lo);
  }
  acc.addf(g.off_pt + c.row_hi, c.pt_hi);
  if (disc) acc.addf(g.off_ut + c.row_hi, c.ut_hi);
  if (c.noedge) acc.addi(g.off_noedge + c.row_hi, 1);
}

}  // namespace msim
