// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
// Plain containers the oracle models write into, and the run_* drivers.
#ifndef MSIM_ORACLE_MODELS_H
#define MSIM_ORACLE_MODELS_H

#include <stdint.h>

#include <map>
#include <vector>

#include "../include/msim.h"

namespace orc {

extern uint64_t g_events_handled;

struct RowLog {  // the reference's std::map<std::string, std::vector<double>> report
  std::vector<double> id, startTime, endTime, state, event;
  void push(double i, double s, double e, double st, double ev) {
    id.push_back(i);
    startTime.push_back(s);
    endTime.push_back(e);
    state.push_back(st);
    event.push_back(ev);
  }
};

struct ReportTables {  // flattened EventReport::wrap() / SummaryReport::asList()
  std::vector<double> pt, ut, events, prev, costs, indiv;  // row-major rows
  int64_t n_pt = 0, n_ut = 0, n_events = 0, n_prev = 0, n_costs = 0, n_indiv = 0;
};

struct SickSickerParam {  // README.org:148-187
  double r_HS1, r_HD, r_S1H, r_S1S2, r_S1D, r_S2D;
  double c_H, c_S1, c_S2, c_Trt, u_H, u_S1, u_S2, u_Trt;
  double discountRate, partitionBy;
  int Trt;
};

// wrap_map on an unordered_map: copied into a std::map first, so rows come out key-sorted (microsimulation.h:1843-1860)
template <class UMap3>
void flatten3(const UMap3 &um, std::vector<double> &rows, int64_t &n) {  // (state, age) -> value
  typedef std::map<typename UMap3::key_type, typename UMap3::mapped_type> Map3;
  const Map3 m(um.begin(), um.end());
  rows.clear();
  for (typename Map3::const_iterator it = m.begin(); it != m.end(); ++it) {
    rows.push_back((double)it->first.first);
    rows.push_back((double)it->first.second);
    rows.push_back((double)it->second);
  }
  n = (int64_t)m.size();
}
template <class UMap4>
void flatten4(const UMap4 &um, std::vector<double> &rows, int64_t &n) {  // (state, event, age) -> number
  typedef std::map<typename UMap4::key_type, typename UMap4::mapped_type> Map4;
  const Map4 m(um.begin(), um.end());
  rows.clear();
  for (typename Map4::const_iterator it = m.begin(); it != m.end(); ++it) {
    rows.push_back((double)std::get<0>(it->first));
    rows.push_back((double)std::get<1>(it->first));
    rows.push_back((double)std::get<2>(it->first));
    rows.push_back((double)it->second);
  }
  n = (int64_t)m.size();
}

// EventReport::wrap(): an empty list when no event was ever recorded
// (microsimulation.h:976)
template <class Rep>
void flatten_summary(const Rep &r, ReportTables *out) {
  *out = ReportTables();
  if (r._events.size() == 0) return;
  flatten3(r._pt, out->pt, out->n_pt);
  flatten3(r._ut, out->ut, out->n_ut);
  flatten4(r._events, out->events, out->n_events);
  flatten3(r._prev, out->prev, out->n_prev);
  flatten3(r._costs, out->costs, out->n_costs);
  out->n_indiv = (int64_t)r._indivUtilities.size();
  for (size_t i = 0; i < r._indivUtilities.size(); i++) {
    out->indiv.push_back(r._indivUtilities[i]);
    out->indiv.push_back(r._indivCosts[i]);
  }
}
template <class Rep>
void flatten(const Rep &r, ReportTables *out) {
  *out = ReportTables();
  if (r._events.size() == 0) return;
  flatten3(r._pt, out->pt, out->n_pt);
  if (r.outputUtilities) flatten3(r._ut, out->ut, out->n_ut);
  flatten4(r._events, out->events, out->n_events);
  flatten3(r._prev, out->prev, out->n_prev);
}

struct PersonReportOptions {  // include/msim.h msim_report_options
  double start, finish, delta, max_time, start_report_age;
  int32_t output_utilities;
};
void run_person_r(const double seed_in[6], int64_t first_person, int64_t n, bool keep_rows, RowLog *log,
                  double counts[9], ReportTables *report, double discount_rate, const PersonReportOptions *opt = 0,
                  double *indiv = 0, double means[7] = 0);
void run_calibration(int64_t first_person, int64_t n, const double par[6], msim_calibration *out);
void run_illness_death(int64_t n, double cure, double zsd, ReportTables *out);
void run_simple_person(int64_t n, RowLog *log);
void run_simple_person2(int64_t n, ReportTables *out);
void run_cancer_management(int64_t n, double odShape, double discountRate, double partitionBy, bool indivp, int64_t first_person,
                           ReportTables *out);
void run_sick_sicker(int64_t n, const SickSickerParam *param, bool indivp, bool substreams, int64_t first_person,
                     ReportTables *out, std::vector<double> *trace);
double run_tick_tock(std::vector<double> *times);
void run_doc_illness_death(int n, std::vector<double> *times, std::vector<double> *qaly_cost);
void run_priority_demo(std::vector<double> *order);

}  // namespace orc

#endif
