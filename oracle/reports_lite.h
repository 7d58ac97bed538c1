// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Restatement of the reference's report classes (Rcpp-dependent, so they
// cannot be compiled here as they lie):
//   EventReport    inst/include/microsimulation.h:860-1005
//   SummaryReport  inst/include/microsimulation.h:1010-1190
//   CostReport     inst/include/microsimulation.h:1195-1296
// As in the reference the tables are std::unordered_map keyed by pair / tuple with
// the reference's hash (boost-style hash_combine over the members,
// microsimulation.h:41-98) — the containers are where the reference's CPU time
// goes for report models, so the timed baseline uses the same ones — and they are
// sorted into std::map order when "wrapped" (flattened to columns here), as
// wrap_map does (microsimulation.h:1843-1860).
#ifndef MSIM_ORACLE_REPORTS_LITE_H
#define MSIM_ORACLE_REPORTS_LITE_H

#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "cprocess_lite.h"

namespace ssim {

// microsimulation.h:50-98: seed ^= hash(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2), member after member from seed 0
template <class T>
inline void hash_combine(std::size_t &seed, T const &v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
struct KeyHash {
  template <class A, class B>
  std::size_t operator()(std::pair<A, B> const &p) const {
    std::size_t seed = 0;
    hash_combine(seed, p.first);
    hash_combine(seed, p.second);
    return seed;
  }
  template <class A, class B, class C>
  std::size_t operator()(std::tuple<A, B, C> const &t) const {
    std::size_t seed = 0;
    hash_combine(seed, std::get<0>(t));
    hash_combine(seed, std::get<1>(t));
    hash_combine(seed, std::get<2>(t));
    return seed;
  }
};

// microsimulation.h:468-496
class Means {
 public:
  Means() : _n(0), _sum(0.0), _sumsq(0.0) {}
  double mean() { return _sum / _n; }
  double var() { return (long double)_n / (_n - 1) * (_sumsq / _n - mean() * mean()); }
  int n() { return _n; }
  double sum() { return _sum; }
  double sumsq() { return _sumsq; }
  double sd() { return sqrt(var()); }
  Means *operator+=(const double value) {
    _n++;
    _sum += (long double)value;
    _sumsq += (long double)value * (long double)value;
    return this;
  }

 private:
  int _n;
  long double _sum, _sumsq;
};

template <class State, class Event, class Time = double, class Utility = double>
class EventReport {
 public:
  typedef std::set<Time, std::greater<Time> > Partition;  // descending
  typedef typename Partition::iterator Iterator;
  typedef std::pair<State, Time> Pair;
  typedef std::tuple<State, Event, Time> Triple;
  EventReport(Utility discountRate = 0.0, bool outputUtilities = true, Time startReportAge = Time(0), int size = 1,
              bool indiv = false)
      : discountRate(discountRate), current(0), outputUtilities(outputUtilities),
        startReportAge(startReportAge), id(0), indiv(indiv) {
    _vector.resize(size);
    setPartition(startReportAge);
  }
  void setIndivN(const int n) {  // microsimulation.h:889-892
    _vector.resize(n);
    indiv = true;
  }
  void individualReset() {  // microsimulation.h:905-916
    if (now() >= startReportAge) mean_utilities += double(current);
    if (indiv && id < int(_vector.size())) _vector[id] = (now() >= startReportAge) ? double(current) : NAN;  // NA_REAL
    current = Utility(0);
    id++;
  }
  void setPartition(const std::vector<Time> v) {
    for (size_t i = 0; i < v.size(); i++) _partition.insert(v[i]);
  }
  void setPartition(const Time start = 0.0, const Time finish = 100.0, const Time delta = 1.0,
                    const Time maxTime = Time(1.0e100)) {
    _partition.clear();
    for (Time t = start; t <= finish; t += delta) _partition.insert(t);
    _partition.insert(maxTime);
  }
  void clear() {
    _ut.clear();
    _pt.clear();
    _events.clear();
    _prev.clear();
    _partition.clear();
    current = Utility(0);
  }
  Utility discountedUtilities(Time a, Time b, Utility utility = 1.0) {
    if (discountRate == 0.0) return utility * (b - a);
    else if (a == b) return 0.0;
    else if (discountRate > 0.0) {
      double alpha = log(1.0 + discountRate);
      return utility / alpha * (exp(-a * alpha) - exp(-b * alpha));
    }
    return 0.0;
  }
  void addBrief(const Time lhs, const Time rhs, const Utility utility = 1) {
    if (rhs >= startReportAge)
      current += discountedUtilities(std::max<Time>(lhs - startReportAge, Time(0)), rhs - startReportAge, utility);
  }
  // microsimulation.h:933-960.  With a descending set, lower_bound(x) is the
  // largest partition point <= x; "--it" moves to the next LARGER point.
  void add(const State state, const Event event, const Time lhs, const Time rhs, const Utility utility = 1) {
    addBrief(lhs, rhs, utility);
    Iterator lo = _partition.lower_bound(lhs);
    Iterator hi = _partition.lower_bound(rhs);
    Iterator last = _partition.begin();
    ++_events[Triple(state, event, *hi)];
    Iterator it = lo;
    for (;;) {
      if (lhs <= (*it) && (*it) < rhs) ++_prev[Pair(state, *it)];
      if (it == last) {
        if (outputUtilities) _ut[Pair(state, *it)] += discountedUtilities(std::max<Time>(lhs, *it), rhs, utility);
        _pt[Pair(state, *it)] += rhs - std::max<Time>(lhs, *it);
      } else {
        Iterator nx = it;
        --nx;
        Time next_value = *nx;
        if (outputUtilities)
          _ut[Pair(state, *it)] +=
              discountedUtilities(std::max<Time>(lhs, *it), std::min<Time>(rhs, next_value), utility);
        _pt[Pair(state, *it)] += std::min<Time>(rhs, next_value) - std::max<Time>(lhs, *it);
      }
      if (it == hi) break;
      --it;
    }
  }
  Utility discountRate, current;
  bool outputUtilities;
  Partition _partition;
  std::unordered_map<Pair, int, KeyHash> _prev;
  std::unordered_map<Pair, Utility, KeyHash> _ut;
  std::unordered_map<Pair, Time, KeyHash> _pt;
  std::unordered_map<Triple, int, KeyHash> _events;
  Time startReportAge;
  std::vector<Utility> _vector;
  Means mean_utilities;
  int id;
  bool indiv;
};

template <class State, class Event = short, class Time = double, class Utility = double, class Cost = double>
class SummaryReport {
 public:
  typedef std::set<Time, std::greater<Time> > Partition;
  typedef typename Partition::iterator Iterator;
  typedef std::pair<State, Time> Pair;
  typedef std::tuple<State, Event, Time> Triple;
  SummaryReport(int n = 1, bool indivp = true, Utility discountRate = 0.0) : n(n), indivp(indivp) {
    resize(indivp ? n : 1);
    setDiscountRate(discountRate);
    setUtility(1.0);
    setCost(0.0);
  }
  void resize(int size) {
    _indivUtilities.resize(size);
    _indivCosts.resize(size);
  }
  void setDiscountRate(Utility discountRate) {
    setUtilityDiscountRate(discountRate);
    setCostDiscountRate(discountRate);
  }
  void setUtilityDiscountRate(Utility discountRate) {
    utilityDiscountRate = discountRate;
    utilityAlpha = log(1.0 + utilityDiscountRate);
  }
  void setCostDiscountRate(Cost discountRate) {
    costDiscountRate = discountRate;
    costAlpha = log(1.0 + costDiscountRate);
  }
  void setPartition(const std::vector<Time> v) {
    for (size_t i = 0; i < v.size(); i++) _partition.insert(v[i]);
  }
  void setPartition(const Time start, const Time finish, const Time delta, const Time maxTime = Time(1.0e100)) {
    _partition.clear();
    for (Time t = start; t <= finish; t += delta) _partition.insert(t);
    _partition.insert(maxTime);
  }
  void setUtility(Utility u) { utility = u; }
  void setCost(Cost c) { cost = c; }
  void clear() {
    _ut.clear();
    _pt.clear();
    _events.clear();
    _prev.clear();
    _partition.clear();
    _costs.clear();
    _indivUtilities.clear();
    _indivCosts.clear();
  }
  // microsimulation.h:1070-1107
  void add(const State state, const Event event, const Time lhs, const Time rhs, int index = 0) {
    if (!indivp) index = 0;
    Iterator lo = _partition.lower_bound(lhs);
    Iterator hi = _partition.lower_bound(rhs);
    Iterator last = _partition.begin();
    ++_events[Triple(state, event, *hi)];
    Iterator it = lo;
    for (;;) {
      if (lhs <= (*it) && (*it) < rhs) ++_prev[Pair(state, *it)];
      Time a = std::max<Time>(lhs, *it), b;
      if (it == last) {
        b = rhs;
      } else {
        Iterator nx = it;
        --nx;
        b = std::min<Time>(rhs, *nx);
      }
      Utility u = discountedUtilityInterval(a, b, utility);
      _ut[Pair(state, *it)] += u;
      _indivUtilities[index] += u;
      Cost c = discountedCostInterval(a, b, cost);
      _costs[Pair(state, *it)] += c;
      _indivCosts[index] += c;
      _pt[Pair(state, *it)] += b - a;
      if (it == hi) break;
      --it;
    }
  }
  void addPointCost(const State state, const Cost cost, int index = 0) {  // microsimulation.h:1109-1116
    if (!indivp) index = 0;
    Time time = ssim::now();
    Time time_lhs = *_partition.lower_bound(time);
    Cost c = discountedCost(time, cost);
    _costs[Pair(state, time_lhs)] += c;
    _indivCosts[index] += c;
  }
  Utility discountedUtilityInterval(Time a, Time b, Utility utility) {
    if (a == b || utility == 0.0) return 0.0;
    else if (utilityDiscountRate == 0.0) return utility * (b - a);
    else if (utilityDiscountRate > 0.0) return utility / utilityAlpha * (exp(-a * utilityAlpha) - exp(-b * utilityAlpha));
    return 0.0;
  }
  Cost discountedCostInterval(Time a, Time b, Cost cost) {
    if (a == b || cost == 0.0) return 0.0;
    else if (costDiscountRate == 0.0) return cost * (b - a);
    else if (costDiscountRate > 0.0) return cost / costAlpha * (exp(-a * costAlpha) - exp(-b * costAlpha));
    return 0.0;
  }
  Cost discountedCost(Time a, Cost cost) {
    if (cost == 0.0) return 0.0;
    else if (costDiscountRate == 0.0) return cost;
    else if (costDiscountRate > 0) return cost * exp(-a * costAlpha);
    return 0;
  }
  int n;
  bool indivp;
  Partition _partition;
  std::unordered_map<Pair, int, KeyHash> _prev;
  std::unordered_map<Pair, Utility, KeyHash> _ut;
  std::unordered_map<Pair, Time, KeyHash> _pt;
  std::unordered_map<Triple, int, KeyHash> _events;
  std::unordered_map<Pair, Cost, KeyHash> _costs;
  std::vector<Utility> _indivUtilities;
  std::vector<Cost> _indivCosts;
  Utility utilityDiscountRate, utilityAlpha, utility;
  Cost costDiscountRate, costAlpha, cost;
};

template <class State, class Time = double, class Cost = double>
class CostReport {
 public:
  typedef std::set<Time, std::greater<Time> > Partition;
  typedef std::pair<State, Time> Pair;
  CostReport(Cost discountRate = 0, Time startReportAge = Time(0))
      : discountRate(discountRate), current(0), startReportAge(startReportAge) {}
  Cost discountedCost(Time a, Cost cost) {
    if (discountRate == 0.0) return cost;
    else if (discountRate > 0.0) return cost / pow(1 + discountRate, a);
    return 0;
  }
  void setPartition(const Time start, const Time finish, const Time delta, const Time maxTime = Time(1.0e100)) {
    _partition.clear();
    for (Time t = start; t <= finish; t += delta) _partition.insert(t);
    _partition.insert(maxTime);
  }
  void clear() {
    _table.clear();
    _partition.clear();
    current = Cost(0);
  }
  void add(const State state, const Time time, const Cost cost) {  // microsimulation.h:1265-1278
    Time time_lhs = *_partition.lower_bound(time);
    Cost c = discountedCost(time, cost);
    _table[Pair(state, time_lhs)] += c;
    if (startReportAge == 0.0) current += c;
    else if (time >= startReportAge) current += discountedCost(time - startReportAge, cost);
  }
  Cost discountRate, current;
  Partition _partition;
  std::unordered_map<Pair, Cost, KeyHash> _table;
  Time startReportAge;
};

}  // namespace ssim

#endif
