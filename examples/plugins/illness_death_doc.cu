// The reference's documented user model, as an out-of-tree plugin.
//
// inst/doc/examples.org:254-312 defines (inline, through Rcpp::sourceCpp) a cProcess subclass `illness_death` with
// running totals of discounted QALYs and costs, competing exponential events that are cancelled at every transition,
// a stop() hook that appends the totals to the model's own report, and an exported driver that runs n individuals
// from R's random number stream.  This file is that model written against plugin.cuh: the handler bodies are the
// reference's, statement for statement; event names became an enum, `report["QALY"].push_back(x)` became
// env.value(0, x).  The printed results of the reference's run (set.seed(12345), n = 5; examples.org:318-332) are the
// known answers tests/test_plugin.py checks.
#include "plugin.cuh"

namespace doc {

enum event_t { onset, recovered, death };

struct Consts {
  double mean_onset, mean_death_healthy, mean_recovery, mean_death_ill;  // 30, 60, 5, 10 in the document
  double discount_rate;                                                  // 0.03
};

// every transition cancels the pending events lazily (they stay in the heap until their time comes) and schedules two
// new ones: after k transitions up to k + 2 actions are pending.  P(k >= 62) is about 1e-11 at the document's rates;
// a person who got there would end the run with MSIM_ERR_HEAP, not with a wrong answer.
#ifndef DOC_HEAP_CAP
#define DOC_HEAP_CAP 64
#endif
constexpr int HEAP_CAP = DOC_HEAP_CAP;  // (the tests build a copy with 3 to see the error code)

template <class Env>
struct illness_death : msim::cProcess<illness_death<Env>, typename Env::Src, HEAP_CAP> {
  MSIM_PLUGIN_MODEL(illness_death, Consts, HEAP_CAP)
  double QALY, utility, cumulative_costs, cost;

  MSIM_DEV void init() {  // examples.org:266-269
    cumulative_costs = QALY = 0.0;
    this->toHealthy();
  }
  MSIM_DEV void toHealthy() {  // :270-275
    utility = 0.95;
    cost = 1000.0;
    this->scheduleAt(now() + R.rexp(K.mean_onset), onset);
    this->scheduleAt(now() + R.rexp(K.mean_death_healthy), death);
  }
  MSIM_DEV void toIllness() {  // :276-281
    utility = 0.95 * 0.8;
    cost = 10000.0;
    this->scheduleAt(now() + R.rexp(K.mean_recovery), recovered);
    this->scheduleAt(now() + R.rexp(K.mean_death_ill), death);
  }
  MSIM_DEV void handleMessage(int kind) {  // :283-294
    QALY += msim::discountedInterval(utility, previous(), now(), K.discount_rate);
    cumulative_costs += msim::discountedInterval(cost, previous(), now(), K.discount_rate);
    this->cancel_events();  // in this case, we cancel all of the competing events
    if (kind == onset)
      toHealthy();  // (sic: the document's C++ version calls toHealthy() here)
    else if (kind == recovered)
      toIllness();
    else
      stop_process();
  }
  MSIM_DEV void stop() {  // :295-298
    env.value(0, QALY);
    env.value(1, cumulative_costs);
  }
};

struct Traits {
  typedef doc::Consts Consts;
  template <class Env> using Model = illness_death<Env>;
  static constexpr int N_VALUES = 2;
  static const char *const *value_names() {
    static const char *const names[2] = {"QALY", "cumulative_costs"};
    return names;
  }
};

}  // namespace doc

MSIM_PLUGIN(illness_death_doc, doc::Traits)
