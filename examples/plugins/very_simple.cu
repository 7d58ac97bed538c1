// The reference's engine-overhead benchmark as a plugin: `VerySimple` of src/person-r.cc:234-241 (callSpeedTest runs
// 10^6 of them): two events per process, an empty handler, no random numbers.  On the CPU it measures the bare cost
// of create_process / schedule / pop / dispatch / clear (allocation, virtual calls); here the event times are
// compile-time constants and nvcc folds the whole life away, so it is only the smallest possible plugin.  The one
// per-person value (events handled) exists so that the run has something to check.
#include "plugin.cuh"

struct Consts {
  int unused;
};

template <class Env>
struct VerySimple : msim::cProcess<VerySimple<Env>, typename Env::Src, 4> {
  MSIM_PLUGIN_MODEL(VerySimple, Consts, 4)
  int handled;
  MSIM_DEV void init() {
    handled = 0;
    scheduleAt(10.0, 0);  // "a message"
    scheduleAt(11.0, 1);  // "another message"
  }
  MSIM_DEV void handleMessage(int) { env.value(0, (double)++handled); }
};

struct Traits {
  typedef ::Consts Consts;
  template <class Env> using Model = VerySimple<Env>;
  static constexpr int N_VALUES = 1;
  static const char *const *value_names() {
    static const char *const names[1] = {"events"};
    return names;
  }
};

MSIM_PLUGIN(very_simple, Traits)
