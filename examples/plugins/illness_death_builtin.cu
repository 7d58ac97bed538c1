// The package's own illness-death model (src/illness-death.cpp; microsimulation_b200/csrc/models.cuh) exported through
// the plugin mechanism: a model defined for the built-in entry points is a plugin model as it stands.  The tests use
// this to check the plugin path's EventReport against msim_callIllnessDeath, table for table.
#include "plugin.cuh"

struct Traits {
  typedef msim::illness_death::Consts Consts;
  template <class Env> using Model = msim::illness_death::SimplePerson<Env>;
  static constexpr int N_VALUES = 0;
  static const char *const *value_names() { return nullptr; }
};

MSIM_PLUGIN(illness_death_builtin, Traits)

// b_weibull and friends are evaluated on the host, as the reference does (illness-death.cpp:17-19, 37)
MSIM_PLUGIN_EXPORT void illness_death_builtin_consts(double cure, double zsd, void *out) {
  msim::illness_death::Consts c = msim_host::illness_consts(cure, zsd);
  memcpy(out, &c, sizeof c);
}
