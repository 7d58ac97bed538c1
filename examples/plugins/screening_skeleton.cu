// A screening-model SKELETON as an out-of-tree plugin: the device building blocks of the external FHCRC model —
// Rpexp, Table, NumericInterpolate (lookups.cuh), a generalised survival model (gsm.cuh) and an EventReport — working
// TOGETHER in one kernel.  screening_skeleton.h says what the skeleton is made of and where each piece comes from in
// the reference tree; it is not the FHCRC model (that lives in mclements/prostata) and its results are pinned by
// nothing but the oracle's restatement of the same skeleton (tests/test_screening_skeleton.py).
//
// Besides the entry points MSIM_PLUGIN exports, the library has screening_skeleton_setup(): the look-up tables are
// uploaded once and stay on the device, in a __device__ struct the model reads (a plugin's parameter struct is copied
// by value into every launch; tables do not belong there).
#include "plugin.cuh"

#include "gsm.cuh"
#include "host_gsm.h"
#include "lookups.cuh"
#include "screening_skeleton.h"

namespace scr {

struct Tables {
  msim::Rpexp mortality;
  msim::DenseTable biopsy;
  msim::Interp sensitivity, survival;
  const msim::GsmModel *sojourn;
  double surv_last;  // the survival curve's last value: below it no cancer death is scheduled
};
__device__ Tables d_tables;

struct Consts {  // the scalar parameters of screening_inputs
  double g0, mubeta0, sebeta0, mubeta1, sebeta1, mubeta2[2], sebeta2[2], tau2;
  double psa_threshold, compliance, cohort, screen;
};

constexpr int HEAP_CAP = 8;  // pending at most: other death, onset or clinical dx, cancer death, screen, biopsy, screen dx

template <class Env>
struct Person : msim::cProcess<Person<Env>, typename Env::Src, HEAP_CAP> {
  MSIM_PLUGIN_MODEL(Person, Consts, HEAP_CAP)
  int state, grade;
  bool dx, screen_dx;
  double age_o, beta0, beta1, beta2, eps, age_dx;

  // R::rnormPos (src/microsimulation.cc:103-107)
  MSIM_DEV double rnormPos(double mean, double sd) {
    double x;
    while ((x = R.rnorm(mean, sd)) < 0.0) {
    }
    return x;
  }

  MSIM_DEV void init() {
    state = SCR_Healthy;
    dx = screen_dx = false;
    age_dx = 0.0;
    age_o = 35.0 + sqrt(-2.0 * log(R.runif(0.0, 1.0)) / K.g0);
    grade = (R.runif(0.0, 1.0) < 0.006 * (age_o - 35.0)) ? 1 : 0;
    beta0 = R.rnorm(K.mubeta0, K.sebeta0);
    beta1 = rnormPos(K.mubeta1, K.sebeta1);
    beta2 = rnormPos(K.mubeta2[grade], K.sebeta2[grade]);
    eps = R.rnorm(0.0, K.tau2);
    int failed = 0;
    scheduleAt(msim::rpexp_rand(d_tables.mortality, R.runif(0.0, 1.0), 0.0, &failed), SCR_toOtherDeath);
    scheduleAt(age_o, SCR_toLocalised);
    if (K.screen > 0.0) scheduleAt(50.0, SCR_toScreen);
  }

  MSIM_DEV void handleMessage(int kind) {
    env.report(state, kind, previous(), now());
    switch (kind) {
      case SCR_toOtherDeath:
      case SCR_toCancerDeath:
        env.value(0, now());
        env.value(1, age_dx);
        env.value(2, screen_dx ? 1.0 : 0.0);
        env.value(3, (double)grade);
        stop_simulation();
        break;
      case SCR_toLocalised:
        state = SCR_Localised;
        scheduleAt(now() + msim::gsm_randU(*d_tables.sojourn, R.runif(0.0, 1.0), 0.0, grade, 10.0), SCR_toClinicalDiagnosis);
        break;
      case SCR_toScreen: {
        const double psa = exp(fmin(log(20.0), beta0 + beta1 * (now() - 35.0) + beta2 * fmax(0.0, now() - age_o) + eps));
        if (psa >= K.psa_threshold) {
          const double key[2] = {psa, now()};
          if (R.runif(0.0, 1.0) < K.compliance * msim::table_lookup(d_tables.biopsy, key)) scheduleAt(now(), SCR_toBiopsy);
        }
        if (now() + K.screen <= 70.0) scheduleAt(now() + K.screen, SCR_toScreen);
        break;
      }
      case SCR_toBiopsy:
        if (state == SCR_Localised && R.runif(0.0, 1.0) < msim::interp_step(d_tables.sensitivity, K.cohort + now()))
          scheduleAt(now(), SCR_toScreenDiagnosis);
        break;
      case SCR_toClinicalDiagnosis:
      case SCR_toScreenDiagnosis:
        if (!dx) {
          dx = true;
          screen_dx = kind == SCR_toScreenDiagnosis;
          age_dx = now();
          state = SCR_Diagnosed;
          RemoveKind(SCR_toScreen);
          RemoveKind(SCR_toBiopsy);
          RemoveKind(SCR_toClinicalDiagnosis);
          RemoveKind(SCR_toScreenDiagnosis);
          const double u = R.runif(0.0, 1.0);
          if (u >= d_tables.surv_last) scheduleAt(now() + msim::interp_invert_decreasing(d_tables.survival, u), SCR_toCancerDeath);
        }
        break;
      default:
        break;
    }
  }
};

struct Traits {
  typedef scr::Consts Consts;
  template <class Env> using Model = Person<Env>;
  static constexpr int N_VALUES = SCR_N_VALUES;
  static const char *const *value_names() {
    static const char *const names[SCR_N_VALUES] = {"age_death", "age_dx", "screen_dx", "grade"};
    return names;
  }
};

// device copies of the tables: one allocation, kept until the next setup
static double *d_blob = nullptr;

}  // namespace scr

MSIM_PLUGIN(screening_skeleton, scr::Traits)

// Uploads the look-up tables of `in` (everything but the scalars, which travel with each call in scr::Consts) and the
// sojourn-time model.  `device` as in msim_plugin_call.  Returns 0 or an MSIM_ERR_* code (see screening_skeleton_last_error).
MSIM_PLUGIN_EXPORT int screening_skeleton_setup(const screening_inputs *in, int device) {
  using namespace msim_host;
  int rc;
  if (!in || !in->gsm || in->n_mort < 1 || in->n_biopsy < 1 || in->n_sens < 2 || in->n_surv < 2) return fail(MSIM_ERR_ARG, "bad tables");
  if (g.device != device && g.inited) return fail(MSIM_ERR_STATE, "this plugin is already bound to another device");
  g.device = device;
  if ((rc = ensure_init())) return rc;
  // the sojourn-time model goes through the library's own upload (terms expanded into per-interval cubics)
  if ((rc = gsm_upload(in->gsm))) return rc;
  std::vector<double> blob;
  auto put = [&](const double *v, size_t n) {
    const size_t at = blob.size();
    blob.insert(blob.end(), v, v + n);
    return at;
  };
  // Rpexp: cumulative hazards at the break points, hazards, times
  std::vector<double> H(in->n_mort);
  rpexp_cumulate(in->n_mort, in->mort_rate, in->mort_age, H.data());
  const size_t o_H = put(H.data(), H.size()), o_h = put(in->mort_rate, in->n_mort), o_t = put(in->mort_age, in->n_mort);
  // Table<double, double, double>: sorted distinct keys per axis, outcomes on the dense grid (0 where no row names the cell)
  std::vector<double> ax[2];
  for (int a = 0; a < 2; a++) {
    for (int r = 0; r < in->n_biopsy; r++) ax[a].push_back(in->biopsy_rows[3 * r + a]);
    std::sort(ax[a].begin(), ax[a].end());
    ax[a].erase(std::unique(ax[a].begin(), ax[a].end()), ax[a].end());
  }
  std::vector<double> grid(ax[0].size() * ax[1].size(), 0.0);
  for (int r = 0; r < in->n_biopsy; r++) {
    const size_t i0 = std::lower_bound(ax[0].begin(), ax[0].end(), in->biopsy_rows[3 * r]) - ax[0].begin();
    const size_t i1 = std::lower_bound(ax[1].begin(), ax[1].end(), in->biopsy_rows[3 * r + 1]) - ax[1].begin();
    grid[i0 * ax[1].size() + i1] = in->biopsy_rows[3 * r + 2];
  }
  const size_t o_a0 = put(ax[0].data(), ax[0].size()), o_a1 = put(ax[1].data(), ax[1].size()), o_grid = put(grid.data(), grid.size());
  // NumericInterpolate x 2: points and slopes
  auto slopes = [](const double *x, const double *y, int n) {
    std::vector<double> s(n - 1);
    for (int i = 0; i < n - 1; i++) s[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    return s;
  };
  const std::vector<double> s_sens = slopes(in->sens_year, in->sens_value, in->n_sens), s_surv = slopes(in->surv_time, in->surv_value, in->n_surv);
  const size_t o_sx = put(in->sens_year, in->n_sens), o_sy = put(in->sens_value, in->n_sens), o_ss = put(s_sens.data(), s_sens.size());
  const size_t o_vx = put(in->surv_time, in->n_surv), o_vy = put(in->surv_value, in->n_surv), o_vs = put(s_surv.data(), s_surv.size());
  if (scr::d_blob) CK(cudaFree(scr::d_blob));
  scr::d_blob = nullptr;
  CK(cudaMalloc(&scr::d_blob, sizeof(double) * blob.size()));
  CK(cudaMemcpy(scr::d_blob, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice));
  const double *d = scr::d_blob;
  scr::Tables T;
  memset(&T, 0, sizeof T);
  T.mortality = msim::Rpexp{in->n_mort, d + o_H, d + o_h, d + o_t};
  T.biopsy.n_axes = 2;
  T.biopsy.n[0] = (int)ax[0].size();
  T.biopsy.n[1] = (int)ax[1].size();
  T.biopsy.axis[0] = d + o_a0;
  T.biopsy.axis[1] = d + o_a1;
  T.biopsy.values = d + o_grid;
  T.sensitivity = msim::Interp{in->n_sens, d + o_sx, d + o_sy, d + o_ss};
  T.survival = msim::Interp{in->n_surv, d + o_vx, d + o_vy, d + o_vs};
  T.sojourn = (const msim::GsmModel *)g.gsm_model.p;
  T.surv_last = in->surv_value[in->n_surv - 1];
  CK(cudaMemcpyToSymbol(scr::d_tables, &T, sizeof T));
  return MSIM_OK;
}
