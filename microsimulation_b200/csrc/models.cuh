// The reference's models, written against the device plugin surface
// (engine.cuh) so that init()/handleMessage() bodies keep the reference's
// statement order — the order of random draws is part of the contract.
//
//   Person        FHCRC prostate natural history   src/person-r.cc:36-190
//   CalibPerson   calibration example              src/calibperson-r.cc:35-171
//   IllnessDeath  illnessDeath::SimplePerson       src/illness-death.cpp:6-71
//   SimplePerson  row-log / EventReport variants   src/simple-example.cc:11-70,
//                                                  src/simple-example2.cc:12-60
//
// Each model is a template over `Env`, which supplies the uniform source
// (`Env::Src`) and the output sinks the reference writes to from inside
// handleMessage: row(...) for the life-history row log, report(...) for
// EventReport::add, count(...) for per-kind tallies.  Constants that the
// reference computes with libm from literals (exp(2.3525), pow(hr, 1/shape),
// b_weibull via gammafn) are evaluated ONCE on the host with the host's libm
// and passed in, so they are bit-identical to the CPU run.  Run-time exp / log / pow are fastmath.cuh's (fm_*):
// within 1.05 ulp of the true value, event times within 2 ulp of the reference's glibc values.
#pragma once
#include "engine.cuh"

namespace msim {

// ------------------------------------------------------------------ person-r
namespace person_r {

enum gleason_t { nogleason, gleasonLt7, gleason7, gleasonGt7 };
enum stage_t { Healthy, Localised, DxLocalised, LocallyAdvanced, DxLocallyAdvanced, Metastatic, DxMetastatic, Death };
enum event_t {
  toDeath, toPCDeath, toLocalised, toDxLocalised, toDxLocallyAdvanced, toLocallyAdvanced, toMetastatic, toDxMetastatic
};
constexpr int N_STATE = 8, N_EVENT = 8, MAX_ROWS = 5, HEAP_CAP = 4;

// (stage, event) pairs handleMessage can see (person-r.cc:108-190): `stage` only ever holds the four undiagnosed
// stages (diagnosis sets a flag), and from each of them a person dies of other causes, is diagnosed, or progresses.
constexpr int REACHABLE_PAIRS[][2] = {{Healthy, toDeath},           {Healthy, toLocalised},
                                      {Localised, toDeath},         {Localised, toDxLocalised},
                                      {Localised, toLocallyAdvanced}, {LocallyAdvanced, toDeath},
                                      {LocallyAdvanced, toDxLocallyAdvanced}, {LocallyAdvanced, toMetastatic},
                                      {Metastatic, toDeath},        {Metastatic, toDxMetastatic},
                                      {Metastatic, toPCDeath}};
constexpr int N_REACHABLE_PAIRS = sizeof(REACHABLE_PAIRS) / sizeof(REACHABLE_PAIRS[0]);

struct Consts {
  double onset_inv_shape;  // 1/exp(2.3525)
  double onset_scale;      // 64.0218
  // dwell-time Weibulls of toLocalised / toLocallyAdvanced / toMetastatic:
  // inv_shape[k] = 1/shape_k; scale[k][gleason] = scale_k*pow(progressionHR(gleason)*dxHR(stage_k), 1/shape_k)
  double inv_shape[3];
  double scale[3][4];
  double pDx[3];
};

// host side: person-r.cc:80-93 (dxHR, progressionHR) and microsimulation.cc:6-8
inline Consts make_consts() {
  Consts c;
  c.onset_inv_shape = 1. / exp(2.3525);
  c.onset_scale = 64.0218;
  const double shape[3] = {exp(1.0353), exp(1.4404), exp(1.4404)};
  const double scale[3] = {19.8617, 16.3863, 1.4242};
  const double dxHR[3] = {1.1308, 0.5900, 1.3147};  // Localised, LocallyAdvanced, Metastatic
  const double prog[4] = {1.4027 * 1.3874 /* nogleason falls to the last branch */, 1, 1.3874, 1.4027 * 1.3874};
  for (int k = 0; k < 3; k++) {
    c.inv_shape[k] = 1. / shape[k];
    for (int g = 0; g < 4; g++) {
      // the reference evaluates this pow at run time (microsimulation.cc:7); keep
      // the compiler from folding it with different rounding
      volatile double hr = prog[g] * dxHR[k], e = 1.0 / shape[k];
      c.scale[k][g] = scale[k] * pow(hr, e);
    }
  }
  c.pDx[0] = 1.1308 / (2.1308);
  c.pDx[1] = 0.5900 / (1.0 + 0.5900);
  c.pDx[2] = 1.3147 / (1.0 + 1.3147);
  return c;
}

template <class Env>
struct Person : cProcess<Person<Env>, typename Env::Src, HEAP_CAP> {
  typedef cProcess<Person<Env>, typename Env::Src, HEAP_CAP> Base;
  using Base::now;
  using Base::previousEventTime;
  using Base::R;
  using Base::scheduleAt;
  using Base::stop_simulation;

  int gleason, stage;
  bool dx;
  int64_t id;
  Env &env;
  const Consts &K;

  MSIM_DEV Person(Env &e, const Consts &k, int64_t i)
      : Base(&e.src), gleason(nogleason), stage(Healthy), dx(false), id(i), env(e), K(k) {
    e.begin_person();
  }

  MSIM_DEV void init() {  // person-r.cc:98-103
    if (R.runif(0.0, 1.0) < 0.2241) scheduleAt(K.onset_scale * fm_pow(-fm_log(R.unif_rand()), K.onset_inv_shape), toLocalised);
    scheduleAt(R.rexp(80.0), toDeath);
  }

  // init() in two steps for the staged kernel (person_staged.cuh): init_draws()
  // takes init()'s uniforms in init()'s order but leaves the onset Weibull's
  // log/pow undone; onset_time() finishes it later, in a warp made up entirely
  // of persons who need it.
  MSIM_DEV static bool init_draws(RMath<typename Env::Src> &R, double *u_onset, double *t_death) {
    bool onset = R.runif(0.0, 1.0) < 0.2241;
    if (onset) *u_onset = R.unif_rand();
    *t_death = R.rexp(80.0);
    return onset;
  }
  MSIM_DEV static double onset_time(const Consts &K, double u_onset) {
    return K.onset_scale * fm_pow(-fm_log(u_onset), K.onset_inv_shape);
  }

  // dwell = now() + rweibullHR(shape_k, scale_k, progressionHR(gleason)*dxHR(stage))
  MSIM_DEV double dwell(int k) { return now() + K.scale[k][gleason] * fm_pow(-fm_log(R.unif_rand()), K.inv_shape[k]); }

  MSIM_DEV void handleMessage(int kind) {  // person-r.cc:108-190
    env.row((double)id, previousEventTime, now(), stage, kind);  // logged BEFORE the state change (:112-116)
    env.report(stage, kind, previousEventTime, now());  // EventReport attachment (BASELINE config 4); no-op unless requested

    if (kind == toDeath || kind == toPCDeath) {  // :120-126
      stop_simulation();
    } else if (kind == toLocalised || kind == toLocallyAdvanced || kind == toMetastatic) {
      // :128-173.  The three progression handlers have the same shape — new
      // stage, [Gleason draws on toLocalised], one Weibull dwell time, one
      // uniform choosing between diagnosis and further progression — so lanes
      // of a warp that handle different ones of them run the same instructions
      // (the expensive log/pow once) with per-lane constants; the order of the
      // draws is the reference's.
      const int k = (kind == toLocalised) ? 0 : (kind == toLocallyAdvanced) ? 1 : 2;
      stage = (k == 0) ? Localised : (k == 1) ? LocallyAdvanced : Metastatic;
      if (k == 0)
        gleason = (R.runif(0.0, 1.0) < 0.6812) ? gleasonLt7 : ((R.runif(0.0, 1.0) < 0.5016) ? gleason7 : gleasonGt7);
      const double dwellTime = dwell(k);
      const int dxKind = (k == 0) ? toDxLocalised : (k == 1) ? toDxLocallyAdvanced : toDxMetastatic;
      const int nextKind = (k == 0) ? toLocallyAdvanced : (k == 1) ? toMetastatic : toPCDeath;
      scheduleAt(dwellTime, (R.runif(0.0, 1.0) < K.pDx[k]) ? dxKind : nextKind);
    } else if (kind == toDxLocalised || kind == toDxLocallyAdvanced || kind == toDxMetastatic) {
      dx = true;  // :175-188
    }
  }
};

}  // namespace person_r

// --------------------------------------------------------------- calibperson
namespace calib {

enum stage_t { DiseaseFree, Precursor, PreClinical, Clinical, Death };
enum event_t { toPrecursor, toPreClinical, toClinical, toDeath, Count };
constexpr int HEAP_SLOTS = 14;  // at most 12 pending: ten Count events + toPrecursor (or its successor) + toDeath
constexpr int HEAP_CAP = -HEAP_SLOTS;  // negative: the slots live in memory the environment provides (engine.cuh)

struct Par {
  double Lam1, sigm1, p2, lam2, mu3, tau3;
};

// The TimeAtRisk bookkeeping of the toDeath handler (calibperson-r.cc:124-135): f(i, min(ctime[i], clinTime)) for
// i = 0.. until clinTime < ctime[i].  A person dies once, so an environment may run this after the event loop, where
// the warp is together again (in the handler only the few lanes that die in the same pass are active).
template <class F>
MSIM_HD void time_at_risk_terms(double clinTime, F f) {
  const double ctime[4] = {20, 40, 60, 80};
  for (int i = 0; i < 4; i++) {
    f(i, fmin(ctime[i], clinTime));
    if (clinTime < ctime[i]) break;
  }
}

// ORDERED: pending actions in an ordered array (engine.cuh, OrderedEvents) instead of the binary heap
template <class Env, bool ORDERED = false>
struct CalibPerson : cProcess<CalibPerson<Env, ORDERED>, typename Env::Src, HEAP_CAP,
                              typename std::conditional<ORDERED, OrderedEvents<HEAP_SLOTS>, EventHeap<HEAP_CAP> >::type> {
  typedef cProcess<CalibPerson<Env, ORDERED>, typename Env::Src, HEAP_CAP,
                   typename std::conditional<ORDERED, OrderedEvents<HEAP_SLOTS>, EventHeap<HEAP_CAP> >::type> Base;
  using Base::now;
  using Base::R;
  using Base::scheduleAt;
  using Base::stop_simulation;

  int stage;
  bool diseasepot;
  double clinTime;
  Env &env;
  const Par &P;

  MSIM_DEV CalibPerson(Env &e, const Par &p) : Base(&e.src), stage(DiseaseFree), diseasepot(false), clinTime(1000), env(e), P(p) {
    this->actions.bind(e.heap_t, e.heap_m, e.heap_stride);
  }

  MSIM_DEV void init() {  // calibperson-r.cc:95-110
    if (R.runif(0, 1) < P.p2) diseasepot = true;
    else diseasepot = false;
    clinTime = 1000;
    stage = DiseaseFree;
    double lam1 = fm_exp(R.rnorm(P.Lam1, P.sigm1));
    scheduleAt(R.rexp(lam1), toPrecursor);
    double x = R.runif(0, 1);
    scheduleAt((65 - 15 * fm_log(-fm_log(x))), toDeath);  // Gumbel
    for (int i = 10; i < 110; i = i + 10) scheduleAt(i, Count);
  }

  MSIM_DEV void handleMessage(int kind) {  // calibperson-r.cc:115-171
    if (kind == toDeath) {
      stage = Death;
      clinTime = fmin(clinTime, now());
      env.at_death(clinTime);  // TimeAtRisk: time_at_risk_terms(), run by the environment
      stop_simulation();
    } else if (kind == toPrecursor) {
      stage = Precursor;
      if (diseasepot) {
        double dwellTime = now() + R.rexp(P.lam2);
        scheduleAt(dwellTime, toPreClinical);
      }
    } else if (kind == toPreClinical) {
      stage = PreClinical;
      double dwellTime = now() + fm_exp(R.rnorm(P.mu3, P.tau3 * P.mu3));
      scheduleAt(dwellTime, toClinical);
    } else if (kind == toClinical) {
      stage = Clinical;
      clinTime = now();
    } else if (kind == Count) {
      int cind = (int)(now() / 10 - 1);
      if (cind > 9) cind = 9;
      env.occupancy(stage, cind);
    }
  }
};

}  // namespace calib

// ------------------------------------------------------------- illness-death
namespace illness_death {

enum state_t { Healthy, Cancer };
enum event_t { toOtherDeath, toCancer, toCancerDeath };
constexpr int N_STATE = 2, N_EVENT = 3, HEAP_CAP = 4, MAX_ROWS = 3, MAX_DRAWS = 7;

struct Consts {
  double cure, zsd;
  double b_other;         // b_weibull(80, 4)          illness-death.cpp:17-19, :37
  double b_cancer_mean;   // 80 / gammafn(1 + 1/3)     the rr-independent factor of b_weibull(80, 3, z)
};

template <class Env>
struct SimplePerson : cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> {
  typedef cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> Base;
  using Base::now;
  using Base::previousEventTime;
  using Base::R;
  using Base::RemoveKind;
  using Base::scheduleAt;
  using Base::stop_process;

  int state;
  double z;
  Env &env;
  const Consts &K;

  MSIM_DEV SimplePerson(Env &e, const Consts &k, int64_t /*id*/) : Base(&e.src), state(Healthy), z(1.0), env(e), K(k) {
    e.begin_person();
  }

  MSIM_DEV void init() {  // illness-death.cpp:35-41
    state = Healthy;
    z = exp(R.rnorm(0.0, K.zsd));
    scheduleAt(R.rweibull(4.0, K.b_other), toOtherDeath);
    if (R.runif(0.0, 1.0) > K.cure) scheduleAt(R.rweibull(3.0, K.b_cancer_mean * pow(z, -1.0 / 3.0)), toCancer);
  }

  MSIM_DEV void handleMessage(int kind) {  // illness-death.cpp:46-71
    env.report(state, kind, previousEventTime, now());
    switch (kind) {
      case toOtherDeath:
      case toCancerDeath:
        stop_process();
        break;
      case toCancer:
        state = Cancer;
        RemoveKind(toOtherDeath);
        if (R.runif(0.0, 1.0) < 0.5) scheduleAt(now() + R.rweibull(1.0, 10.0), toCancerDeath);
        break;
      default:
        break;
    }
  }
};

}  // namespace illness_death

// ------------------------------------------- simple-example / simple-example2
namespace simple_example {

enum state_t { Healthy, Cancer, Death };
enum event_t { toOtherDeath, toCancer, toCancerDeath };
constexpr int N_STATE = 3, N_EVENT = 3, HEAP_CAP = 4, MAX_ROWS = 2, MAX_DRAWS = 4;

struct Consts {
  int unused;
};

template <class Env>
struct SimplePerson : cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> {
  typedef cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> Base;
  using Base::now;
  using Base::previousEventTime;
  using Base::R;
  using Base::scheduleAt;
  using Base::stop_process;

  int state;
  int64_t id;
  Env &env;

  MSIM_DEV SimplePerson(Env &e, const Consts &, int64_t i) : Base(&e.src), state(Healthy), id(i), env(e) {
    e.begin_person();
  }

  MSIM_DEV void init() {  // simple-example.cc:26-32, simple-example2.cc:28-32
    state = Healthy;
    double tm = R.rweibull(8.0, 85.0);
    scheduleAt(tm, toOtherDeath);
    scheduleAt(R.rweibull(3.0, 90.0), toCancer);
  }

  MSIM_DEV void handleMessage(int kind) {  // simple-example.cc:41-70, simple-example2.cc:37-60
    env.row((double)id, previousEventTime, now(), state, kind);  // callSimplePerson
    env.report(state, kind, previousEventTime, now());           // callSimplePerson2
    switch (kind) {
      case toOtherDeath:
      case toCancerDeath:
        stop_process();
        break;
      case toCancer:
        state = Cancer;
        if (R.runif(0.0, 1.0) < 0.5) scheduleAt(now() + R.rweibull(2.0, 10.0), toCancerDeath);
        break;
      default:
        break;
    }
  }
};

}  // namespace simple_example

// ------------------------------------------------- Sick-Sicker (README.org:148-301)
// The reference's documented SummaryReport example: a continuous-time
// Healthy / Sick / Sicker / Dead model with utilities and costs that change at
// each transition, 31 years of follow-up, exponential transition times.
namespace sick_sicker {

enum state_t { Healthy, Sick, Sicker, Dead };
enum event_t { toS1, toS2, toH, toD, toEOF };
// every transition cancels the pending events lazily (they stay in the heap until their time comes) and
// schedules up to three new ones, so the heap holds a few entries per transition made so far
constexpr int N_STATE = 4, N_EVENT = 5, HEAP_CAP = 160;

struct Par {
  double r_HS1, r_HD, r_S1H, r_S1S2, r_S1D, r_S2D;
  double c_H, c_S1, c_S2, c_Trt, u_H, u_S1, u_S2, u_Trt;
  int Trt;
};

// ORDERED: pending actions in a small ordered array with cancelled actions removed (engine.cuh, OrderedEvents): at
// most toEOF and three competing transitions are live at any time
constexpr int LIVE_CAP = 8;
template <class Env, bool ORDERED = false>
struct SickSicker : cProcess<SickSicker<Env, ORDERED>, typename Env::Src, HEAP_CAP,
                             typename std::conditional<ORDERED, OrderedEvents<LIVE_CAP, HeapStore<LIVE_CAP> >, EventHeap<HEAP_CAP> >::type> {
  typedef cProcess<SickSicker<Env, ORDERED>, typename Env::Src, HEAP_CAP,
                   typename std::conditional<ORDERED, OrderedEvents<LIVE_CAP, HeapStore<LIVE_CAP> >, EventHeap<HEAP_CAP> >::type> Base;
  using Base::Cancel;
  using Base::now;
  using Base::previousEventTime;
  using Base::R;
  using Base::scheduleAt;
  using Base::stop_simulation;

  int state;
  Env &env;
  const Par &P;

  MSIM_DEV SickSicker(Env &e, const Par &p, int64_t /*id*/) : Base(&e.src), state(Healthy), env(e), P(p) { e.begin_person(); }

  MSIM_DEV double rexpRate(double rate) { return R.rexp(1.0 / rate); }  // README.org:200

  MSIM_DEV void init() {  // README.org:226-232
    state = Healthy;
    scheduleAt(rexpRate(P.r_HS1), toS1);
    scheduleAt(rexpRate(P.r_HD), toD);
    scheduleAt(31.0, toEOF);  // end of follow-up
  }

  // RemoveKind(toH); RemoveKind(toS1); RemoveKind(toS2); RemoveKind(toD)  (README.org:236-241) in one scan
  MSIM_DEV void cancelEvents() {
    Cancel([](int k) { return k == toH || k == toS1 || k == toS2 || k == toD; });
  }

  MSIM_DEV void handleMessage(int kind) {  // README.org:245-285
    env.summary(state, kind, previousEventTime, now());
    switch (kind) {
      case toH:
        state = Healthy;
        env.set_uc(1, P.u_H, P.c_H);
        cancelEvents();
        scheduleAt(now() + rexpRate(P.r_HS1), toS1);
        scheduleAt(now() + rexpRate(P.r_HD), toD);
        break;
      case toS1:
        state = Sick;
        env.set_uc(2, P.Trt ? P.u_Trt : P.u_S1, P.c_S1 + (P.Trt ? P.c_Trt : 0.0));
        cancelEvents();
        scheduleAt(now() + rexpRate(P.r_S1H), toH);
        scheduleAt(now() + rexpRate(P.r_S1S2), toS2);
        scheduleAt(now() + rexpRate(P.r_S1D), toD);
        break;
      case toS2:
        state = Sicker;
        env.set_uc(3, P.u_S2, P.c_S2 + (P.Trt ? P.c_Trt : 0.0));
        cancelEvents();
        scheduleAt(now() + rexpRate(P.r_S2D), toD);
        break;
      case toD:
      case toEOF:
        stop_simulation();
        break;
      default:
        break;
    }
  }
};

// The report's utility and cost are NOT reset between persons in the
// reference (README.org:292-299: one Report, one process object, n runs), so a
// person's first interval is valued with whatever the previous person's last
// transition set.  code 0 = nothing set yet (the constructor's utility 1, cost
// 0), 1/2/3 = the values toH / toS1 / toS2 set.
MSIM_HD void uc_of_code(const Par &P, int code, double *u, double *c) {
  if (code == 1) {
    *u = P.u_H;
    *c = P.c_H;
  } else if (code == 2) {
    *u = P.Trt ? P.u_Trt : P.u_S1;
    *c = P.c_S1 + (P.Trt ? P.c_Trt : 0.0);
  } else if (code == 3) {
    *u = P.u_S2;
    *c = P.c_S2 + (P.Trt ? P.c_Trt : 0.0);
  } else {
    *u = 1.0;
    *c = 0.0;
  }
}

}  // namespace sick_sicker


// ------------- SummaryReport + addPointCost (test/test_microsimulation.R:782-865, 1006-1093)
// The reference's dev-script example with costs: other-cause death and cancer onset as Weibulls, a one-off cost when
// cancer starts and a management cost every two years after (SummaryReport::addPointCost, microsimulation.h:1105-1112).
namespace cancer_mgmt {

enum state_t { Healthy, Cancer, Death };
enum event_t { toOtherDeath, toCancer, toCancerDeath, toCancerManagement };
constexpr int N_STATE = 3, N_EVENT = 4, HEAP_CAP = 4;  // pending: toOtherDeath, toCancerManagement, toCancerDeath

struct Par {
  double odShape;
};

template <class Env>
struct SimplePerson : cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> {
  typedef cProcess<SimplePerson<Env>, typename Env::Src, HEAP_CAP> Base;
  using Base::now;
  using Base::previous;
  using Base::R;
  using Base::scheduleAt;
  using Base::stop_simulation;

  int state;
  double utility;
  Env &env;
  const Par &P;

  MSIM_DEV SimplePerson(Env &e, const Par &p, int64_t /*id*/) : Base(&e.src), state(Healthy), utility(1.0), env(e), P(p) { e.begin_person(); }

  MSIM_DEV void init() {  // test_microsimulation.R:1038-1044
    state = Healthy;
    utility = 1.0;
    scheduleAt(R.rweibull(P.odShape, 85.0), toOtherDeath);
    scheduleAt(R.rweibull(3.0, 90.0), toCancer);
  }

  MSIM_DEV void handleMessage(int kind) {  // test_microsimulation.R:1048-1071
    env.summary(state, kind, previous(), now());
    switch (kind) {
      case toOtherDeath:
      case toCancerDeath:
        stop_simulation();
        break;
      case toCancer:
        state = Cancer;
        utility = 0.8;
        env.point_cost(state, now(), 15000.0);
        scheduleAt(now() + 2.0, toCancerManagement);
        if (R.runif(0.0, 1.0) < 0.5) scheduleAt(now() + R.rweibull(2.0, 10.0), toCancerDeath);
        break;
      case toCancerManagement:
        env.point_cost(state, now(), 1000.0);
        scheduleAt(now() + 2.0, toCancerManagement);
        break;
      default:
        break;
    }
  }
};

}  // namespace cancer_mgmt

}  // namespace msim
