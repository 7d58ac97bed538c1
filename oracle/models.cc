// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// The reference's models, transcribed against the restated plugin layer
// (cprocess_lite.h / reports_lite.h) because the originals include Rcpp:
//   Person (FHCRC natural history)   src/person-r.cc:36-226
//   CalibPerson                      src/calibperson-r.cc:35-200
//   illnessDeath::SimplePerson       src/illness-death.cpp:6-94
//   SimplePerson (row log)           src/simple-example.cc:3-83
//   SimplePerson (EventReport)       src/simple-example2.cc:4-83
// plus the documentation models whose printed output pins the oracle:
//   SickSicker (SummaryReport)       README.org:214-287
//   tick_tock                        inst/doc/examples.org:121-160
//   illness_death with QALYs/costs   inst/doc/examples.org:294-345
// Handler bodies keep the reference's statement order: the order of draws is
// part of the contract.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "models.h"
#include "reports_lite.h"

using namespace ssim;

namespace orc {

uint64_t g_events_handled = 0;  // instrumentation for the work model

// ---------------------------------------------------------------- person-r
namespace person_r {

enum gleason_t { nogleason, gleasonLt7, gleason7, gleasonGt7 };
enum stage_t { Healthy, Localised, DxLocalised, LocallyAdvanced, DxLocallyAdvanced, Metastatic, DxMetastatic, Death };
enum event_t {
  toDeath, toPCDeath, toLocalised, toDxLocalised, toDxLocallyAdvanced, toLocallyAdvanced, toMetastatic, toDxMetastatic
};

static RowLog *g_log = 0;
static bool g_keep_rows = true;
static double g_counts[9];
static std::map<std::string, Rng *> rng;
typedef EventReport<short, short, double> Report;
static Report *g_report = 0;

double dxHR(stage_t stage) {
  return stage == Healthy ? -1 : (stage == Localised ? 1.1308 : (stage == LocallyAdvanced ? 0.5900 : 1.3147));
}
double progressionHR(gleason_t gleason) {
  return gleason == gleasonLt7 ? 1 : (gleason == gleason7 ? 1.3874 : 1.4027 * 1.3874);
}

class Person : public cProcess {
 public:
  gleason_t gleason;
  stage_t stage;
  bool dx;
  int id;
  Person(int i = 0) : gleason(nogleason), stage(Healthy), dx(false), id(i) {}
  void init();
  virtual void handleMessage(const cMessage *msg);
};

void Person::init() {
  rng["NH"]->set();
  if (R::runif(0.0, 1.0) < 0.2241) scheduleAt(R::rweibull(exp(2.3525), 64.0218), toLocalised);
  scheduleAt(R::rexp(80.0), toDeath);
}

void Person::handleMessage(const cMessage *msg) {
  double pDx;
  g_events_handled++;
  if (g_keep_rows) g_log->push(id, previousEventTime, now(), stage, msg->kind);
  g_counts[msg->kind] += 1;
  g_counts[8] += now();
  if (g_report) g_report->add(stage, msg->kind, previousEventTime, now());

  rng["NH"]->set();

  if (msg->kind == toDeath) {
    Sim::stop_simulation();
  } else if (msg->kind == toPCDeath) {
    Sim::stop_simulation();
  } else if (msg->kind == toLocalised) {
    stage = Localised;
    gleason = (R::runif(0.0, 1.0) < 0.6812) ? gleasonLt7 : ((R::runif(0.0, 1.0) < 0.5016) ? gleason7 : gleasonGt7);
    Time dwellTime = now() + rweibullHR(exp(1.0353), 19.8617, progressionHR(gleason) * dxHR(stage));
    pDx = 1.1308 / (2.1308);
    if (R::runif(0.0, 1.0) < pDx) scheduleAt(dwellTime, toDxLocalised);
    else scheduleAt(dwellTime, toLocallyAdvanced);
  } else if (msg->kind == toLocallyAdvanced) {
    stage = LocallyAdvanced;
    Time dwellTime = now() + rweibullHR(exp(1.4404), 16.3863, progressionHR(gleason) * dxHR(stage));
    pDx = 0.5900 / (1.0 + 0.5900);
    if (R::runif(0.0, 1.0) < pDx) scheduleAt(dwellTime, toDxLocallyAdvanced);
    else scheduleAt(dwellTime, toMetastatic);
  } else if (msg->kind == toMetastatic) {
    stage = Metastatic;
    Time dwellTime = now() + rweibullHR(exp(1.4404), 1.4242, progressionHR(gleason) * dxHR(stage));
    pDx = 1.3147 / (1.0 + 1.3147);
    if (R::runif(0.0, 1.0) < pDx) scheduleAt(dwellTime, toDxMetastatic);
    else scheduleAt(dwellTime, toPCDeath);
  } else if (msg->kind == toDxLocalised) {
    dx = true;
  } else if (msg->kind == toDxLocallyAdvanced) {
    dx = true;
  } else if (msg->kind == toDxMetastatic) {
    dx = true;
  }
}

}  // namespace person_r

// Driver loop of callPersonSimulation (person-r.cc:195-226) with two
// extensions used only by tests/bench: start at `first_person` (the skipped
// persons' substream jumps are done one by one, exactly as the loop would)
// and optional EventReport attachment (BASELINE config 4).
void run_person_r(const double seed_in[6], int64_t first_person, int64_t n, bool keep_rows, RowLog *log,
                  double counts[9], ReportTables *report, double discount_rate, const PersonReportOptions *opt,
                  double *indiv, double means[7]) {
  using namespace person_r;
  double seed[6];
  for (int i = 0; i < 6; i++) seed[i] = seed_in[i];
  rlite::RNGkind(rlite::USER_UNIF);
  rlite::set_user_unif_rand(user_unif_rand);
  RngStream::SetPackageSeed(seed);
  g_log = log;
  g_keep_rows = keep_rows;
  memset(g_counts, 0, sizeof g_counts);
  Report rep(discount_rate, opt ? opt->output_utilities != 0 : true, opt ? opt->start_report_age : 0.0);
  if (report) {
    rep.clear();
    if (opt) {  // EventReport::setPartition(start, finish, delta, maxTime)
      rep.setPartition(opt->start, opt->finish, opt->delta, opt->max_time);
    } else {
      std::vector<double> ages;
      for (int a = 0; a <= 100; a++) ages.push_back(a);
      ages.push_back(1.0e+6);
      rep.setPartition(ages);
    }
    if (indiv) rep.setIndivN((int)n);
    g_report = &rep;
  } else {
    g_report = 0;
  }
  rng["NH"] = new Rng();
  rng["S"] = new Rng();
  rng["NH"]->set();
  if (first_person > 0) {
    rng["NH"]->AdvanceSubstream(0, (int32_t)first_person);
    rng["S"]->AdvanceSubstream(0, (int32_t)first_person);
  }
  Person person;
  for (int64_t i = first_person; i < first_person + n; i++) {
    rng["NH"]->nextSubstream();
    rng["S"]->nextSubstream();
    person = Person((int)i);
    Sim::create_process(&person);
    Sim::run_simulation();
    if (g_report) g_report->individualReset();  // at the end of the individual, while now() is its last event's time
    Sim::clear();
  }
  delete rng["NH"];
  delete rng["S"];
  if (counts) memcpy(counts, g_counts, sizeof g_counts);
  if (report) flatten(rep, report);
  if (report && indiv)
    for (int64_t i = 0; i < n; i++) indiv[i] = rep._vector[(size_t)i];
  if (report && means) {
    Means &m = rep.mean_utilities;
    means[0] = m.n(); means[1] = m.mean(); means[2] = m.var(); means[3] = m.sd(); means[4] = m.sd() / sqrt((double)m.n());
    means[5] = m.sum(); means[6] = m.sumsq();
  }
  g_report = 0;
}

// ------------------------------------------------------------ calibperson
namespace calib {

enum stage_t { DiseaseFree, Precursor, PreClinical, Clinical, Death };
enum event_t { toPrecursor, toPreClinical, toClinical, toDeath, Count };
static std::string stage_names[5] = {"DiseaseFree", "Precursor", "PreClinical", "Clinical", "Death"};
static Rng *rng;
static std::map<std::string, std::vector<double> > report;

class CalibPerson : public cProcess {
 public:
  stage_t stage;
  bool diseasepot;
  double Lam1, sigm1, p2, lam2, mu3, tau3;
  double clinTime;
  int id;
  CalibPerson() {}
  CalibPerson(const double *par, int i = 0) {
    Lam1 = par[0];
    sigm1 = par[1];
    p2 = par[2];
    lam2 = par[3];
    mu3 = par[4];
    tau3 = par[5];
    id = i;
    stage = DiseaseFree;
  }
  void init();
  virtual void handleMessage(const cMessage *msg);
};

void CalibPerson::init() {
  if (R::runif(0, 1) < p2) diseasepot = true;
  else diseasepot = false;
  clinTime = 1000;
  stage = DiseaseFree;
  double lam1 = exp(R::rnorm(Lam1, sigm1));
  scheduleAt(R::rexp(lam1), toPrecursor);
  double x = R::runif(0, 1);
  scheduleAt((65 - 15 * log(-log(x))), toDeath);  // Gumbel
  for (int i = 10; i < 110; i = i + 10) scheduleAt(i, Count);
}

void CalibPerson::handleMessage(const cMessage *msg) {
  double ctime[] = {20, 40, 60, 80};
  int cind;
  g_events_handled++;
  if (msg->kind == toDeath) {
    stage = Death;
    clinTime = std::min(clinTime, now());
    for (unsigned int i = 0; i < 4; i++) {
      if (i < report["TimeAtRisk"].size()) report["TimeAtRisk"][i] += std::min(ctime[i], clinTime);
      else report["TimeAtRisk"].push_back(std::min(ctime[i], clinTime));
      if (clinTime < ctime[i]) break;
    }
    Sim::stop_simulation();
  } else if (msg->kind == toPrecursor) {
    stage = Precursor;
    if (diseasepot) {
      simtime_t dwellTime = now() + R::rexp(lam2);
      scheduleAt(dwellTime, toPreClinical);
    }
  } else if (msg->kind == toPreClinical) {
    stage = PreClinical;
    simtime_t dwellTime = now() + exp(R::rnorm(mu3, tau3 * mu3));
    scheduleAt(dwellTime, toClinical);
  } else if (msg->kind == toClinical) {
    stage = Clinical;
    clinTime = now();
  } else if (msg->kind == Count) {
    cind = std::min(9, int(now() / 10 - 1));
    std::string stagestr = stage_names[stage];
    if (report.find(stagestr) == report.end()) report[stagestr].assign(10, 0);
    report[stagestr][cind] += 1;
  }
}

}  // namespace calib

// callCalibrationSimulation (calibperson-r.cc:176-200).  `new Rng()` takes
// the package seed; first_person > 0 skips that many substreams first.
void run_calibration(int64_t first_person, int64_t n, const double par[6], msim_calibration *out) {
  using namespace calib;
  rlite::RNGkind(rlite::USER_UNIF);
  rlite::set_user_unif_rand(user_unif_rand);
  report.clear();
  rng = new Rng();
  rng->set();
  if (first_person > 0) rng->AdvanceSubstream(0, (int32_t)first_person);
  report.insert(std::make_pair("TimeAtRisk", std::vector<double>()));
  CalibPerson person;
  for (int64_t i = 0; i < n; i++) {
    person = CalibPerson(par, 0);
    rng->nextSubstream();
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
  delete rng;
  memset(out, 0, sizeof *out);
  for (int s = 0; s < 4; s++) {
    std::map<std::string, std::vector<double> >::iterator it = report.find(stage_names[s]);
    if (it == report.end()) continue;
    out->present[s] = 1;
    for (int k = 0; k < 10; k++) out->occupancy[s * 10 + k] = it->second[k];
  }
  std::vector<double> &tar = report["TimeAtRisk"];
  out->n_time_at_risk = (int32_t)tar.size();
  for (size_t k = 0; k < tar.size(); k++) out->time_at_risk[k] = tar[k];
}

// ---------------------------------------------------------- illness-death
namespace illnessDeath {

enum state_t { Healthy, Cancer };
enum event_t { toOtherDeath, toCancer, toCancerDeath };
static EventReport<short, short, double> report;
static double cure, zsd;

double b_weibull(double mean, double a, double rr = 1.0) { return mean / R::gammafn(1.0 + 1.0 / a) * pow(rr, -1.0 / a); }

class SimplePerson : public cProcess {
 public:
  state_t state;
  int id;
  double z;
  SimplePerson(const int i = 0) : id(i) {}
  void init();
  virtual void handleMessage(const cMessage *msg);
};

void SimplePerson::init() {
  state = Healthy;
  z = exp(R::rnorm(0.0, zsd));
  scheduleAt(R::rweibull(4.0, b_weibull(80.0, 4.0)), toOtherDeath);
  if (R::runif(0.0, 1.0) > cure) scheduleAt(R::rweibull(3.0, b_weibull(80.0, 3.0, z)), toCancer);
}

void SimplePerson::handleMessage(const cMessage *msg) {
  g_events_handled++;
  report.add(state, msg->kind, previousEventTime, now());
  switch (msg->kind) {
    case toOtherDeath:
    case toCancerDeath:
      Sim::stop_process();
      break;
    case toCancer:
      state = Cancer;
      RemoveKind(toOtherDeath);
      if (R::runif(0.0, 1.0) < 0.5) scheduleAt(now() + R::rweibull(1.0, 10.0), toCancerDeath);
      break;
    default:
      break;
  }
}

}  // namespace illnessDeath

static std::vector<double> ages_0_100_1e6() {
  std::vector<double> ages;
  for (int a = 0; a <= 100; a++) ages.push_back(a);
  ages.push_back(1.0e+6);
  return ages;
}

// callIllnessDeath (illness-death.cpp:73-94); draws from the global R stream.
void run_illness_death(int64_t n, double cure_, double zsd_, ReportTables *out) {
  using namespace illnessDeath;
  SimplePerson person;
  cure = cure_;
  zsd = zsd_;
  report.clear();
  report.setPartition(ages_0_100_1e6());
  for (int64_t i = 0; i < n; i++) {
    person = SimplePerson((int)i);
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
  flatten(report, out);
}

// ---------------------------------------------------------- simple-example
namespace simple1 {

enum state_t { Healthy, Cancer, Death };
enum event_t { toOtherDeath, toCancer, toCancerDeath };

class SimplePerson : public cProcess {
 public:
  int id;
  state_t state;
  RowLog *log;
  SimplePerson() : id(-1) {}
  void init();
  virtual void handleMessage(const cMessage *msg);
};

void SimplePerson::init() {
  id++;
  state = Healthy;
  double tm = R::rweibull(8.0, 85.0);
  scheduleAt(tm, toOtherDeath);
  scheduleAt(R::rweibull(3.0, 90.0), toCancer);
}

void SimplePerson::handleMessage(const cMessage *msg) {
  g_events_handled++;
  log->push(double(id), previousEventTime, now(), double(state), double(msg->kind));
  switch (msg->kind) {
    case toOtherDeath:
    case toCancerDeath:
      Sim::stop_process();
      break;
    case toCancer:
      state = Cancer;
      if (R::runif(0.0, 1.0) < 0.5) scheduleAt(now() + R::rweibull(2.0, 10.0), toCancerDeath);
      break;
    default:
      break;
  }
}

}  // namespace simple1

// callSimplePerson (simple-example.cc:72-83): ONE person object re-used, so
// previousEventTime/startTime carry over exactly as in the reference.
void run_simple_person(int64_t n, RowLog *log) {
  simple1::SimplePerson person;
  person.log = log;
  for (int64_t i = 0; i < n; i++) {
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
}

namespace simple2 {

enum state_t { Healthy, Cancer, Death };
enum event_t { toOtherDeath, toCancer, toCancerDeath };
static EventReport<short, short, double> report;

class SimplePerson : public cProcess {
 public:
  state_t state;
  int id;
  SimplePerson(const int i = 0) : id(i) {}
  void init();
  virtual void handleMessage(const cMessage *msg);
};

void SimplePerson::init() {
  state = Healthy;
  scheduleAt(R::rweibull(8.0, 85.0), toOtherDeath);
  scheduleAt(R::rweibull(3.0, 90.0), toCancer);
}

void SimplePerson::handleMessage(const cMessage *msg) {
  g_events_handled++;
  report.add(state, msg->kind, previousEventTime, now());
  switch (msg->kind) {
    case toOtherDeath:
    case toCancerDeath:
      Sim::stop_process();
      break;
    case toCancer:
      state = Cancer;
      if (R::runif(0.0, 1.0) < 0.5) scheduleAt(now() + R::rweibull(2.0, 10.0), toCancerDeath);
      break;
    default:
      break;
  }
}

}  // namespace simple2

void run_simple_person2(int64_t n, ReportTables *out) {  // simple-example2.cc:62-83
  using namespace simple2;
  SimplePerson person;
  report.clear();
  report.setPartition(ages_0_100_1e6());
  for (int64_t i = 0; i < n; i++) {
    person = SimplePerson((int)i);
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
  flatten(report, out);
}

// ------------------------------------------------- Sick-Sicker (README.org)
namespace sicksicker {

enum state_t { Healthy, Sick, Sicker, Dead };
enum event_t { toS1, toS2, toH, toD, toEOF };
typedef SummaryReport<short, short> Report;

static double rexpRate(double rate) { return R::rexp(1.0 / rate); }

class SickSicker : public cProcess {
 public:
  int id;
  state_t state;
  const SickSickerParam *param;
  Report *report;
  std::vector<double> *trace;
  Rng *rng;
  SickSicker(const SickSickerParam *param, Report *report) : id(-1), param(param), report(report), trace(0), rng(0) {}
  void init();
  void handleMessage(const cMessage *msg);
  void cancelEvents();
};

void SickSicker::init() {
  id++;
  state = Healthy;
  scheduleAt(rexpRate(param->r_HS1), toS1);
  scheduleAt(rexpRate(param->r_HD), toD);
  scheduleAt(31.0, toEOF);
}
void SickSicker::cancelEvents() {
  RemoveKind(toH);
  RemoveKind(toS1);
  RemoveKind(toS2);
  RemoveKind(toD);
}
void SickSicker::handleMessage(const cMessage *msg) {
  g_events_handled++;
  if (trace) {
    trace->push_back(id);
    trace->push_back(state);
    trace->push_back(msg->kind);
    trace->push_back(previousEventTime);
    trace->push_back(now());
  }
  report->add(state, msg->kind, this->previousEventTime, now(), id);
  switch (msg->kind) {
    case toH:
      state = Healthy;
      report->setUtility(param->u_H);
      report->setCost(param->c_H);
      cancelEvents();
      scheduleAt(now() + rexpRate(param->r_HS1), toS1);
      scheduleAt(now() + rexpRate(param->r_HD), toD);
      break;
    case toS1:
      state = Sick;
      report->setUtility(param->Trt ? param->u_Trt : param->u_S1);
      report->setCost(param->c_S1 + (param->Trt ? param->c_Trt : 0.0));
      cancelEvents();
      scheduleAt(now() + rexpRate(param->r_S1H), toH);
      scheduleAt(now() + rexpRate(param->r_S1S2), toS2);
      scheduleAt(now() + rexpRate(param->r_S1D), toD);
      break;
    case toS2:
      state = Sicker;
      report->setUtility(param->u_S2);
      report->setCost(param->c_S2 + (param->Trt ? param->c_Trt : 0.0));
      cancelEvents();
      scheduleAt(now() + rexpRate(param->r_S2D), toD);
      break;
    case toD:
    case toEOF:
      Sim::stop_simulation();
      break;
    default:
      break;
  }
}

}  // namespace sicksicker

// callSim of README.org:289-301.  substreams=false: global R stream (the
// documented run).  substreams=true: person i draws from substream
// first_person+i+1 of a stream made from the package seed — the regime the
// device engine runs this model in.
void run_sick_sicker(int64_t n, const SickSickerParam *param, bool indivp, bool substreams, int64_t first_person,
                     ReportTables *out, std::vector<double> *trace) {
  using namespace sicksicker;
  Report report((int)n, indivp);
  report.setPartition(0.0, 31.0, param->partitionBy);
  report.setDiscountRate(param->discountRate);
  SickSicker person(param, &report);
  person.trace = trace;
  Rng *rng = 0;
  if (substreams) {
    rlite::RNGkind(rlite::USER_UNIF);
    rlite::set_user_unif_rand(user_unif_rand);
    rng = new Rng();
    rng->set();
    // A shard that starts at person first_person > 0 must begin with the utility and cost the persons before it left
    // in the report (see below): the 256 persons before the shard are simulated first (one of them has set them with
    // probability 1 - 1e-980), then the tables are emptied and the report's utility and cost kept.
    const int64_t warm = first_person < 256 ? first_person : 256;
    if (first_person - warm > 0) rng->AdvanceSubstream(0, (int32_t)(first_person - warm));
    if (warm > 0) {
      Report scratch((int)warm, true);
      scratch.setPartition(0.0, 31.0, param->partitionBy);
      scratch.setDiscountRate(param->discountRate);
      SickSicker before(param, &scratch);
      for (int64_t i = 0; i < warm; i++) {
        rng->nextSubstream();
        Sim::create_process(&before);
        Sim::run_simulation();
        Sim::clear();
      }
      report.setUtility(scratch.utility);
      report.setCost(scratch.cost);
    }
  }
  // NB the reference never resets utility/cost between persons: the values
  // set by the previous person's last transition carry into the next
  // person's first interval (README.org:252-281).  Kept.
  for (int64_t i = 0; i < n; i++) {
    if (rng) rng->nextSubstream();
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
  if (rng) delete rng;
  flatten_summary(report, out);
}

// ------------------------- SummaryReport + addPointCost (test/test_microsimulation.R:782-865, 1006-1093)
// The dev script's "simple example" with cancer management: other-cause death and cancer onset as Weibulls, a
// one-off cost when cancer starts and a management cost every two years after (SummaryReport::addPointCost,
// microsimulation.h:1105-1112), intervals and point costs booked to the individual `id`.  Onset is the Weibull of the
// script's second version (:1043); the first version's gsm onset (:818) needs a fitted model the script does not keep.
namespace cancermgmt {

enum state_t { Healthy, Cancer, Death };
enum event_t { toOtherDeath, toCancer, toCancerDeath, toCancerManagement };
typedef SummaryReport<short, short> Report;

class SimplePerson : public cProcess {
 public:
  int id;
  state_t state;
  double utility, odShape;
  Report *report;
  SimplePerson(double odShape, Report *report) : id(-1), odShape(odShape), report(report) {}
  void init() {
    id++;
    state = Healthy;
    utility = 1.0;
    scheduleAt(R::rweibull(odShape, 85.0), toOtherDeath);
    scheduleAt(R::rweibull(3.0, 90.0), toCancer);
  }
  void handleMessage(const cMessage *msg) {
    g_events_handled++;
    report->add(state, msg->kind, this->previous(), now(), id);
    switch (msg->kind) {
      case toOtherDeath:
      case toCancerDeath:
        Sim::stop_simulation();
        break;
      case toCancer:
        state = Cancer;
        utility = 0.8;
        report->addPointCost(state, 15000.0, id);
        scheduleAt(now() + 2.0, toCancerManagement);
        if (R::runif(0.0, 1.0) < 0.5) scheduleAt(now() + R::rweibull(2.0, 10.0), toCancerDeath);
        break;
      case toCancerManagement:
        report->addPointCost(state, 1000.0, id);
        scheduleAt(now() + 2.0, toCancerManagement);
        break;
      default:
        break;
    }
  }
};

}  // namespace cancermgmt

// simulations(n, discountRate, odShape, partitionBy) of the dev script, person i on substream first_person+i+1 of a
// stream made from the package seed
void run_cancer_management(int64_t n, double odShape, double discountRate, double partitionBy, bool indivp, int64_t first_person,
                           ReportTables *out) {
  using namespace cancermgmt;
  Report report((int)n, indivp);
  report.setPartition(0.0, 100.0, partitionBy);
  report.setDiscountRate(discountRate);
  SimplePerson person(odShape, &report);
  rlite::RNGkind(rlite::USER_UNIF);
  rlite::set_user_unif_rand(user_unif_rand);
  Rng *rng = new Rng();
  rng->set();
  if (first_person > 0) rng->AdvanceSubstream(0, (int32_t)first_person);
  for (int64_t i = 0; i < n; i++) {
    rng->nextSubstream();
    Sim::create_process(&person);
    Sim::run_simulation();
    Sim::clear();
  }
  delete rng;
  flatten_summary(report, out);
}

// ----------------------------------------------- doc examples (examples.org)
namespace docs {

class tick_tock : public cProcess {
 public:
  double finish_time;
  std::vector<double> *times;
  void init() {
    scheduleAt(R::rexp(1.0), "tick");
    scheduleAt(R::rexp(10.0), "stop");
  }
  void handleMessage(const cMessage *msg) {
    times->push_back(now());
    msg->name == "tick"   ? scheduleAt(now() + R::rexp(1.0), "tock")
    : msg->name == "tock" ? scheduleAt(now() + R::rexp(1.0), "tick")
                          : Sim::stop_process();
  }
  void stop() { finish_time = now(); }
};

class illness_death : public cProcess {
 public:
  double QALY, utility, cumulative_costs, cost;
  std::vector<double> *times;
  std::vector<double> *out;
  void init() {
    cumulative_costs = QALY = 0.0;
    this->toHealthy();
  }
  void toHealthy() {
    utility = 0.95;
    cost = 1000.0;
    this->scheduleAt(now() + R::rexp(30.0), "onset");
    this->scheduleAt(now() + R::rexp(60.0), "death");
  }
  void toIllness() {
    utility = 0.95 * 0.8;
    cost = 10000.0;
    this->scheduleAt(now() + R::rexp(5.0), "recovered");
    this->scheduleAt(now() + R::rexp(10.0), "death");
  }
  void handleMessage(const cMessage *msg) {
    times->push_back(now());
    QALY += discountedInterval(utility, previous(), now(), 0.03);
    cumulative_costs += discountedInterval(cost, previous(), now(), 0.03);
    this->cancel_events();
    if (msg->name == "onset") toHealthy();
    else if (msg->name == "recovered") toIllness();
    else Sim::stop_process();
  }
  void stop() {
    out->push_back(QALY);
    out->push_back(cumulative_costs);
  }
};

}  // namespace docs

double run_tick_tock(std::vector<double> *times) {  // examples.org:152-158
  docs::tick_tock sim;
  sim.times = times;
  Sim::create_process(&sim);
  Sim::run_simulation();
  Sim::clear();
  return sim.finish_time;
}

void run_doc_illness_death(int n, std::vector<double> *times, std::vector<double> *qaly_cost) {  // examples.org:334-343
  docs::illness_death sim;
  sim.times = times;
  sim.out = qaly_cost;
  for (int i = 0; i < n; i++) {
    Sim::create_process(&sim);
    Sim::run_simulation();
    Sim::clear();
  }
}

// priority tie-break script, test/test_microsimulation.R:206-242
namespace prio {
class P : public cProcess {
 public:
  std::vector<double> *order;
  void init() {
    scheduleAt(10.0, (short)0, 1);
    scheduleAt(10.0, (short)1, 0);
    scheduleAt(10.0, (short)2, -1);
    scheduleAt(10.0, (short)3, 2);
  }
  void handleMessage(const cMessage *msg) { order->push_back(msg->kind); }
};
}  // namespace prio

void run_priority_demo(std::vector<double> *order) {
  prio::P p;
  p.order = order;
  Sim::create_process(&p);
  Sim::run_simulation();
  Sim::clear();
}

}  // namespace orc
