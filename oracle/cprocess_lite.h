// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Restatement of the Rcpp-dependent model-plugin layer of the reference,
// which cannot be compiled here (RcppCommon.h is absent):
//   cMessage                    inst/include/microsimulation.h:218-245
//   Cancel / RemoveKind / ...   inst/include/microsimulation.h:252-302
//   cProcess                    inst/include/microsimulation.h:325-447
//   now() / simTime()           src/microsimulation.cc:10-16
//   rweibullHR                  src/microsimulation.cc:6-8
//   Rng + user_unif_rand hook   inst/include/microsimulation.h:742-757,
//                               src/microsimulation.cc:19-85
//   discountedInterval/Point    inst/include/microsimulation.h:811-836
// It sits on either engine: engine_lite.h/rngstream_lite.h (portable oracle)
// or the reference's own siena/ssim.h + RngStream.h (oracle/_ref build,
// -DMSIM_ORACLE_USE_REFERENCE).
#ifndef MSIM_ORACLE_CPROCESS_LITE_H
#define MSIM_ORACLE_CPROCESS_LITE_H

#ifdef MSIM_ORACLE_USE_REFERENCE
#include <RngStream.h>
#include <siena/ssim.h>
#else
#include "engine_lite.h"
#include "rngstream_lite.h"
#endif

#include <math.h>
#include <functional>
#include <sstream>
#include <string>

#include "rmath_lite.h"

namespace ssim {

using std::string;

#define WithRNG(rng, expr) (rng->set(), expr)

class cMessage : public ssim::Event {
 public:
  cMessage(const short k = -1, const string n = "", const ProcessId process_id = -1,
           const ProcessId sender_process_id = -1, const short priority = 0)
      : ssim::Event(), kind(k), name(n), sendingTime(-1.0), timestamp(0), process_id(process_id),
        sender_process_id(sender_process_id) {
    this->priority = priority;
  }
  short kind;
  string name;
  Time sendingTime, timestamp;
  ProcessId process_id, sender_process_id;
  string str() const {
    std::ostringstream os;
    os << "kind=" << kind << ",name=" << name;
    return os.str();
  }
};

// lazy cancellation across all processes: matching entries become A_Ignore
inline void Cancel(std::function<bool(const cMessage *msg)> pred) {
  Sim::ignore_event([pred](const Event *e) {
    const cMessage *msg = dynamic_cast<const cMessage *>(e);
    return (msg != 0 && pred(msg));
  });
}
inline void RemoveKind(short kind) {
  Cancel([kind](const cMessage *msg) { return msg->kind == kind; });
}
inline void CancelKind(short kind) { RemoveKind(kind); }
inline void RemoveName(string name) {
  Cancel([name](const cMessage *msg) { return msg->name == name; });
}
inline void CancelName(string name) { RemoveName(name); }
inline void CancelEvents() {
  Cancel([](const cMessage *) { return true; });
}

class cProcess : public ssim::ProcessWithPId {
 public:
  Time startTime, previousEventTime;
  cProcess(Time startTime = Time(0.0)) : startTime(startTime), previousEventTime(startTime) {}
  virtual void handleMessage(const cMessage *msg) = 0;
  virtual void init() = 0;
  // absolute time t -> delay t-clock -> the engine stores clock + delay
  virtual void scheduleAt(Time t, cMessage *msg, short priority = 0) {
    msg->timestamp = t;
    msg->sendingTime = Sim::clock();
    msg->process_id = msg->sender_process_id = pid();
    msg->priority = priority;
    Sim::self_signal_event(msg, t - Sim::clock());
  }
  virtual void scheduleAt(Time t, string name, short priority = 0) {
    scheduleAt(t, new cMessage(-1, name), priority);
  }
  virtual void scheduleAt(Time t, const char *name, short priority = 0) {
    scheduleAt(t, new cMessage(-1, string(name)), priority);
  }
  virtual void scheduleAt(Time t, short k, short priority = 0) {
    scheduleAt(t, new cMessage(k, ""), priority);
  }
  virtual void scheduleAt(Time t, int k, short priority = 0) {  // enum kinds
    scheduleAt(t, new cMessage((short)k, ""), priority);
  }
  virtual void send(ProcessId process_id, Time t, cMessage *msg, short priority = 0) {
    msg->timestamp = t;
    msg->process_id = process_id;
    msg->sender_process_id = pid();
    msg->sendingTime = Sim::clock();
    msg->priority = priority;
    Sim::signal_event(process_id, msg, t - Sim::clock());
  }
  void cancel(std::function<bool(const cMessage *msg)> pred) {
    Cancel([pred, this](const cMessage *msg) { return pred(msg) && msg->process_id == this->pid(); });
  }
  void cancel_kind(short kind) {
    cancel([kind](const cMessage *msg) { return msg->kind == kind; });
  }
  void cancel_name(string name) {
    cancel([name](const cMessage *msg) { return msg->name == name; });
  }
  void cancel_events() {
    cancel([](const cMessage *) { return true; });
  }
  double previous() { return previousEventTime; }
  void initialize() {
    init();
    previousEventTime = startTime;
  }
  virtual void process_event(const ssim::Event *e) {
    const cMessage *msg = dynamic_cast<const cMessage *>(e);
    if (msg != 0) {
      handleMessage(msg);
      previousEventTime = Sim::clock();
    }
  }
};

typedef Time simtime_t;
Time simTime();
Time now();
double rweibullHR(double shape, double scale, double hr);

// RngStream that can be selected as the source of unif_rand()
class Rng : public RngStream {
 public:
  Rng();
  virtual ~Rng();
  void seed(const double seed[6]) { SetSeed(seed); }
  void set();
  void nextSubstream() { ResetNextSubstream(); }
  int id;
};

extern "C" {
void r_create_current_stream();
void r_remove_current_stream();
void r_set_user_random_seed(double *seed);
void r_get_user_random_seed(double *seed);
void r_next_rng_substream();
void r_rng_advance_substream(double *seed, int *n);
double *user_unif_rand();
}

inline double discountedInterval(double start, double end, double discountRate) {
  if (discountRate == 0.0) return end - start;
  return (pow(1.0 + discountRate, -start) - pow(1.0 + discountRate, -end)) / log(1.0 + discountRate);
}
inline double discountedInterval(double y, double start, double end, double discountRate) {
  if (discountRate == 0.0) return y * (end - start);
  return y * (pow(1.0 + discountRate, -start) - pow(1.0 + discountRate, -end)) / log(1.0 + discountRate);
}
inline double discountedPoint(double time, double discountRate) {
  return discountRate <= 0.0 ? 1.0 : pow(1.0 + discountRate, -time);
}
inline double discountedPoint(double y, double time, double discountRate) {
  return discountRate <= 0.0 ? y : y * pow(1.0 + discountRate, -time);
}

}  // namespace ssim

#endif
