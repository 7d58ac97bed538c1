// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// CPU restatement of the reference's discrete-event engine, used when the
// oracle is built WITHOUT the reference tree (the GPU box has no
// /root/reference).  It restates, in our own code:
//   * the public simulator surface   inst/include/siena/ssim.h:108-562
//   * the action ordering            src/ssim.cc:51-63   (time, then event priority)
//   * schedule / schedule_now        src/ssim.cc:92-105
//   * create_process / clear         src/ssim.cc:108-127
//   * ignore_event (lazy cancel)     src/ssim.cc:131-142
//   * the main loop                  src/ssim.cc:148-235
//   * stop_process / stop_simulation src/ssim.cc:241-253
//   * the 0-based binary heap        inst/include/heap.h:75-117
// The `oracle/_ref` build swaps this file for the reference's own
// siena/ssim.h + src/ssim.cc + heap.h, compiled where they lie; tests assert
// both builds give identical results (tests/test_oracle_ref.py).
#ifndef MSIM_ORACLE_ENGINE_LITE_H
#define MSIM_ORACLE_ENGINE_LITE_H

#include <functional>
#include <string>

namespace ssim {

typedef int ProcessId;
const ProcessId NULL_PROCESSID = -1;
typedef double Time;
const Time INIT_TIME = 0;

class Event {
 public:
  Event() : priority(0), refcount(0) {}
  virtual ~Event() {}
  virtual std::string str() const { return "(event)"; }
  short priority;

 private:
  mutable unsigned refcount;
  friend class Sim;
  friend struct EngineState;
};

typedef std::function<bool(const Event *)> EventPredicate;

class Process {
 public:
  virtual ~Process() {}
  virtual void initialize(void) {}
  virtual void process_event(const Event *) {}
  virtual void stop(void) {}
};

class ProcessWithPId : public Process {
 public:
  ProcessWithPId() : process_id(NULL_PROCESSID) {}
  ProcessId activate();
  ProcessId pid() const { return process_id; }

 private:
  ProcessId process_id;
};

class Sim {
 public:
  static ProcessId create_process(Process *p);
  static int stop_process(ProcessId pid);
  static void stop_process();
  static void clear();
  static void self_signal_event(const Event *e);
  static void self_signal_event(const Event *e, Time delay);
  static void signal_event(ProcessId p, const Event *e);
  static void signal_event(ProcessId p, const Event *e, Time d);
  static void advance_delay(Time d);
  static ProcessId this_process();
  static Time clock();
  static void run_simulation();
  static void stop_simulation();
  static void set_stop_time(Time t = INIT_TIME);
  static void ignore_event(EventPredicate pred);
};

}  // namespace ssim

#endif
