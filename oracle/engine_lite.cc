// ORACLE — TEST INFRASTRUCTURE ONLY (see engine_lite.h).
// Restatement of src/ssim.cc + inst/include/heap.h of the reference.
#include "engine_lite.h"

#include <vector>

namespace ssim {

namespace {

enum Kind { K_EVENT, K_INIT, K_STOP, K_IGNORE };  // src/ssim.cc:44-49

struct Pending {  // src/ssim.cc:51-63 "Action"
  Time when;
  Kind kind;
  ProcessId pid;
  const Event *ev;
};

// (time, priority) order, src/ssim.cc:60-62.  Like the reference this reads
// ev->priority only when the times tie.
inline bool before(const Pending &x, const Pending &y) {
  if (x.when < y.when) return true;
  if (x.when == y.when) return x.ev->priority < y.ev->priority;
  return false;
}

struct Slot {  // src/ssim.cc:67-75 "PDescr"
  Process *proc;
  bool dead;
  Time free_at;
};

}  // namespace

struct EngineState {
  std::vector<Pending> q;  // binary heap, root at index 0 (heap.h:75-117)
  std::vector<Slot> procs;
  Time horizon = INIT_TIME;
  Time t = INIT_TIME;
  ProcessId cur = NULL_PROCESSID;
  bool live = false;
  bool in_loop = false;

  // heap.h:75-88 — append, then swap upward while strictly smaller than parent
  void push(const Pending &p) {
    q.push_back(p);
    size_t k = q.size() - 1;
    while (k > 0) {
      size_t up = (k - 1) / 2;
      if (!before(q[k], q[up])) break;
      Pending tmp = q[k];
      q[k] = q[up];
      q[up] = tmp;
      k = up;
    }
  }

  // heap.h:90-117 — move last to root, sink; right child chosen only if it
  // is strictly smaller than the left one; stop unless child < node.
  Pending pop() {
    Pending top = q[0];
    size_t last = q.size() - 1;
    if (last == 0) {
      q.pop_back();
      return top;
    }
    q[0] = q[last];
    q.pop_back();
    last--;
    size_t k = 0;
    for (;;) {
      size_t c = 2 * k + 1;
      if (c > last) break;
      if (c + 1 <= last && before(q[c + 1], q[c])) c = c + 1;
      if (!before(q[c], q[k])) break;
      Pending tmp = q[k];
      q[k] = q[c];
      q[c] = tmp;
      k = c;
    }
    return top;
  }

  void hold(const Event *e) {
    if (e) ++(e->refcount);
  }
  void drop(const Event *e) {
    if (e && --(e->refcount) == 0) delete e;
  }
  // src/ssim.cc:92-105
  void at_delay(Time d, Kind k, ProcessId p, const Event *e) {
    hold(e);
    push(Pending{t + d, k, p, e});
  }
  void at_now(Kind k, ProcessId p, const Event *e) {
    hold(e);
    push(Pending{t, k, p, e});
  }
};

static EngineState S;

ProcessId ProcessWithPId::activate() {  // src/ssim.cc:285-291
  if (process_id == NULL_PROCESSID) return process_id = Sim::create_process(this);
  return NULL_PROCESSID;
}

ProcessId Sim::create_process(Process *p) {  // src/ssim.cc:108-113
  S.procs.push_back(Slot{p, false, INIT_TIME});
  ProcessId id = (ProcessId)S.procs.size() - 1;
  S.at_now(K_INIT, id, 0);
  return id;
}

void Sim::clear() {  // src/ssim.cc:115-127
  S.live = false;
  S.t = INIT_TIME;
  S.cur = NULL_PROCESSID;
  S.procs.clear();
  for (size_t i = 0; i < S.q.size(); i++) S.drop(S.q[i].ev);
  S.q.clear();
}

void Sim::ignore_event(EventPredicate pred) {  // src/ssim.cc:131-142
  for (size_t i = 0; i < S.q.size(); i++) {
    Pending &p = S.q[i];
    if (p.kind == K_EVENT && p.ev != 0 && pred(p.ev)) p.kind = K_IGNORE;
  }
}

void Sim::run_simulation() {  // src/ssim.cc:148-235
  if (S.in_loop) return;
  S.in_loop = true;
  S.live = true;
  while (S.live && !S.q.empty()) {
    Pending a = S.pop();
    if (a.kind == K_IGNORE) {  // dropped without touching the clock (:171-175)
      S.drop(a.ev);
      continue;
    }
    S.t = a.when;
    if (S.horizon != INIT_TIME && S.t > S.horizon) break;
    S.cur = a.pid;
    Slot &s = S.procs[S.cur];
    if (s.dead) {
      // terminated process: event silently discarded (no error handler is
      // ever installed by the reference's models)
    } else if (S.t < s.free_at) {
      // busy: likewise discarded
    } else {
      switch (a.kind) {
        case K_EVENT:
          s.proc->process_event(a.ev);
          break;
        case K_INIT:
          s.proc->initialize();
          break;
        case K_STOP:
          s.proc->stop();
          S.procs[S.cur].dead = true;
          break;
        default:
          break;
      }
      S.procs[S.cur].free_at = S.t;
    }
    S.drop(a.ev);
  }
  S.in_loop = false;
  S.live = false;
}

void Sim::set_stop_time(Time t) { S.horizon = t; }
void Sim::stop_process() { S.at_now(K_STOP, S.cur, 0); }  // src/ssim.cc:241-243
int Sim::stop_process(ProcessId pid) {
  if (S.procs[pid].dead) return -1;
  S.at_now(K_STOP, pid, 0);
  return 0;
}
void Sim::stop_simulation() { S.live = false; }  // src/ssim.cc:251-253
void Sim::advance_delay(Time d) {
  if (S.live) S.t += d;
}
ProcessId Sim::this_process() { return S.cur; }
Time Sim::clock() { return S.t; }
void Sim::self_signal_event(const Event *e) { S.at_now(K_EVENT, S.cur, e); }
void Sim::self_signal_event(const Event *e, Time d) { S.at_delay(d, K_EVENT, S.cur, e); }
void Sim::signal_event(ProcessId p, const Event *e) { S.at_now(K_EVENT, p, e); }
void Sim::signal_event(ProcessId p, const Event *e, Time d) { S.at_delay(d, K_EVENT, p, e); }

}  // namespace ssim
