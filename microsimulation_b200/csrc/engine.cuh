// Device event engine: one simulated process (person) per thread.
//
// Mirrors, per thread, the reference's static single-threaded simulator:
//   ssim::Sim main loop, schedule, clear      src/ssim.cc:92-127, 148-235
//   Action ordering (time, then priority)     src/ssim.cc:51-63
//   heap<Action>::insert / pop_first          inst/include/heap.h:75-117
//   Sim::ignore_event (lazy cancellation)     src/ssim.cc:131-142
//   Sim::stop_process / stop_simulation       src/ssim.cc:241-253
// and the OMNeT++-style plugin surface models are written against:
//   cMessage{kind, priority}, cProcess::scheduleAt / previous / cancel*,
//   RemoveKind / CancelEvents, now()          inst/include/microsimulation.h:218-447
//
// The pending-event set is a fixed-capacity binary heap held in the thread's
// registers / local memory (CAP entries); it is the SAME algorithm as heap.h —
// strict `<` on sift-up, right child only if strictly smaller than the left —
// so ties in (time, priority) resolve exactly as on the CPU, including
// cancelled entries, which stay in the heap as A_Ignore and are skipped at pop
// time without advancing the clock.  Models with many pending actions may use OrderedEvents instead (below): an
// ordered array that pops in the heap's order whenever that order does not depend on the heap's shape, and says so
// when it might (the run is then repeated with the heap).
//
// Differences by design: there is one process per simulation (the built models
// never call cProcess::send), cMessage names (std::string) are not available on
// the device — kinds only — and a full heap is an error code, not UB.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "mrg32k3a.cuh"
#include "rmath.cuh"

namespace msim {

enum ActionType : int { A_Event = 0, A_Init = 1, A_Stop = 2, A_Ignore = 3 };

// meta word: bits 0-15 kind (int16), bits 16-29 priority + 8192, bits 30-31 type
MSIM_DEV int pack_meta(int type, int kind, int priority) {
  return (int)(((uint32_t)kind & 0xffffu) | ((((uint32_t)(priority + 8192)) & 0x3fffu) << 16) | ((uint32_t)type << 30));
}
MSIM_DEV int meta_kind(int m) { return (int)(int16_t)((uint32_t)m & 0xffffu); }
MSIM_DEV int meta_prio(int m) { return (int)(((uint32_t)m >> 16) & 0x3fffu) - 8192; }
MSIM_DEV int meta_type(int m) { return (int)((uint32_t)m >> 30); }

// Storage of the generic heap: CAP > 4 keeps the entries in the thread's own
// arrays (local memory); CAP < 0 means -CAP entries in memory the caller binds
// (shared memory: slot k of this thread at base[k * stride]), which the
// calibration model uses — its 12 pending actions in local memory were that
// kernel's top stall.
template <int CAP>
struct HeapStore {
  double t_[CAP];
  int m_[CAP];
  static constexpr int capacity = CAP;
  MSIM_DEV double &T(int k) { return t_[k]; }
  MSIM_DEV int &M(int k) { return m_[k]; }
};
template <int CAP>
struct HeapStoreBound {
  double *t_;
  int *m_;
  int stride_;
  static constexpr int capacity = CAP;
  MSIM_DEV void bind(double *t, int *m, int stride) {
    t_ = t;
    m_ = m;
    stride_ = stride;
  }
  MSIM_DEV double &T(int k) { return t_[k * stride_]; }
  MSIM_DEV int &M(int k) { return m_[k * stride_]; }
};

template <int CAP>
struct EventHeap : std::conditional<(CAP > 0), HeapStore<(CAP > 0 ? CAP : 1)>, HeapStoreBound<(CAP > 0 ? 1 : -CAP)> >::type {
  typedef typename std::conditional<(CAP > 0), HeapStore<(CAP > 0 ? CAP : 1)>, HeapStoreBound<(CAP > 0 ? 1 : -CAP)> >::type Store;
  using Store::M;
  using Store::T;
  int n;
  static constexpr bool tie = false;       // ties resolve as in heap.h: nothing to report
  static constexpr bool compacts = false;  // cancelled entries stay (A_Ignore), as in ssim.cc:131-142

  MSIM_DEV void reset() { n = 0; }
  MSIM_DEV bool before(double ta, int ma, double tb, int mb) const {
    return ta < tb || (ta == tb && meta_prio(ma) < meta_prio(mb));  // ssim.cc:60-62
  }

  // heap.h:75-88
  MSIM_DEV bool insert(double time, int m) {
    if (n >= Store::capacity) return false;
    int k = n++;
    while (k > 0) {
      int up = (k - 1) >> 1;
      double tu = T(up);
      int mu = M(up);
      if (!before(time, m, tu, mu)) break;
      T(k) = tu;
      M(k) = mu;
      k = up;
    }
    T(k) = time;
    M(k) = m;
    return true;
  }

  // heap.h:90-117 (caller guarantees n > 0)
  MSIM_DEV void pop_first(double &time, int &m) {
    time = T(0);
    m = M(0);
    int last = --n;  // index of the element that moves to the root
    if (last == 0) return;
    double tx = T(last);
    int mx = M(last);
    last--;  // new last valid index
    int k = 0;
    for (;;) {
      int c = 2 * k + 1;
      if (c > last) break;
      double tc = T(c);
      int mc = M(c);
      if (c + 1 <= last) {
        double tr = T(c + 1);
        int mr = M(c + 1);
        if (before(tr, mr, tc, mc)) {
          c = c + 1;
          tc = tr;
          mc = mr;
        }
      }
      if (!before(tc, mc, tx, mx)) break;
      T(k) = tc;
      M(k) = mc;
      k = c;
    }
    T(k) = tx;
    M(k) = mx;
  }

  // visit the pending entries in array order (Sim::ignore_event's scan)
  template <class F>
  MSIM_DEV void for_each_meta(F f) {
    for (int i = 0; i < n; i++) f(M(i));
  }
};

// CAP = 4: the same heap, written out as straight-line code over four scalar
// slots so that it lives entirely in registers (no dynamically indexed local
// array).  Every built model except the calibration example keeps at most
// three actions pending.  Comparisons and moves are exactly those heap.h
// performs for sizes 1..4, so ties resolve identically.
template <>
struct EventHeap<4> {
  double t0, t1, t2, t3;
  int m0, m1, m2, m3;
  int n;
  static constexpr bool tie = false;
  static constexpr bool compacts = false;

  MSIM_DEV void reset() { n = 0; }

  MSIM_DEV static bool before(double ta, int ma, double tb, int mb) {
    return ta < tb || (ta == tb && meta_prio(ma) < meta_prio(mb));
  }
  MSIM_DEV static void swap(double &ta, int &ma, double &tb, int &mb) {
    double t = ta;
    ta = tb;
    tb = t;
    int m = ma;
    ma = mb;
    mb = m;
  }

  MSIM_DEV bool insert(double time, int m) {
    if (n == 0) {
      t0 = time;
      m0 = m;
    } else if (n == 1) {  // slot 1, parent 0
      t1 = time;
      m1 = m;
      if (before(t1, m1, t0, m0)) swap(t0, m0, t1, m1);
    } else if (n == 2) {  // slot 2, parent 0
      t2 = time;
      m2 = m;
      if (before(t2, m2, t0, m0)) swap(t0, m0, t2, m2);
    } else if (n == 3) {  // slot 3, parent 1, grandparent 0
      t3 = time;
      m3 = m;
      if (before(t3, m3, t1, m1)) {
        swap(t1, m1, t3, m3);
        if (before(t1, m1, t0, m0)) swap(t0, m0, t1, m1);
      }
    } else {
      return false;
    }
    n++;
    return true;
  }

  MSIM_DEV void pop_first(double &time, int &m) {
    time = t0;
    m = m0;
    if (n == 2) {  // last (slot 1) becomes the root, nothing below it
      t0 = t1;
      m0 = m1;
    } else if (n == 3) {  // slot 2 -> root; only child is slot 1
      t0 = t2;
      m0 = m2;
      if (before(t1, m1, t0, m0)) swap(t0, m0, t1, m1);
    } else if (n == 4) {  // slot 3 -> root; children 1 and 2, both leaves afterwards
      t0 = t3;
      m0 = m3;
      if (before(t2, m2, t1, m1)) {
        if (before(t2, m2, t0, m0)) swap(t0, m0, t2, m2);
      } else {
        if (before(t1, m1, t0, m0)) swap(t0, m0, t1, m1);
      }
    }
    n--;
  }

  // visit the pending entries in array order (Sim::ignore_event's scan)
  template <class F>
  MSIM_DEV void for_each_meta(F f) {
    if (n > 0) f(m0);
    if (n > 1) f(m1);
    if (n > 2) f(m2);
    if (n > 3) f(m3);
  }
};

// The pending-event set as an ORDERED array, for models that keep many actions pending (the calibration model: 12).
// Popping the minimum is then a load (the binary heap's sift-down over 12 entries was 44% of that kernel's
// instructions).  It pops in exactly the heap's order as long as no two pending actions compare equal — neither
// before the other in (time, priority): then the minimum is unique at every pop, whatever the data structure.  Where
// two do compare equal, heap.h's order depends on the heap's shape, which this queue does not model: it sets `tie`
// instead, and the caller repeats the run with EventHeap (host side: msim_calibration_sweep).  Event times are
// continuous draws, so this essentially never happens; exactness does not rest on that.
// Entries are ascending in [head, tail) of caller-bound memory (the default) or of the thread's own small array
// (StoreT = HeapStore<CAP>); insertion scans from the late end.
// Cancelled actions are REMOVED instead of lingering as A_Ignore: popping an ignored action has no effect in the
// reference (no clock update, no dispatch: ssim.cc:171-175), so the live actions' pop order — all that is
// observable — is the same, and a model that cancels at every transition (Sick-Sicker) keeps a handful of entries
// instead of hundreds.
template <int CAP, class StoreT = HeapStoreBound<CAP> >
struct OrderedEvents : StoreT {
  typedef StoreT Store;
  using Store::M;
  using Store::T;
  int n, head, tail;
  bool tie;
  static constexpr bool compacts = true;

  // drop the entries whose meta word satisfies pred
  template <class Pred>
  MSIM_DEV void remove_if(Pred pred) {
    int w = head;
    for (int i = head; i < tail; i++) {
      const int m = M(i);
      if (!pred(m)) {
        if (w != i) {
          T(w) = T(i);
          M(w) = m;
        }
        w++;
      }
    }
    tail = w;
    n = tail - head;
  }

  MSIM_DEV void reset() {
    n = head = tail = 0;
    tie = false;
  }
  MSIM_DEV static bool before(double ta, int ma, double tb, int mb) {
    return ta < tb || (ta == tb && meta_prio(ma) < meta_prio(mb));  // ssim.cc:60-62
  }
  MSIM_DEV bool insert(double time, int m) {
    if (tail == CAP) {
      if (head == 0) return false;
      for (int i = 0; i < n; i++) {  // make room at the late end
        T(i) = T(head + i);
        M(i) = M(head + i);
      }
      head = 0;
      tail = n;
    }
    int k = tail++;
    n++;
    while (k > head) {
      const double tu = T(k - 1);
      const int mu = M(k - 1);
      if (!before(time, m, tu, mu)) {
        if (!before(tu, mu, time, m)) tie = true;
        break;
      }
      T(k) = tu;
      M(k) = mu;
      k--;
    }
    T(k) = time;
    M(k) = m;
    return true;
  }
  MSIM_DEV void pop_first(double &time, int &m) {  // caller guarantees n > 0
    time = T(head);
    m = M(head);
    head++;
    n--;
  }
  template <class F>
  MSIM_DEV void for_each_meta(F f) {
    for (int i = head; i < tail; i++) f(M(i));
  }
};

// CRTP base: `struct Person : cProcess<Person, Src, 4> { void init(); void
// handleMessage(int kind); }`.  Handler bodies read like the reference's.
template <class Derived, class Src, int CAP, class Actions = EventHeap<CAP> >
struct cProcess {
  Actions actions;
  double clock_;
  double startTime, previousEventTime;
  bool running_, terminated_;
  int error_;  // MSIM_ERR_HEAP if a schedule did not fit
  RMath<Src> R;

  MSIM_DEV explicit cProcess(Src *src) : R(src) {
    actions.reset();
    clock_ = 0.0;
    startTime = previousEventTime = 0.0;
    running_ = false;
    terminated_ = false;
    error_ = 0;
  }

  MSIM_DEV double now() const { return clock_; }  // microsimulation.cc:10-12
  MSIM_DEV double simTime() const { return clock_; }
  MSIM_DEV double previous() const { return previousEventTime; }

  // cProcess::scheduleAt (microsimulation.h:342-360) + Sim::self_signal_event
  // (ssim.cc:269-271) + SimImpl::schedule (ssim.cc:92-98): the stored time is
  // clock + (t - clock), two roundings, as in the reference.
  MSIM_DEV void scheduleAt(double t, int kind, int priority = 0) {
    double delay = t - clock_;
    if (!actions.insert(clock_ + delay, pack_meta(A_Event, kind, priority))) error_ = -4;
  }

  // Cancel(pred) -> Sim::ignore_event: mark matching A_Event entries A_Ignore
  template <class Pred>
  MSIM_DEV void Cancel(Pred pred) {
    if constexpr (Actions::compacts) {
      actions.remove_if([&](int m) { return meta_type(m) == A_Event && pred(meta_kind(m)); });
    } else {
      actions.for_each_meta([&](int &m) {
        if (meta_type(m) == A_Event && pred(meta_kind(m))) m = (int)(((uint32_t)m & 0x3fffffffu) | ((uint32_t)A_Ignore << 30));
      });
    }
  }
  MSIM_DEV void RemoveKind(int kind) {
    Cancel([kind](int k) { return k == kind; });
  }
  MSIM_DEV void CancelKind(int kind) { RemoveKind(kind); }
  MSIM_DEV void cancel_kind(int kind) { RemoveKind(kind); }  // single process: same set
  MSIM_DEV void CancelEvents() {
    Cancel([](int) { return true; });
  }
  MSIM_DEV void cancel_events() { CancelEvents(); }

  // Sim::stop_process(): A_Stop at the current time; the process is marked
  // terminated when it is popped and later events are drained (ssim.cc:189-213)
  MSIM_DEV void stop_process() {
    if (!actions.insert(clock_, pack_meta(A_Stop, -1, 0))) error_ = -4;
  }
  // Sim::stop_simulation(): the loop exits before the next pop (ssim.cc:251)
  MSIM_DEV void stop_simulation() { running_ = false; }

  MSIM_DEV void stop() {}  // Process::stop hook; models may shadow it

  // create_process + run_simulation + clear for this one process.
  // create_process pushes A_Init at t = 0 into an empty heap, so the first
  // pop is always that action: initialize() runs first (microsimulation.h:430).
  MSIM_DEV void run() {
    begin();
    loop();
  }

  // For models whose init() has been carried out elsewhere (its draws taken and
  // its events scheduled on this object): the state right after A_Init.
  MSIM_DEV void resume_after_init() {
    running_ = true;
    previousEventTime = startTime;
    loop();
  }

  // Sim::run_simulation's pop/dispatch loop (ssim.cc:162-232)
  MSIM_DEV void loop() {
    if (alive())
      while (step()) {
      }
    running_ = false;
  }

  // The same loop in pieces, for schedulers that interleave persons: begin() is everything up
  // to the loop's first test after A_Init, alive() that test, step() one pass (one pop; an event is dispatched unless
  // the action is ignored or the process has terminated) followed by the test for the next pass.
  MSIM_DEV void begin() {
    Derived *self = static_cast<Derived *>(this);
    running_ = true;
    self->init();
    previousEventTime = startTime;
  }
  MSIM_DEV bool alive() const { return running_ && actions.n > 0; }
  MSIM_DEV bool step() {  // requires alive()
    Derived *self = static_cast<Derived *>(this);
    double t;
    int m;
    actions.pop_first(t, m);
    int type = meta_type(m);
    if (type != A_Ignore) {  // an ignored action leaves the clock untouched (ssim.cc:171-175)
      clock_ = t;
      if (!terminated_) {  // else drained (ssim.cc:189-192)
        if (type == A_Event) {
          self->handleMessage(meta_kind(m));
          previousEventTime = clock_;  // microsimulation.h:441
        } else {                       // A_Stop
          self->stop();
          terminated_ = true;
        }
      }
    }
    return alive();
  }
};

// uniform source: the person's MRG32k3a substream
struct MrgSource {
  Mrg32k3a g;
  MSIM_DEV double unif_rand() { return g.u01(); }
};

}  // namespace msim
