// Staged execution of the FHCRC natural-history model (src/person-r.cc) —
// the schedule the person-r entry points use on the GPU.
//
// Why.  Run one-person-per-thread straight through, this model leaves ~30% of
// a warp's lanes active (ncu: 9.6 threads per warp instruction, profiles/r01a):
// 78% of persons draw three uniforms and die of other causes, while the rest
// need a Weibull onset time (log + pow) and then, for half of those, a chain of
// progression / diagnosis events with more log + pow.  Nearly every warp holds
// a few such persons, so every warp walks the long code with a few lanes on.
//
// How.  A person's life is cut at the two points where it may need the
// expensive arithmetic, and persons that reach a cut are parked, with their
// RngStream state, in a per-warp queue in shared memory.  Each pass of a
// warp's loop runs ONE stage on up to 32 persons —
//   stage A   32 NEW persons (all lanes): substream jump, init()'s draws;
//             persons without an onset draw are ready (a single toDeath event);
//   stage B1  when >= 32 persons wait for their onset time: log + pow on all
//             lanes; persons who die before onset are ready;
//   stage B2  when >= 32 persons wait with onset before death: ready;
// — and then, in ONE place for all stages (run_life), the ready persons run
// the event loop (Sim::run_simulation's, engine.cuh) from just after A_Init.
// No block-level synchronisation is involved: queues are per warp, and pushes
// and pops are warp ballots.  Every draw a person takes, and the order it takes
// them in, is the reference's; only WHEN the arithmetic happens moves.  A person's
// stages run through the same Person<Env> object (models.cuh) the generic kernel
// uses: scheduleAt(), the 4-slot heap and handleMessage() are shared code, so
// ties and pop order are the reference's by construction.
//
// Row log.  Rows must come out in person order but persons now finish out of
// order, so the simulation writes, per person, its FIRST row (end time + a
// 16-bit tag: kind, state, number of further rows) to arrays indexed by person,
// and appends the further rows (0.23 per person) to a small overflow region
// owned by the person's chunk of ROW_CHUNK consecutive persons, tagged with
// (person, row number).  A bandwidth-bound pass per chunk then merges both in
// shared memory and streams the final five f64 columns out fully coalesced
// (reads 10 B per person + 20 B per further row, writes 40 B per row).
#pragma once
#include "models.cuh"
#include "report.cuh"

namespace msim {
constexpr int MSIM_ERR_MODEL_DEV = -8;  // include/msim.h MSIM_ERR_MODEL

namespace person_r {

// a person parked after stage A: init()'s uniforms taken, onset time not yet computed
struct ParkedOnset {
  Mrg32k3a g;  // RngStream state after init()'s draws
  double u_onset, t_death;
};
// a person whose init() events are known: toLocalised at t_onset (if has_onset) and toDeath at t_death
struct Life {
  Mrg32k3a g;
  double t_onset, t_death;
  bool has_onset;
};

// Stage A: init()'s draws.  Returns true if the person can run now (no onset draw: a single toDeath event).
template <class Env>
MSIM_DEV bool stage_a(Env &env, ParkedOnset *park, Life *life) {
  RMath<typename Env::Src> R(&env.src);
  double u_onset = 0.0, t_death = 0.0;
  const bool onset = Person<Env>::init_draws(R, &u_onset, &t_death);
  if (!onset) {
    life->g = env.src.g;
    life->t_onset = 0.0;
    life->t_death = t_death;
    life->has_onset = false;
    return true;
  }
  park->g = env.src.g;
  park->u_onset = u_onset;
  park->t_death = t_death;
  return false;
}

// Stage B1: the onset time.  Returns true if the person can run now: death pops before onset (strictly
// earlier, because toLocalised was inserted first — heap.h:75-88), so nothing but toDeath will be handled.
template <class Env>
MSIM_DEV bool stage_b1(const Consts &K, const ParkedOnset &in, Life *life) {
  life->g = in.g;
  life->t_onset = Person<Env>::onset_time(K, in.u_onset);
  life->t_death = in.t_death;
  life->has_onset = true;
  return in.t_death < life->t_onset;
}

// The one place a person's events run: init()'s scheduleAt calls in init()'s order (person-r.cc:100-102), then
// Sim::run_simulation's loop from just after A_Init.
template <bool FOLD_SINGLE, class Env>
MSIM_DEV void run_life(Env &env, const Consts &K, int64_t id, const Life &life, int *error) {
  env.src.g = life.g;
  Person<Env> person(env, K, id);
  // Writing the one-event case (78% of persons) apart lets the compiler fold heap and dispatch away for it.  (In round
  // 1 the second copy of the handler and report code cost the report kernel more than that in instruction fetch; with
  // the leaner arithmetic of round 2 it wins there too: 5.57 against 5.63 ms per 1e8 persons.)
  if (FOLD_SINGLE && !life.has_onset) {
    person.scheduleAt(life.t_death, toDeath);
    person.resume_after_init();
  } else {
    if (life.has_onset) person.scheduleAt(life.t_onset, toLocalised);
    person.scheduleAt(life.t_death, toDeath);
    person.resume_after_init();
  }
  if (person.error_) *error = person.error_;
}

}  // namespace person_r

#ifdef __CUDACC__

#ifndef MSIM_SBLOCK
#define MSIM_SBLOCK 320
#endif
#ifndef MSIM_SMINB
#define MSIM_SMINB 2
#endif
constexpr int SBLOCK = MSIM_SBLOCK;  // threads per block of the staged kernel
constexpr int SWARPS = SBLOCK / 32;
constexpr int QCAP = 64;        // a queue holds < 32 entries between pushes, < 64 after one
constexpr int XROWS = 4;        // rows a person may log beyond its first (person_r::MAX_ROWS - 1)
constexpr int ROW_CHUNK = 1024;  // persons per block of the row-log passes
constexpr int ROW_OV_MIN = 64;    // smallest region for a chunk's further rows, whatever the plan says
#ifndef MSIM_ROW_WINDOW
#define MSIM_ROW_WINDOW 1536
#endif
#ifndef MSIM_ROW_MINB
#define MSIM_ROW_MINB 4
#endif
// rows a block of the final pass stages at a time (a chunk holds 1262 on average), and blocks per SM.  With one load
// per thread in flight at a time (round 2's first form) 1536 rows and 6 blocks per SM (40 registers) measured best:
// 5.06 ms per 1e8 persons for the whole row-log step against 5.29 (2048 rows, 55 registers), 5.14 (2048, 5 blocks),
// 5.10 (1280, 7), 5.13 (1024, 8).  With everything a thread needs requested up front (the kernel as it is now) the
// registers matter more than the sixth block: 4 blocks (64 registers) or 5 (48, spilling 36 bytes) 4.73 ms, 6 blocks
// (40 registers, spilling) 4.84  [gpurun_out/variants, round 2]
constexpr int ROW_WINDOW = MSIM_ROW_WINDOW;

struct alignas(16) WarpQueues {
  uint32_t q1_s[6][QCAP];
  double q1_u[QCAP], q1_t[QCAP];
  uint32_t q1_idx[QCAP];
  uint32_t q2_s[6][QCAP];
  double q2_on[QCAP], q2_t[QCAP];
  uint32_t q2_idx[QCAP];
};
// rows a lane logs beyond its person's first, until the warp appends them (row-log runs only)
struct alignas(16) WarpRows {
  double x_start[XROWS][32], x_end[XROWS][32];
  uint32_t x_meta[XROWS][32];
};
constexpr int FOLD_GROUP = 16;  // blocks whose report slices are added by the last of them to finish

constexpr int PEER_MAX = 8;  // GPUs of one node that merge their result vectors inside the launch
// Merge of a sharded run's result vector across the GPUs of one node (msim_peer_*; device side further down): every
// rank's mailbox as mapped into this process (NVLink peer memory).
struct PeerParams {
  unsigned long long *mail[PEER_MAX];  // [2][world][cap] pairs of words (payload 32 bits | epoch << 32)
  int rank, world, cap, total;         // total: reachable cells + counts (peer_cell)
  unsigned epoch;                      // of this launch: 1, 2, ...
  int mode;                            // 1: this launch's vectors are exchanged and added before it ends (-> dense);
                                       // 2: this launch adds the PREVIOUS launch's vectors (-> out) and sends its own
  double *out;                         // mode 2: the merged vector, in dense's layout
  const int *cell;                     // [total] entry of the vector that travels -> cell of dense (peer_cell_of, host)
};

struct StagedParams {
  StreamStart st;  // stride = M^(wave_blocks*SBLOCK)
  // Two WAVES of wave_blocks blocks each when the grid is two blocks per SM: the blocks that are the first on their
  // SMs take the tiles [0, wave_split), the others [wave_split, tiles); base_b = substream state of the person that
  // tile wave_split starts with.  One wave (wave_blocks = gridDim.x, wave_split = tiles) otherwise.  Why: see the kernel.
  int wave_blocks;
  int64_t wave_split;
  uint32_t base_b[6];
  unsigned long long *wave_clock;  // [3]: ~(first block's start), wave 0's end, wave 1's end (globaltimer ns; max); null: one wave
  unsigned *wave_sync;             // [2] blocks that joined each wave; zero before the launch
  int64_t first_person, n;
  unsigned outputs;
  person_r::Consts consts;
  ReportGeom geom;
  ReportCompact compact;  // the rows of the report this model can reach (block tables are laid out by these)
  double *dense;   // dense report accumulators in global memory (zeroed before launch)
  double *slices;  // gridDim.x slices of 2 * compact.n_states * n_bin doubles (pt, ut); every block zeroes its own
  unsigned *fold_sync;  // [groups + 2] arrival counters of the fold (+ one for the peer merge), zero before the first launch, left zero by every launch
  double *counts;  // N_COUNTS doubles
  double *indiv;   // [n] every individual's discounted total at individualReset (outputs & 8), NaN below startReportAge
  int *error;
  // row log, first pass
  double *first_end;           // [n] end time of each person's first row
  unsigned short *first_tag;   // [n] kind | state << 4 | (rows - 1) << 8
  // further rows: chunk c owns entries [c*ov_cap, (c+1)*ov_cap); tag = person-in-chunk << 16 | row << 12 | state << 8 | kind
  unsigned *ov_tag;
  double *ov_start, *ov_end;
  unsigned *ov_fill;           // [chunks] entries appended to each chunk's region (may exceed ov_cap: then the run is repeated)
  int ov_cap;
  PeerParams peer;  // merge of the dense vector across the GPUs of a node INSIDE this launch (below); world <= 1 = none
};

// Report accumulators of a block.  Counters: tables in shared memory laid out by the model's reachable states and
// (state, event) pairs (ReportCompact), native 32-bit atomics, flushed once per block.  FP sums
// (pt, ut): red.global.add.f64 into a slice of global memory that only THIS
// block writes — fire and forget, no return value to wait for.  Shared memory
// has no native f64 add on sm_100a (atomicAdd there is a compare-and-swap
// loop: 2.7 ms per 1e8 persons for the 19% of intervals that are not a
// person's first; per-warp private tables with plain load/add/store for the
// first intervals were also slower than the REDs).  The slices are added in a fixed order inside the same launch:
// the last block of every group of FOLD_GROUP to finish adds the group's slices, the last group to finish adds the
// groups (no second kernel; the order of addition does not depend on which block does it).  (All blocks
// accumulating into ONE global vector: 38 ms per 1e8 persons, a few hundred hot cells serialise in L2.)
// red.shared.add.s32: one instruction, no return value.  (Written in PTX: the compiler turns some atomicAdd calls
// on shared memory into a loop that groups the lanes by address first — up to 32 passes when the addresses differ,
// as they do here.)
__device__ __forceinline__ void red_shared_add(int *ptr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(ptr)), "r"(v) : "memory");
}

struct StagedAcc {
  double *f;                 // this block's global slice: pt[slot][bin], then ut[slot][bin]
  int *i;                    // shared: dstart[slot][bin], noedge[slot][bin], events[pair][bin]
  const ReportCompact *cm;   // in shared memory
  int nb, n_event, off_ut, off_noedge, off_events;
  int *err;
  // one-entry write-combining cache for the `dstart` counter: consecutive +1s
  // to the same cell (every person's first interval starts in the same cell)
  // cost nothing until the cell changes
  int pend_idx;
  int pend_n;
  __device__ __forceinline__ void addc(int idx) {
    if (idx == pend_idx) {
      pend_n++;
    } else {
      if (pend_n) red_shared_add(&i[pend_idx], pend_n);
      pend_idx = idx;
      pend_n = 1;
    }
  }
  __device__ __forceinline__ void flush() {
    if (pend_n) red_shared_add(&i[pend_idx], pend_n);
    pend_n = 0;
    pend_idx = -1;
  }
  // what report_add (report.cuh) adds, on this block's tables
  __device__ __forceinline__ void add(const ReportCells &c, bool disc) {
    const int ss = cm->state_slot[c.state], ps = cm->pair_slot[c.state * n_event + c.event];
    if (ss < 0 || ps < 0) {
      *err = MSIM_ERR_MODEL_DEV;
      return;
    }
    const int row = ss * nb;
    red_shared_add(&i[off_events + ps * nb + c.bin_hi], 1);
    addc(row + c.bin_f0);
    if (c.row_lo >= 0) {
      const int lo = row + (c.row_lo - c.state * nb);
      atomicAdd(&f[lo], c.pt_lo);
      if (disc) atomicAdd(&f[off_ut + lo], c.ut_lo);
    }
    atomicAdd(&f[row + c.bin_hi], c.pt_hi);
    if (disc) atomicAdd(&f[off_ut + row + c.bin_hi], c.ut_hi);
    if (c.noedge) red_shared_add(&i[off_noedge + row + c.bin_hi], 1);
  }
};

// OUT = the MSIM_OUT_* combination when it is known at compile time (each combination is its own kernel), or
// OUT_RUNTIME: read from the parameters (used where individual totals are kept).  Which is faster is a matter of what
// the compiler makes of it.  Round 1 (1e8 persons): without a report the specialised kernels won (counts 3.97 vs 4.30
// ms, row log 5.92 vs 6.15), with one the general kernel did (6.2 vs 9.6 ms — the specialised one ran the substream
// jump and the draws in fragments of a few lanes).  Round 2, after the arithmetic was rewritten (mrg32k3a.cuh,
// fastmath.cuh) and the block tables shrunk: the specialised report kernel wins too (5.35 vs 5.56 ms at 256 threads,
// 103 registers), and at 92 registers it fits 320-thread blocks — 20 warps per SM, 5 per scheduler — without spills:
// 4.77 ms.  Block sizes, all kernels specialised, counts / report + counts / row log in ms per 1e8 persons:
// 256 x 2: 3.25 / 5.31 / 5.03; 288 x 2: 3.45 / 5.34 / 5.20; 320 x 2: 3.11 / 4.79 / 4.84; 352 x 2: 3.26 / 5.25 / 5.21;
// 384 x 2 (80 registers, spills): 2.99 / 4.92 / 4.90.  (Warps per SM that are not a multiple of 4 leave the four
// schedulers unevenly loaded.)
constexpr unsigned OUT_RUNTIME = 0xffu;
// UNIT: instantiated for the built models' partition {0, 1, ..., 100, 1e6} (ReportGeom::unit; report_cells<true>)
template <unsigned OUT, bool UNIT = false>
struct StagedEnvT {
  typedef MrgSource Src;
  Src src;
  unsigned outputs_rt;
  __device__ __forceinline__ unsigned out() const { return OUT == OUT_RUNTIME ? outputs_rt : OUT; }
  const ReportGeom *geom;  // in shared memory
  StagedAcc acc;
  ReportCursor cursor;
  // per-kind tallies: eight 16-bit fields in two words (each kind occurs at most once per person).  Only kept when
  // there is no report: with one, the per-kind counts are the column sums of the block's event table.
  unsigned long long kinds_lo, kinds_hi;
  double sum_end;
  // row log of the person in hand
  int nrows;
  double first_end;
  unsigned first_tag;
  WarpRows *wq;
  int lane;
  double current;  // EventReport::current: the individual's discounted total so far (outputs & 8)
  __device__ __forceinline__ void begin_person() {
    report_cursor_start(*geom, cursor);
    nrows = 0;
    current = 0.0;
  }
  __device__ __forceinline__ void row(double, double start, double end, int state, int kind) {
    if (out() & 4u) {
      if (!(out() & 2u)) {
        const unsigned long long inc = 1ull << ((kind & 3) * 16);
        if (kind < 4) kinds_lo += inc;
        else kinds_hi += inc;
      }
      sum_end += end;
    }
    if (out() & 1u) {
      if (nrows == 0) {
        first_end = end;
        first_tag = (unsigned)kind | ((unsigned)state << 4);
      } else if (nrows <= XROWS) {
        wq->x_start[nrows - 1][lane] = start;
        wq->x_end[nrows - 1][lane] = end;
        wq->x_meta[nrows - 1][lane] = ((unsigned)nrows << 12) | ((unsigned)state << 8) | (unsigned)kind;
      }
      nrows++;
    }
  }
  __device__ __forceinline__ void report(int state, int kind, double lhs, double rhs) {
    if (out() & 2u) {
      // (a short form of the cells for a person's first interval — no left partial, no left exponential — was tried
      // here, twice: 5.99 against 5.56 ms per 1e8 persons with the run-time-flag kernel, 4.93 against 4.77 with this
      // one; the second copy of the cell code costs more in instruction fetch and scheduling than the arithmetic it skips)
      if (UNIT && !(cursor.valid && cursor.t == lhs)) {  // person-r reports from previous() to now(): every interval starts where the last one ended
        *acc.err = MSIM_ERR_MODEL_DEV;
        return;
      }
      const ReportCells c = report_cells<UNIT>(*geom, cursor, state, kind, lhs, rhs);
      acc.add(c, geom->discount_rate != 0.0);
      if (out() & 8u) current += report_brief(*geom, c, lhs, rhs);  // addBrief, microsimulation.h:929-932
    }
  }
};

// a finished person's first row and row count (rows beyond the first stay staged until the warp appends them)
template <class StagedEnv>
__device__ __forceinline__ void staged_finish(const StagedParams &p, StagedEnv &env, uint32_t idx, int *err) {
  if (env.out() & 1u) {
    if (env.nrows > XROWS + 1) *err = -4;
    const int extra = env.nrows > 0 ? (env.nrows > XROWS + 1 ? XROWS : env.nrows - 1) : 0;
    p.first_end[idx] = env.first_end;
    p.first_tag[idx] = (unsigned short)(env.first_tag | ((unsigned)extra << 8) | (env.nrows > 0 ? 0x8000u : 0u));
  }
  // EventReport::individualReset (microsimulation.h:905-916) at the end of the person's simulation: now() is the
  // time of its last event (the cursor's), the total counts only if that is past startReportAge (else NA_REAL)
  if ((env.out() & 10u) == 10u)
    p.indiv[idx] = (env.cursor.t >= env.geom->start_report_age) ? env.current : __longlong_as_double(0x7ff8000000000000ll);
}

// a lane appends its person's further rows to the region of the person's chunk
template <class StagedEnv>
__device__ __forceinline__ void staged_append_extra(const StagedParams &p, StagedEnv &env, bool active, uint32_t idx, int lane) {
  const int cnt = active ? (env.nrows > XROWS + 1 ? XROWS : (env.nrows > 1 ? env.nrows - 1 : 0)) : 0;
  if (cnt == 0) return;
  const unsigned chunk = idx / ROW_CHUNK, local = idx % ROW_CHUNK;
  const unsigned at = atomicAdd(&p.ov_fill[chunk], (unsigned)cnt);
  for (int r = 0; r < cnt; r++) {
    if (at + r < (unsigned)p.ov_cap) {
      const size_t e = (size_t)chunk * p.ov_cap + at + r;
      p.ov_tag[e] = (local << 16) | env.wq->x_meta[r][lane];
      p.ov_start[e] = env.wq->x_start[r][lane];
      p.ov_end[e] = env.wq->x_end[r][lane];
    }
  }
}

// shared-memory bytes of the block tables for a report laid out by `cm`
inline size_t staged_counter_bytes(const ReportGeom &g, const ReportCompact &cm) {
  return sizeof(int) * (size_t)(2 * cm.n_states + cm.n_pairs) * g.n_bin;
}
inline int staged_slice_doubles(const ReportGeom &g, const ReportCompact &cm) { return 2 * cm.n_states * g.n_bin; }

// ---- all-reduce of the dense vector over NVLink peer memory, by the block that finished the fold ------------------
// Every rank owns a mailbox (cudaMalloc + CUDA IPC, mapped into every other rank's process) with one slot per source
// rank and epoch parity.  SEND: the last block of rank r stores its vector into slot r of every rank's mailbox (its
// own included).  COLLECT: it reads all ranks' vectors of one epoch from its OWN mailbox and adds them in rank order —
// the same order on every rank, so all ranks hold bit-identical sums.  No flags and no fences: every 8-byte word that
// travels carries 32 bits of payload and the 32-bit epoch (a double = two such words, one 16-byte store), an 8-byte
// store arrives whole, and the reader polls each word until it shows the epoch — the latency of ONE store across
// NVLink instead of store, fence (a round trip), flag, fence.
//   mode 1: send(e), collect(e) -> dense.  The launch ends with the merged vector in place; it waits for the slowest
//           rank (as an all-reduce after the launch would).
//   mode 2: collect(e - 1) -> out, send(e).  A launch adds up what the ranks sent a whole launch ago, so nobody waits
//           and the ranks' launches need not end together; the result of launch e is complete after launch e + 1, or
//           after peer_collect_kernel (msim_peer_flush) at the end of a sequence.
// Either way a step of a sharded run is ONE launch (no NCCL kernel): what the strong-scaling run (1e8 persons over
// 8 GPUs, 0.65 ms per step) needs.  Slots alternate with the epoch's parity, and a rank sends epoch e only after it has
// collected e - 1 (mode 2) or e (mode 1) from every peer — which those peers sent after they were done with the slot's
// previous content — so a slot is never overwritten while its owner may still read it.  A peer that does not show up
// within PEER_TIMEOUT_NS raises MSIM_ERR_PEER instead of hanging the GPU.
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;
constexpr int ERR_PEER = -9;  // MSIM_ERR_PEER

__device__ __forceinline__ void peer_store_pair(unsigned long long *p, unsigned long long w0, unsigned long long w1) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void peer_load_pair(const unsigned long long *p, unsigned long long &w0, unsigned long long &w1) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// What travels is the vector WITHOUT the cells the model cannot reach (ReportCompact: person-r uses 4 of 8 states and
// 11 of 64 (state, event) pairs — 2 763 of 9 810 doubles): entry c of the compact vector is cell q.cell[c] of `dense`
// (a table the host fills with peer_cell_of: one block does the whole exchange, index arithmetic would be in its way).
inline int peer_cell_of(const ReportGeom &g, const ReportCompact &cm, int c) {
  const int nb = g.n_bin, n_sb = cm.n_states * nb;
  if (c < 4 * n_sb) {
    const int t = c / n_sb, r = c - t * n_sb, slot = r / nb, b = r - slot * nb;
    const int off = t == 0 ? g.off_pt : t == 1 ? g.off_ut : t == 2 ? g.off_dstart : g.off_noedge;
    return off + (int)cm.slot_state[slot] * nb + b;
  }
  c -= 4 * n_sb;
  if (c < cm.n_pairs * nb) {
    const int ps = c / nb, b = c - ps * nb;
    return g.off_events + ((int)cm.pair_state[ps] * g.n_event + (int)cm.pair_event[ps]) * nb + b;
  }
  return g.total + (c - cm.n_pairs * nb);  // the counts follow the report's cells
}

// this rank's vector (from `dense`) into slot `rank` of every mailbox, tagged with `epoch`
template <int THREADS>
__device__ __forceinline__ void peer_send(const PeerParams &q, const double *dense, unsigned epoch) {
  const size_t slot = ((size_t)(epoch & 1u) * q.world + q.rank) * (size_t)q.cap;
  const unsigned long long ep = (unsigned long long)epoch << 32;
  for (int c0 = threadIdx.x; c0 < q.total; c0 += 4 * THREADS) {  // four entries' values requested together, then sent
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = c0 + i * THREADS;
      v[i] = c < q.total ? __ldcg(&dense[__ldg(&q.cell[c])]) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int c = c0 + i * THREADS;
      if (c >= q.total) break;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[i]);
      const unsigned long long w0 = (bits & 0xffffffffull) | ep, w1 = (bits >> 32) | ep;
      for (int r = 0; r < q.world; r++) peer_store_pair(q.mail[r] + 2 * (slot + c), w0, w1);
    }
  }
}

// all ranks' vectors of `epoch`, from this rank's mailbox, added in rank order into dst (dense's layout)
template <int THREADS>
__device__ __forceinline__ void peer_collect(const PeerParams &q, double *dst, unsigned epoch, int *error) {
  const unsigned long long *box = q.mail[q.rank] + 2 * ((size_t)(epoch & 1u) * q.world) * (size_t)q.cap;
  const unsigned long long ep = (unsigned long long)epoch << 32;
  const unsigned long long t0 = peer_now_ns();
  bool late = false;
  for (int c = threadIdx.x; c < q.total; c += THREADS) {
    const int cell = __ldg(&q.cell[c]);
    // all ranks' words for this entry are requested before the first is looked at (one memory latency, not `world`)
    unsigned long long w0[PEER_MAX], w1[PEER_MAX];
#pragma unroll
    for (int r = 0; r < PEER_MAX; r++)
      if (r < q.world) peer_load_pair(box + 2 * ((size_t)r * q.cap + c), w0[r], w1[r]);
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < PEER_MAX; r++) {
      if (r >= q.world) break;
      unsigned spins = 0;
      while ((w0[r] & 0xffffffff00000000ull) != ep || (w1[r] & 0xffffffff00000000ull) != ep) {
        if (late || ((++spins & 0xfffu) == 0 && peer_now_ns() - t0 > PEER_TIMEOUT_NS)) {
          late = true;
          break;
        }
        peer_load_pair(box + 2 * ((size_t)r * q.cap + c), w0[r], w1[r]);
      }
      t += __longlong_as_double((long long)((w0[r] & 0xffffffffull) | (w1[r] << 32)));
    }
    dst[cell] = t;
  }
  if (late) atomicCAS(error, 0, ERR_PEER);
}

// by the block that finished the fold: dense is complete
__device__ __noinline__ void staged_peer_merge(const StagedParams &p) {
  __threadfence();
  __syncthreads();  // this block's own stores to `dense` (the fold) are visible to all its threads
  peer_send<SBLOCK>(p.peer, p.dense, p.peer.epoch);
  if (p.peer.mode == 1) {
    __syncthreads();  // dense is overwritten below: every thread has read what it sends
    peer_collect<SBLOCK>(p.peer, p.dense, p.peer.epoch, p.error);
  }
}

// mode 2, by the FIRST block of the launch to finish its persons: the previous launch's vectors, while the other blocks
// are still at work (it does not need this launch's results, so it is kept off the launch's critical path)
__device__ __noinline__ void staged_peer_collect_previous(const StagedParams &p) {
  peer_collect<SBLOCK>(p.peer, p.peer.out, p.peer.epoch - 1, p.error);
}

// the last launch's vectors, for a sequence of mode-2 launches (msim_peer_flush)
struct PeerCollectParams {
  PeerParams peer;
  int *error;
};
__global__ void __launch_bounds__(256) peer_collect_kernel(const __grid_constant__ PeerCollectParams p) {
  peer_collect<256>(p.peer, p.peer.out, p.peer.epoch, p.error);
}

// PEER: with the merge across GPUs (above) in the launch's tail.  A separate instantiation: compiled into the one kernel,
// the two calls cost runs that do not merge 0.4% (4.83 against 4.81 ms per 1e8 persons: another register allocation).
#ifdef MSIM_TRACE
// development aid (scripts/trace_run.py, build with -DMSIM_TRACE): where a launch's time goes.  Every block leaves time
// stamps in global memory; the block that ends the fold prints, per stamp, the earliest and latest block (ns after the
// first block's entry) — one printf at the very end, so the stamps themselves are not disturbed.
__device__ unsigned long long g_trace[1024][4];
__device__ unsigned long long g_trace_warp[1024][32];  // when each warp of the block had no persons left
__device__ unsigned g_trace_sm[1024];
#define MSIM_STAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_trace[blockIdx.x][k] = peer_now_ns(); } while (0)
#else
#define MSIM_STAMP(k) do { } while (0)
#endif

template <unsigned OUT, bool PEER = false, bool UNIT = false>
__global__ void __launch_bounds__(SBLOCK, MSIM_SMINB) person_staged_kernel(const __grid_constant__ StagedParams p) {
  typedef StagedEnvT<OUT, UNIT> StagedEnv;
  const unsigned outs = (OUT == OUT_RUNTIME) ? p.outputs : OUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // shared layout: [per-warp queues | per-warp row staging (row log) | report counters]
  WarpQueues *queues = reinterpret_cast<WarpQueues *>(smem_raw);
  unsigned char *sp = smem_raw + sizeof(WarpQueues) * SWARPS;
  WarpRows *rows_stage = reinterpret_cast<WarpRows *>(sp);
  if (outs & 1u) sp += sizeof(WarpRows) * SWARPS;
  const bool rep = (outs & 2u) != 0;
  const int nb = p.geom.n_bin;
  const int n_sb = rep ? p.compact.n_states * nb : 0;          // cells of one [slot][bin] table
  const int n_i = rep ? 2 * n_sb + p.compact.n_pairs * nb : 0;  // counters: dstart, noedge, events
  const int n_f = 2 * n_sb;                                     // FP sums: pt, ut
  const int n_sb_d = n_sb > 0 ? n_sb : 1, nb_d = nb > 0 ? nb : 1;  // divisors (the tables are empty without a report)
  int *s_i = reinterpret_cast<int *>(sp);
  double *g_f = p.slices + (size_t)blockIdx.x * n_f;  // this block's slice of the FP sums
  __shared__ unsigned s_kind[8];
  __shared__ double s_sum[SWARPS];
  // the report geometry in shared memory: its tables are indexed per lane (a constant-bank load with a
  // different index in every lane is serialised)
  __shared__ ReportGeom s_geom;
  __shared__ ReportCompact s_cm;
  __shared__ int s_error;
  __shared__ int s_last;
  __shared__ long long s_tile_end;
  __shared__ int s_wave[2];  // this block's wave and its place in it
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  MSIM_STAMP(0);
  const unsigned lt_mask = (1u << lane) - 1u;
  WarpQueues &Q = queues[warp];
  if (tid < 8) s_kind[tid] = 0;
  if (tid == 0) s_error = 0;
  if (tid == 0) {
    // Which wave, and which of its blocks, this block is.  The favoured block of an SM is the one in the SM's lower
    // warp slots (%warpid of its first warp 0, 1 or 3 against 10, 11 or 9 on every SM of a B200: scripts/trace_run.py);
    // the block index does not tell (the block scheduler hands 6 of the first 148 blocks to SMs that have one), nor
    // does the order in which the two arrive at a counter (they start within 0.4 us of one another).  A wave that is
    // full — a GPU shared with another kernel, so that an SM sees its blocks one after the other — sends the block to
    // the other one: every place in every wave is taken exactly once whatever the slots are.
    int wave = 0;
    unsigned rank = blockIdx.x;
    if (p.wave_clock) {
      unsigned wid;
      asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
      wave = wid >= (unsigned)(SWARPS * 3 / 5) ? 1 : 0;
      rank = atomicAdd(&p.wave_sync[wave], 1u);
      if (rank >= (unsigned)p.wave_blocks) {
        wave ^= 1;
        rank = atomicAdd(&p.wave_sync[wave], 1u);
      }
    }
    s_wave[0] = wave;
    s_wave[1] = (int)rank;
    s_tile_end = wave ? (long long)((p.n + 31) / 32) : (long long)p.wave_split;
  }
  for (int k = tid; k < (int)(sizeof(ReportGeom) / 4); k += SBLOCK)
    reinterpret_cast<int *>(&s_geom)[k] = reinterpret_cast<const int *>(&p.geom)[k];
  for (int k = tid; k < (int)(sizeof(ReportCompact) / 4); k += SBLOCK)
    reinterpret_cast<int *>(&s_cm)[k] = reinterpret_cast<const int *>(&p.compact)[k];
  for (int k = tid; k < n_i; k += SBLOCK) s_i[k] = 0;
  for (int k = tid; k < n_f; k += SBLOCK) g_f[k] = 0.0;
  __syncthreads();

  StagedEnv env;
  int err = 0;
  env.outputs_rt = p.outputs;
  env.geom = &s_geom;
  env.acc.f = g_f;
  env.acc.i = s_i;
  env.acc.cm = &s_cm;
  env.acc.nb = nb;
  env.acc.n_event = p.geom.n_event;
  env.acc.off_ut = n_sb;
  env.acc.off_noedge = n_sb;
  env.acc.off_events = 2 * n_sb;
  env.acc.err = &err;
  env.acc.pend_idx = -1;
  env.acc.pend_n = 0;
  env.cursor.valid = false;
  env.kinds_lo = env.kinds_hi = 0;
  env.sum_end = 0.0;
  env.nrows = 0;
  env.first_end = 0.0;
  env.first_tag = 0;
  env.wq = &rows_stage[warp];
  env.lane = lane;
  const person_r::Consts &K = p.consts;

  // The two blocks of an SM do not advance alike: the schedulers favour the warps of the block that arrived first, and
  // with equal shares that block was done at 3.97 ms of a 4.5 ms launch while its neighbour ran the last 0.45 ms with
  // half the SM's warps (1e8 persons, scripts/trace_run.py).  So the first wave of blocks gets the larger share of the
  // tiles (wave_split; the host sets it from the waves' finishing times of earlier launches, wave_clock), each wave
  // dealing its range of tiles to its warps in turn.  (Handing tiles out by ticket from a counter per SM instead evens
  // the blocks out as well, but costs more than it gains: profiles/experiments/r02_tiles_by_ticket.md.)
  const bool wave_b = s_wave[0] != 0;
  const unsigned w_thread = (unsigned)s_wave[1] * SBLOCK + (unsigned)tid;
  Mrg32k3a sub;
  {
    StreamStart st = p.st;
    if (wave_b)
      for (int k = 0; k < 6; k++) st.base[k] = p.base_b[k];
    sub = thread_start_state(st, w_thread);
  }
  // (the end of the wave's range of tiles is read from shared memory in every pass of the loop: held in registers it
  // takes the kernel to its register limit, with spills)
  const int64_t tile_first = (wave_b ? p.wave_split : 0) + (w_thread >> 5), GW = (int64_t)p.wave_blocks * SWARPS;
  if (tid == 0 && p.wave_clock) atomicMax(&p.wave_clock[0], ~peer_now_ns());
  int c1 = 0, c2 = 0;  // queue fills (warp-uniform)

  // Each pass of the loop runs ONE stage on up to 32 persons and then, in one place for all stages, the events
  // of the persons that became ready in it: new persons while there are tiles (stage A), parked persons
  // whenever a full warp of them waits (B1: onset times; B2: lives with onset before death), and whatever is
  // left once the tiles are used up.
  MSIM_STAMP(1);
  for (int64_t tile = tile_first;;) {
    const bool last = tile >= *(volatile long long *)&s_tile_end;
    bool ready = false, park1 = false, park2 = false;
    person_r::Life life;
    person_r::ParkedOnset po;
    uint32_t idx = 0;
    if (c2 >= 32 || (last && c1 == 0 && c2 > 0)) {
      // ---- stage B2: the top <= 32 entries of queue 2
      const int m = c2 > 32 ? 32 : c2;
      c2 -= m;
      if (lane < m) {
        const int e = c2 + lane;
#pragma unroll
        for (int k = 0; k < 6; k++) life.g.s[k] = Q.q2_s[k][e];
        life.t_onset = Q.q2_on[e];
        life.t_death = Q.q2_t[e];
        life.has_onset = true;
        idx = Q.q2_idx[e];
        ready = true;
      }
    } else if (c1 >= 32 || (last && c1 > 0)) {
      // ---- stage B1: the top <= 32 entries of queue 1; those with onset before death go to queue 2
      const int m = c1 > 32 ? 32 : c1;
      c1 -= m;
      if (lane < m) {
        const int e = c1 + lane;
#pragma unroll
        for (int k = 0; k < 6; k++) po.g.s[k] = Q.q1_s[k][e];
        po.u_onset = Q.q1_u[e];
        po.t_death = Q.q1_t[e];
        idx = Q.q1_idx[e];
        ready = person_r::stage_b1<StagedEnv>(K, po, &life);
        park2 = !ready;
      }
      __syncwarp();
    } else if (!last) {
      // ---- stage A: 32 new persons
      const int64_t i64 = tile * 32 + lane;
      idx = (uint32_t)i64;
      // the jump to the lane's NEXT person first, while the warp is together (placed after the draws, whose loops
      // diverge, its 18 modular multiplications ran in fragments of 6-8 lanes)
      env.src.g = sub;
      sub.jump(p.st.stride);
      if (i64 < p.n) {
        ready = person_r::stage_a(env, &po, &life);
        park1 = !ready;
      }
      tile += GW;
    } else {
      break;
    }
    // ---- parked persons join their queue (push by lane order)
    const unsigned m1 = __ballot_sync(0xffffffffu, park1), m2 = __ballot_sync(0xffffffffu, park2);
    if (park1) {
      const int e = c1 + __popc(m1 & lt_mask);
#pragma unroll
      for (int k = 0; k < 6; k++) Q.q1_s[k][e] = po.g.s[k];
      Q.q1_u[e] = po.u_onset;
      Q.q1_t[e] = po.t_death;
      Q.q1_idx[e] = idx;
    }
    if (park2) {
      const int e = c2 + __popc(m2 & lt_mask);
#pragma unroll
      for (int k = 0; k < 6; k++) Q.q2_s[k][e] = life.g.s[k];
      Q.q2_on[e] = life.t_onset;
      Q.q2_t[e] = life.t_death;
      Q.q2_idx[e] = idx;
    }
    c1 += __popc(m1);
    c2 += __popc(m2);
    // ---- the persons that are ready run their events
    if (ready) {
      person_r::run_life<true>(env, K, p.first_person + idx, life, &err);
      staged_finish(p, env, idx, &err);
    }
    __syncwarp();
    if (outs & 1u) staged_append_extra(p, env, ready, idx, lane);
    __syncwarp();
  }

  MSIM_STAMP(2);
#ifdef MSIM_TRACE
  if (lane == 0 && blockIdx.x < 1024) {
    g_trace_warp[blockIdx.x][warp] = peer_now_ns();
    if (warp == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned wid;
      asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
      g_trace_sm[blockIdx.x] = smid | ((unsigned)s_wave[0] << 16) | (wid << 20);
    }
  }
#endif
  env.acc.flush();
  if (err) atomicCAS(&s_error, 0, err);
  if (outs & 4u) {
    if (!rep) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        unsigned v = (unsigned)(((k < 4 ? env.kinds_lo : env.kinds_hi) >> ((k & 3) * 16)) & 0xffffull);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&s_kind[k], v);
      }
    }
    double v = env.sum_end;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_sum[warp] = v;
  }
  __syncthreads();
  if (tid == 0 && p.wave_clock) atomicMax(&p.wave_clock[1 + s_wave[0]], peer_now_ns());  // the block's persons are done
  // the block's counters -> the dense vector (cell by cell through the compact layout); with a report the per-kind
  // event counts are the column sums of the event table
  for (int k = tid; k < n_i; k += SBLOCK) {
    const int v = s_i[k];
    if (v == 0) continue;
    int cell;
    if (k < 2 * n_sb) {
      const int t = k / n_sb_d, r = k - t * n_sb, slot = r / nb_d, b = r - slot * nb;
      cell = (t == 0 ? p.geom.off_dstart : p.geom.off_noedge) + (int)s_cm.slot_state[slot] * nb + b;
    } else {
      const int r = k - 2 * n_sb, ps = r / nb_d, b = r - ps * nb;
      const int ev = s_cm.pair_event[ps];
      cell = p.geom.off_events + ((int)s_cm.pair_state[ps] * p.geom.n_event + ev) * nb + b;
      if (outs & 4u) atomicAdd(&s_kind[ev & 7], (unsigned)v);
    }
    atomicAdd(&p.dense[cell], (double)v);
  }
  if (outs & 4u) {
    __syncthreads();
    if (tid < 8 && s_kind[tid]) atomicAdd(&p.counts[tid], (double)s_kind[tid]);
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < SWARPS; w++) t += s_sum[w];
      atomicAdd(&p.counts[8], t);
    }
  }
  if (tid == 0 && s_error) atomicCAS(p.error, 0, s_error);
  MSIM_STAMP(3);
  if (!rep) return;

  // ---- the FP sums: slices -> dense, inside this launch.  The last block of a group to arrive adds the group's
  // slices (into the group's first slice), the last group to arrive adds the groups, both in index order, so the
  // result does not depend on who does the adding.
  const int n_groups = (gridDim.x + FOLD_GROUP - 1) / FOLD_GROUP;
  const int group = blockIdx.x / FOLD_GROUP, g0 = group * FOLD_GROUP;
  const int g_size = min(FOLD_GROUP, (int)gridDim.x - g0);
  if (PEER && p.peer.world > 1 && p.peer.mode == 2 && p.peer.epoch > 1) {
    if (tid == 0) s_last = (atomicAdd(&p.fold_sync[n_groups + 1], 1u) == 0u);  // the first to get here
    __syncthreads();
    if (s_last) staged_peer_collect_previous(p);
    __syncthreads();
  }
  __threadfence();  // this thread's REDs are performed before the arrival below is visible
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&p.fold_sync[group], 1u) == (unsigned)(g_size - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int k = tid; k < n_f; k += SBLOCK) {
    double t = 0.0;
    for (int b = 0; b < g_size; b++) t += __ldcg(&p.slices[(size_t)(g0 + b) * n_f + k]);
    __stcg(&p.slices[(size_t)g0 * n_f + k], t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    p.fold_sync[group] = 0;  // left as found, for the next launch
    s_last = (atomicAdd(&p.fold_sync[n_groups], 1u) == (unsigned)(n_groups - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int k = tid; k < n_f; k += SBLOCK) {
    double t = 0.0;
    for (int gi = 0; gi < n_groups; gi++) t += __ldcg(&p.slices[(size_t)gi * FOLD_GROUP * n_f + k]);
    const int tb = k / n_sb_d, r = k - tb * n_sb, slot = r / nb_d, b = r - slot * nb;
    p.dense[(tb == 0 ? p.geom.off_pt : p.geom.off_ut) + (int)s_cm.slot_state[slot] * nb + b] = t;
  }
  if (tid == 0) p.fold_sync[n_groups] = p.fold_sync[n_groups + 1] = 0;
#ifdef MSIM_TRACE
  if (tid == 0) {
    const unsigned long long now = peer_now_ns();
    const int nblk = min((int)gridDim.x, 1024);
    unsigned long long lo[4], hi[4];
    for (int k = 0; k < 4; k++) lo[k] = ~0ull, hi[k] = 0;
    for (int b = 0; b < nblk; b++)
      for (int k = 0; k < 4; k++) {
        const unsigned long long t = *(volatile unsigned long long *)&g_trace[b][k];
        lo[k] = t < lo[k] ? t : lo[k];
        hi[k] = t > hi[k] ? t : hi[k];
      }
    const unsigned long long t0 = lo[0];
    printf("trace n %lld entry %llu..%llu loop %llu..%llu drained %llu..%llu flushed %llu..%llu folded %llu (ns)\n", (long long)p.n,
           lo[0] - t0, hi[0] - t0, lo[1] - t0, hi[1] - t0, lo[2] - t0, hi[2] - t0, lo[3] - t0, hi[3] - t0, now - t0);
    if (p.n == 100000000)  // the full table once: block, SM, start of the loop, every warp's end, counters flushed (us)
      for (int b = 0; b < nblk; b++) {
        printf("tb %d sm %u wave %u loop %.1f warps", b, *(volatile unsigned *)&g_trace_sm[b] & 0xffffu, (*(volatile unsigned *)&g_trace_sm[b] >> 16 & 1u) + 100 * (*(volatile unsigned *)&g_trace_sm[b] >> 20),
               (double)(g_trace[b][1] - t0) * 1e-3);
        for (int w = 0; w < SWARPS; w++) printf(" %.0f", (double)(*(volatile unsigned long long *)&g_trace_warp[b][w] - t0) * 1e-3);
        printf(" flushed %.0f\n", (double)(g_trace[b][3] - t0) * 1e-3);
      }
  }
#endif
  if (PEER && p.peer.world > 1) staged_peer_merge(p);
}

// ---- row log passes -----------------------------------------------------------

struct RowPassParams {
  const double *first_end;
  const unsigned short *first_tag;
  const unsigned *ov_tag;
  const double *ov_start, *ov_end;
  const unsigned *ov_fill;
  int ov_cap;
  int64_t n, first_person;
  unsigned long long *chunk_rows;  // per-chunk totals, then (in place) exclusive offsets
  unsigned long long *n_rows;
  unsigned long long *max_fill;    // largest ov_fill seen (tells the host whether ov_cap was enough)
  int64_t n_chunks;
  double *rowlog;  // 5 columns (endTime, event, id, startTime, state), column stride `capacity`
  int64_t capacity;
};

// an integer below 2^52 as a double, exactly, without the conversion instruction (three per row otherwise, on a pipe
// that is a few lanes wide): 2^52 + v has v as its mantissa bits
__device__ __forceinline__ double row_u52_to_double(unsigned long long v) {
  return __longlong_as_double((long long)(0x4330000000000000ull | v)) - 4503599627370496.0;
}

__device__ __forceinline__ int tag_rows(unsigned short t) { return (t & 0x8000u) ? 1 + ((t >> 8) & 7) : 0; }

// rows per chunk of ROW_CHUNK persons; largest overflow fill
__global__ void __launch_bounds__(256) rowpass_count_kernel(RowPassParams p) {
  __shared__ unsigned s_w[8];
  const int64_t lo = (int64_t)blockIdx.x * ROW_CHUNK;
  unsigned v = 0;
  for (int k = threadIdx.x; k < ROW_CHUNK; k += 256) {
    int64_t i = lo + k;
    if (i < p.n) v += tag_rows(p.first_tag[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; w++) t += s_w[w];
    p.chunk_rows[blockIdx.x] = t;
    atomicMax(p.max_fill, (unsigned long long)p.ov_fill[blockIdx.x]);
  }
}

// one block: exclusive scan of the chunk totals in place; total row count out.  Eight consecutive chunks per thread
// and pass (8192 per pass: 12 passes for 1e8 persons instead of 96 — the pass count is what this one block's time is made of)
__global__ void __launch_bounds__(1024) rowpass_scan_kernel(RowPassParams p) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_carry;
  constexpr int PER = 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t c0 = 0; c0 < p.n_chunks; c0 += 1024 * PER) {
    const int64_t c = c0 + (int64_t)tid * PER;
    unsigned long long x[PER], v = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      x[k] = (c + k < p.n_chunks) ? p.chunk_rows[c + k] : 0;
      v += x[k];
    }
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    const unsigned long long carry = s_carry;
    unsigned long long run = carry + woff + incl - v;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      if (c + k < p.n_chunks) p.chunk_rows[c + k] = run;
      run += x[k];
    }
    __syncthreads();
    if (tid == 1023) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (tid == 0) *p.n_rows = s_carry;
}

// One block per chunk: every person's row offset by a block scan of the row counts, first rows and further
// rows placed in shared memory in final order, then streamed out (all global accesses coalesced).  The pass is
// bandwidth-bound and its warps wait for loads, so everything a thread needs from global memory is requested up front:
// its four persons' tags and first end times before the scan, its further row (tag, start, end) before the placing.
__global__ void __launch_bounds__(256, MSIM_ROW_MINB) rowpass_final_kernel(RowPassParams p) {
  __shared__ unsigned short s_off[ROW_CHUNK];
  __shared__ double s_end[ROW_WINDOW], s_start[ROW_WINDOW];
  __shared__ unsigned s_meta[ROW_WINDOW];
  __shared__ unsigned s_w[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t c = blockIdx.x, lo = c * ROW_CHUNK;
  constexpr int PER = ROW_CHUNK / 256;
  static_assert(PER == 4, "a thread owns four consecutive persons: one 8-byte load of tags, two 16-byte loads of end times");
  // ---- thread t owns persons [PER*t, PER*t + PER): tags, first end times, and (below) the scan of their row counts
  const int64_t i0 = lo + (int64_t)tid * PER;
  unsigned short tag[PER];
  double fe[PER];
  if (i0 + PER <= p.n) {
    const uint2 t2 = *reinterpret_cast<const uint2 *>(p.first_tag + i0);  // i0 is a multiple of 4: 8-byte aligned
    tag[0] = (unsigned short)(t2.x & 0xffffu);
    tag[1] = (unsigned short)(t2.x >> 16);
    tag[2] = (unsigned short)(t2.y & 0xffffu);
    tag[3] = (unsigned short)(t2.y >> 16);
    const double2 a = *reinterpret_cast<const double2 *>(p.first_end + i0), b = *reinterpret_cast<const double2 *>(p.first_end + i0 + 2);
    fe[0] = a.x;
    fe[1] = a.y;
    fe[2] = b.x;
    fe[3] = b.y;
  } else {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      tag[k] = (i0 + k < p.n) ? p.first_tag[i0 + k] : (unsigned short)0;
      fe[k] = (i0 + k < p.n) ? p.first_end[i0 + k] : 0.0;
    }
  }
  // this thread's further rows (entries of the chunk's region: as many as the chunk appended, at most ov_cap)
  unsigned m = p.ov_fill[c];
  if (m > (unsigned)p.ov_cap) m = (unsigned)p.ov_cap;
  const size_t ov0 = (size_t)c * p.ov_cap;
  const unsigned long long base = p.chunk_rows[c];
  constexpr int OV_PER = 2;  // entries per thread requested up front (512 per block; a region holds 522 at the default plan)
  unsigned otag[OV_PER];
  double oend[OV_PER], ostart[OV_PER];
#pragma unroll
  for (int k = 0; k < OV_PER; k++) {
    const unsigned e = tid + 256u * k;
    otag[k] = 0;
    oend[k] = ostart[k] = 0.0;
    if (e < m) {
      otag[k] = p.ov_tag[ov0 + e];
      oend[k] = p.ov_end[ov0 + e];
      ostart[k] = p.ov_start[ov0 + e];
    }
  }
  unsigned cnt[PER], v = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) {
    cnt[k] = tag_rows(tag[k]);
    v += cnt[k];
  }
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned run = incl - v, total = 0;
  for (int w = 0; w < 8; w++) {
    if (w < warp) run += s_w[w];
    total += s_w[w];
  }
  unsigned off[PER];
#pragma unroll
  for (int k = 0; k < PER; k++) {
    off[k] = run;
    s_off[tid * PER + k] = (unsigned short)run;
    run += cnt[k];
  }
  __syncthreads();
  for (unsigned w0 = 0; w0 < total; w0 += ROW_WINDOW) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const unsigned o = off[k] - w0;
      if ((tag[k] & 0x8000u) && o < (unsigned)ROW_WINDOW) {  // (unsigned wrap-around rejects rows before the window)
        s_end[o] = fe[k];
        s_start[o] = 0.0;  // a person's first interval starts at startTime = 0 (microsimulation.h:430)
        s_meta[o] = ((unsigned)(tid * PER + k) << 16) | (((unsigned)tag[k] >> 4) & 0xfu) << 8 | ((unsigned)tag[k] & 0xfu);
      }
    }
#pragma unroll
    for (int k = 0; k < OV_PER; k++) {
      if (tid + 256u * k < m) {
        const unsigned o = (unsigned)s_off[otag[k] >> 16] + ((otag[k] >> 12) & 0xfu) - w0;
        if (o < (unsigned)ROW_WINDOW) {
          s_end[o] = oend[k];
          s_start[o] = ostart[k];
          s_meta[o] = (otag[k] & 0xffff0000u) | (otag[k] & 0xfffu);
        }
      }
    }
    for (unsigned e = tid + 256u * OV_PER; e < m; e += 256) {  // regions planned larger than 512 entries
      const unsigned t = p.ov_tag[ov0 + e];
      const unsigned o = (unsigned)s_off[t >> 16] + ((t >> 12) & 0xfu) - w0;
      if (o < (unsigned)ROW_WINDOW) {
        s_end[o] = p.ov_end[ov0 + e];
        s_start[o] = p.ov_start[ov0 + e];
        s_meta[o] = (t & 0xffff0000u) | (t & 0xfffu);
      }
    }
    __syncthreads();
    const unsigned nw = (total - w0 < (unsigned)ROW_WINDOW) ? total - w0 : (unsigned)ROW_WINDOW;
    for (unsigned r = tid; r < nw; r += 256) {
      const unsigned long long g = base + w0 + r;
      if (g < (unsigned long long)p.capacity) {
        const unsigned mt = s_meta[r];
        p.rowlog[0 * p.capacity + g] = s_end[r];
        p.rowlog[1 * p.capacity + g] = row_u52_to_double(mt & 0xffu);
        p.rowlog[2 * p.capacity + g] = row_u52_to_double((unsigned long long)(p.first_person + lo + (mt >> 16)));
        p.rowlog[3 * p.capacity + g] = s_start[r];
        p.rowlog[4 * p.capacity + g] = row_u52_to_double((mt >> 8) & 0xfu);
      }
    }
    __syncthreads();
  }
}

// Tried and dropped (round 2): the three passes in ONE launch — chunks handed out by ticket, each chunk's row offset
// from a chained scan with decoupled look-back inside the kernel.  It saves reading the tags twice and the one-block
// scan between two launches, but measured 5.21 ms per 1e8 persons for the whole row-log step against 5.06 for the
// three launches: the look-back's L2 round trips sit in front of every block's bandwidth-bound part, and count + scan
// together only cost 0.35 ms.

#endif  // __CUDACC__

}  // namespace msim
