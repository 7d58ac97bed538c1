// TEST INFRASTRUCTURE — host emulation of the device engine.
//
// Compiles the SAME per-person templates the CUDA kernels instantiate
// (microsimulation_b200/csrc/{mrg32k3a,rmath,engine,models,report,seqstream}.cuh)
// with a plain C++ compiler and drives them with the kernels' work
// decomposition (grid x BLOCK threads, tile loop, stride jump, per-position
// draw counts + chunked chain resolution, three-wave MT regeneration), one
// "thread" at a time.  CPU-only tests compare this against the oracle so that
// heap, RNG, report and stream-offset logic is checked where there is no GPU.
// It is NOT part of the product: libmsim_b200.so contains none of it.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/msim.h"
#include "../../microsimulation_b200/csrc/host_state.h"
#include "../../microsimulation_b200/csrc/fastmath.cuh"
#include "../../microsimulation_b200/csrc/models.cuh"
#include "../../microsimulation_b200/csrc/calib_shared.cuh"
#include "../../microsimulation_b200/csrc/person_staged.cuh"
#include "../../microsimulation_b200/csrc/report.cuh"
#include "../../microsimulation_b200/csrc/report_host.h"
#include "../../microsimulation_b200/csrc/gsm.cuh"
#include "../../microsimulation_b200/csrc/lookups.cuh"
#include "../../microsimulation_b200/csrc/seqstream.cuh"
#include "../../microsimulation_b200/csrc/summary.cuh"

using namespace msim;

namespace {

constexpr int BLOCK = 256;
constexpr unsigned OUT_ROWLOG = 1u, OUT_REPORT = 2u, OUT_COUNTS = 4u;

struct PlainAcc {
  double *d;
  void addf(int idx, double v) { d[idx] += v; }
  void addi(int idx, int v) { d[idx] += v; }
  void addc(int idx) { d[idx] += 1; }
};

struct Rows {
  std::vector<double> end, event, id, start, state;
};

template <class SrcT>
struct HostEnv {
  typedef SrcT Src;
  Src src;
  unsigned outputs;
  Rows *rows;
  const ReportGeom *geom;
  PlainAcc acc;
  double *counts;
  ReportCursor cursor;
  void begin_person() { report_cursor_start(*geom, cursor); }
  void row(double id, double start, double end, int state, int kind) {
    if (outputs & OUT_COUNTS) {
      counts[kind] += 1;
      counts[8] += end;
    }
    if (outputs & OUT_ROWLOG) {
      rows->end.push_back(end);
      rows->event.push_back(kind);
      rows->id.push_back(id);
      rows->start.push_back(start);
      rows->state.push_back(state);
    }
  }
  void report(int state, int kind, double lhs, double rhs) {
    if (outputs & OUT_REPORT) report_add(*geom, acc, cursor, state, kind, lhs, rhs);
  }
};

template <class SrcT>
struct HostNullEnv {
  typedef SrcT Src;
  Src src;
  int code = 0;
  void begin_person() {}
  void summary(int, int, double, double) {}
  void set_uc(int k, double, double) { code = k; }
  void row(double, double, double, int, int) {}
  void report(int, int, double, double) {}
};

struct Tables {
  std::vector<MrgJump> lo, hi;
  Tables() : lo(JUMP_TAB), hi(JUMP_TAB) {
    const MrgJump &M = mrg_substream_jump();
    lo[0] = mrg_jump_identity();
    for (int j = 1; j < JUMP_TAB; j++) lo[j] = mrg_jump_mul(M, lo[j - 1]);
    MrgJump Mt = mrg_jump_mul(M, lo[JUMP_TAB - 1]);
    hi[0] = mrg_jump_identity();
    for (int j = 1; j < JUMP_TAB; j++) hi[j] = mrg_jump_mul(Mt, hi[j - 1]);
  }
};
const Tables &tables() {
  static Tables T;
  return T;
}

void make_start(const double seed[6], int64_t first_person_plus, int grid, StreamStart *st, int block = BLOCK) {
  uint32_t s[6];
  seed_to_u32(seed, s);
  Mrg32k3a b;
  memcpy(b.s, s, sizeof s);
  b.jump(mrg_jump_pow(mrg_substream_jump(), (uint64_t)first_person_plus));
  memcpy(st->base, b.s, sizeof b.s);
  st->lo = tables().lo.data();
  st->hi = tables().hi.data();
  st->stride = mrg_jump_pow(mrg_substream_jump(), (uint64_t)grid * block);
}

void rows_to_table(const Rows &r, const char *const *cols, msim_table *t) {
  int64_t n = (int64_t)r.id.size();
  t->nrow = n;
  t->ncol = 5;
  t->colnames = cols;
  t->data = (double *)malloc(sizeof(double) * 5 * (n ? n : 1));
  const std::vector<double> *src[5] = {&r.end, &r.event, &r.id, &r.start, &r.state};
  for (int j = 0; j < 5; j++)
    if (n) memcpy(t->data + j * n, src[j]->data(), sizeof(double) * n);
}

// rows come out in "thread-major" order; the kernel's scan/look-back restores
// person order, which the emulation does by sorting on (id, emission order)
void sort_rows_by_person(Rows &r) {
  size_t n = r.id.size();
  std::vector<size_t> idx(n);
  for (size_t i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return r.id[a] < r.id[b]; });
  Rows o;
  for (size_t i : idx) {
    o.end.push_back(r.end[i]);
    o.event.push_back(r.event[i]);
    o.id.push_back(r.id[i]);
    o.start.push_back(r.start[i]);
    o.state.push_back(r.state[i]);
  }
  r = o;
}

// R's MT state -> raw words of the current batch and n_batches-1 regenerations
std::vector<uint32_t> mt_words_emulated(const uint32_t state[625], int64_t n_batches) {
  std::vector<uint32_t> w((size_t)n_batches * 624);
  uint32_t mt[624];
  memcpy(mt, state + 1, sizeof mt);
  memcpy(w.data(), mt, sizeof mt);
  for (int64_t b = 1; b < n_batches; b++) {
    for (int wave = 0; wave < 3; wave++) {
      int kk[256];
      uint32_t v[256];
      bool act[256];
      for (int t = 0; t < 256; t++) act[t] = mt_wave_value(mt, wave, t, &kk[t], &v[t]);
      for (int t = 0; t < 256; t++)
        if (act[t]) mt[kk[t]] = v[t];
    }
    memcpy(w.data() + (size_t)b * 624, mt, sizeof mt);
  }
  return w;
}

template <template <class> class Model, class Consts>
int sequential(int64_t n, const Consts &consts, int max_draws, unsigned outputs, const ReportGeom &geom,
               uint32_t state[625], Rows *rows, std::vector<double> *dense) {
  const int64_t mti0 = state[0];
  const int64_t n_pos = (int64_t)max_draws * n;
  const int64_t words_needed = mti0 + n_pos + max_draws + 1;
  const int64_t n_batches = (words_needed + 623) / 624;
  const int64_t mt_end = n_batches * 624;
  std::vector<uint32_t> w = mt_words_emulated(state, n_batches);
  if (n == 0) return 0;
  // draws per start position
  std::vector<unsigned char> len((size_t)n_pos);
  for (int64_t p = 0; p < n_pos; p++) {
    typedef HostNullEnv<StreamSource> Env;
    Env env;
    env.src.w = w.data();
    env.src.pos = mti0 + p;
    env.src.end = mt_end;
    env.src.overrun = 0;
    Model<Env> person(env, consts, 0);
    person.run();
    int64_t used = env.src.pos - (mti0 + p);
    len[p] = (env.src.overrun || used > 250 || person.error_) ? 255 : (unsigned char)used;
  }
  // chunked chain resolution
  const int64_t n_chunks = (n_pos + CHAIN_CHUNK - 1) / CHAIN_CHUNK;
  std::vector<unsigned char> exit_off((size_t)n_chunks * CHAIN_MAX_DRAWS), entry((size_t)n_chunks);
  std::vector<int> count((size_t)n_chunks * CHAIN_MAX_DRAWS);
  std::vector<int64_t> base((size_t)n_chunks), starts((size_t)n);
  std::vector<unsigned char> chunk(CHAIN_CHUNK);
  for (int64_t c = 0; c < n_chunks; c++) {
    for (int k = 0; k < CHAIN_CHUNK; k++) chunk[k] = (c * CHAIN_CHUNK + k < n_pos) ? len[c * CHAIN_CHUNK + k] : 255;
    for (int e = 0; e < max_draws; e++) {
      int cnt = 0;
      exit_off[c * CHAIN_MAX_DRAWS + e] = (unsigned char)chain_walk(chunk.data(), e, &cnt);
      count[c * CHAIN_MAX_DRAWS + e] = cnt;
    }
  }
  int e = 0;
  int64_t b = 0;
  for (int64_t c = 0; c < n_chunks; c++) {
    entry[c] = (unsigned char)e;
    base[c] = b;
    chain_scan_step(&exit_off[c * CHAIN_MAX_DRAWS], &count[c * CHAIN_MAX_DRAWS], n, &e, &b);
  }
  int64_t end_word = 0;
  for (int64_t c = 0; c < n_chunks; c++) {
    if (entry[c] == 255 || base[c] >= n) continue;
    for (int k = 0; k < CHAIN_CHUNK; k++) chunk[k] = (c * CHAIN_CHUNK + k < n_pos) ? len[c * CHAIN_CHUNK + k] : 255;
    if (!chain_emit(chunk.data(), entry[c], base[c], n, mti0 + c * CHAIN_CHUNK, starts.data(), &end_word)) return -6;
  }
  // the n real persons
  typedef HostEnv<StreamSource> Env;
  double counts[9] = {0};
  Env env;
  env.outputs = outputs;
  env.rows = rows;
  env.geom = &geom;
  env.acc.d = dense ? dense->data() : 0;
  env.counts = counts;
  for (int64_t i = 0; i < n; i++) {
    env.src.w = w.data();
    env.src.pos = starts[i];
    env.src.end = mt_end;
    env.src.overrun = 0;
    Model<Env> person(env, consts, i);
    person.run();
    if (env.src.overrun) return -6;
    if (person.error_) return person.error_;
  }
  int64_t batch = (end_word - 1) / 624;
  memcpy(state + 1, w.data() + batch * 624, sizeof(uint32_t) * 624);
  state[0] = (uint32_t)(end_word - batch * 624);
  return 0;
}

}  // namespace

template <class SrcT>
struct HostSummaryEnv {
  typedef SrcT Src;
  Src src;
  const SummaryGeom *geom;
  PlainAcc acc;
  double u, c, tot_u, tot_c;
  void begin_person() {}
  void row(double, double, double, int, int) {}
  void report(int, int, double, double) {}
  void set_uc(int, double uu, double cc) {
    u = uu;
    c = cc;
  }
  void summary(int state, int kind, double lhs, double rhs) {
    double du, dc;
    summary_add(*geom, acc, state, kind, lhs, rhs, u, c, &du, &dc);
    tot_u += du;
    tot_c += dc;
  }
};


struct EmuCalibEnv {
  typedef MrgSource Src;
  Src src;
  double heap_store_t[calib::HEAP_SLOTS];
  int heap_store_m[calib::HEAP_SLOTS];
  double *heap_t = heap_store_t;
  int *heap_m = heap_store_m;
  int heap_stride = 1;
  double occ[40];
  double tar[4];
  int tar_n;
  void occupancy(int stage, int cind) {
    if (stage < 4) occ[stage * 10 + cind] += 1;
  }
  void at_death(double clinTime) {
    calib::time_at_risk_terms(clinTime, [this](int i, double v) {
      tar[i] += v;
      if (i + 1 > tar_n) tar_n = i + 1;
    });
  }
};

// the pending-event set as the binary heap (ordered = 0) or as the ordered array (ordered = 1; *tie_out says whether
// any person saw two pending actions compare equal, in which case the device repeats the run with the heap)
template <bool ORDERED>
static int emu_calibration_t(const double seed[6], int64_t first_person, int64_t n, const double runpar[6], int grid,
                             msim_calibration *out, int *tie_out) {
  StreamStart st;
  make_start(seed, first_person + 1, grid, &st);
  calib::Par par = {runpar[0], runpar[1], runpar[2], runpar[3], runpar[4], runpar[5]};
  EmuCalibEnv env;
  memset(env.occ, 0, sizeof env.occ);
  memset(env.tar, 0, sizeof env.tar);
  env.tar_n = 0;
  // person order (the sums are order-dependent in the last bits; emulate the reference's order)
  std::vector<Mrg32k3a> subs((size_t)n);
  for (int blk = 0; blk < grid; blk++)
    for (int tid = 0; tid < BLOCK; tid++) {
      Mrg32k3a sub = thread_start_state(st, blk * BLOCK + tid);
      for (int64_t idx = (int64_t)blk * BLOCK + tid; idx < n; idx += (int64_t)grid * BLOCK) {
        subs[idx] = sub;
        sub.jump(st.stride);
      }
    }
  for (int64_t i = 0; i < n; i++) {
    env.src.g = subs[i];
    calib::CalibPerson<EmuCalibEnv, ORDERED> person(env, par);
    person.run();
    if (person.error_) return person.error_;
    if (person.actions.tie) *tie_out = 1;
  }
  memset(out, 0, sizeof *out);
  for (int s = 0; s < 4; s++) {
    double tot = 0;
    for (int k = 0; k < 10; k++) {
      out->occupancy[s * 10 + k] = env.occ[s * 10 + k];
      tot += env.occ[s * 10 + k];
    }
    out->present[s] = tot > 0;
  }
  out->n_time_at_risk = env.tar_n;
  for (int k = 0; k < env.tar_n; k++) out->time_at_risk[k] = env.tar[k];
  return 0;
}

// The modular arithmetic of mrg32k3a.cuh against 128-bit integer arithmetic: random inputs and the extremes
// (residues m - 1, the largest products, values that make the folded sum carry).  Returns the number of mismatches.
template <uint32_t C>
static int64_t mrg_selftest_component(uint64_t seed, int64_t n) {
  const uint32_t m = 0u - C;
  int64_t bad = 0;
  uint64_t x = seed * 6364136223846793005ull + 1442695040888963407ull;
  auto next = [&]() {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
  };
  const uint32_t edge[10] = {0u, 1u, 2u, C - 1, C, C + 1, m - 2, m - 1, m / 2, 0x80000000u};
  for (int64_t i = 0; i < n; i++) {
    uint32_t v[6];
    for (int k = 0; k < 6; k++) {
      const uint64_t r = next();
      v[k] = ((r >> 60) < 6) ? edge[(r >> 32) % 10] : (uint32_t)(r % m);  // about a third of the operands are extremes
    }
    const unsigned __int128 want = ((unsigned __int128)v[0] * v[3] + (unsigned __int128)v[1] * v[4] + (unsigned __int128)v[2] * v[5]) % m;
    if (mrg_dot3<C>(v[0], v[1], v[2], v[3], v[4], v[5]) != (uint32_t)want) bad++;
    const uint64_t y = next() | ((i & 7) == 0 ? 0xffffffff00000000ull : 0ull);  // any 64-bit value, many with a full high word
    if (mrg_reduce<C>(y) != (uint32_t)(y % m)) bad++;
    const uint32_t hi = (uint32_t)(next() % (0xffffffffu / C)), lo = (i & 3) ? (uint32_t)next() : 0xffffffffu - (uint32_t)(next() % (2 * C));
    if (mrg_fold_small<C>(hi, lo) != (uint32_t)((((uint64_t)hi << 32) | lo) % m)) bad++;
  }
  return bad;
}

extern "C" {

int emu_rng_uniforms(const double seed[6], int64_t first_substream, int64_t n_substreams, int32_t draws, int grid,
                     double *out) {
  StreamStart st;
  make_start(seed, first_substream, grid, &st);
  for (int blk = 0; blk < grid; blk++)
    for (int tid = 0; tid < BLOCK; tid++) {
      unsigned gth = blk * BLOCK + tid;
      Mrg32k3a sub = thread_start_state(st, gth);
      for (int64_t j = gth; j < n_substreams; j += (int64_t)grid * BLOCK) {
        Mrg32k3a cur = sub;
        for (int k = 0; k < draws; k++) out[j * draws + k] = cur.u01();
        sub.jump(st.stride);
      }
    }
  return 0;
}

int emu_person_run(const msim_person_shard *sh, int grid, msim_table *rowlog, double counts[9], msim_report *report) {
  if (!mrg_seed_ok(sh->seed)) return MSIM_ERR_SEED;
  StreamStart st;
  make_start(sh->seed, sh->first_person + 1, grid, &st);
  person_r::Consts K = person_r::make_consts();
  ReportGeom geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate);
  std::vector<double> dense(geom.total, 0.0);
  Rows rows;
  typedef HostEnv<MrgSource> Env;
  Env env;
  env.outputs = sh->outputs;
  env.rows = &rows;
  env.geom = &geom;
  env.acc.d = dense.data();
  double cnt[9] = {0};
  env.counts = cnt;
  int64_t n_tiles = (sh->n + BLOCK - 1) / BLOCK;
  for (int blk = 0; blk < grid; blk++)
    for (int tid = 0; tid < BLOCK; tid++) {
      Mrg32k3a sub = thread_start_state(st, blk * BLOCK + tid);
      for (int64_t tile = blk; tile < n_tiles; tile += grid) {
        int64_t idx = tile * BLOCK + tid;
        if (idx < sh->n) {
          env.src.g = sub;
          person_r::Person<Env> person(env, K, sh->first_person + idx);
          person.run();
          if (person.error_) return person.error_;
        }
        sub.jump(st.stride);
      }
    }
  if (counts) memcpy(counts, cnt, sizeof cnt);
  if (rowlog) {
    sort_rows_by_person(rows);
    static const char *const cols[5] = {"endTime", "event", "id", "startTime", "state"};
    rows_to_table(rows, cols, rowlog);
  }
  if (report) dense_to_report_impl(dense.data(), geom, report);
  return 0;
}

// The staged schedule of person_staged.cuh, one warp at a time, lanes in order:
// the same stage functions and the same queue discipline (push by lane order,
// pop the top <= 32 entries) as person_staged_kernel.
// share > 0: the kernel's two waves of blocks (grid must be even): the first grid / 2 blocks deal the tiles
// [0, split) to their warps in turn, the others [split, tiles), each wave from the substream state of its range's first
// person (host_person.h, person_enqueue_staged; person_staged.cuh "waves").  share <= 0: one wave.
int emu_person_run_staged_waves(const msim_person_shard *sh, int grid, double share, msim_table *rowlog, double counts[9],
                                msim_report *report) {
  if (!mrg_seed_ok(sh->seed)) return MSIM_ERR_SEED;
  const int SB = 128, SW = 4;
  const bool two = share > 0;
  if (two && (grid % 2)) return MSIM_ERR_ARG;
  const int wave_blocks = two ? grid / 2 : grid;
  const int64_t all_tiles = (sh->n + 31) / 32;
  const int64_t split = two ? std::min<int64_t>(all_tiles, (int64_t)(share * (double)all_tiles + 0.5)) : all_tiles;
  StreamStart st, st_b;
  make_start(sh->seed, sh->first_person + 1, wave_blocks, &st, SB);
  make_start(sh->seed, sh->first_person + 1 + 32 * split, wave_blocks, &st_b, SB);
  person_r::Consts K = person_r::make_consts();
  ReportGeom geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate);
  std::vector<double> dense(geom.total, 0.0);
  Rows rows;
  typedef HostEnv<MrgSource> Env;
  Env env;
  env.outputs = sh->outputs;
  env.rows = &rows;
  env.geom = &geom;
  env.acc.d = dense.data();
  double cnt[9] = {0};
  env.counts = cnt;
  const int64_t GW = (int64_t)wave_blocks * SW;
  struct E1 {
    person_r::ParkedOnset p;
    int64_t idx;
  };
  struct E2 {
    person_r::Life p;
    int64_t idx;
  };
  int err = 0;
  for (int64_t gw_all = 0; gw_all < (two ? 2 : 1) * GW; gw_all++) {
    const bool wave_b = gw_all >= GW;
    const int64_t gw = wave_b ? gw_all - GW : gw_all;
    const StreamStart &stw = wave_b ? st_b : st;
    const int64_t n_tiles = wave_b ? all_tiles : split;  // end of the wave's range
    Mrg32k3a sub[32];
    for (int lane = 0; lane < 32; lane++) sub[lane] = thread_start_state(stw, (unsigned)(gw * 32 + lane));
    std::vector<E1> q1;
    std::vector<E2> q2;
    for (int64_t tile = (wave_b ? split : 0) + gw;;) {
      const bool last = tile >= n_tiles;
      std::vector<E2> ready;
      if (q2.size() >= 32 || (last && q1.empty() && !q2.empty())) {
        size_t m = std::min<size_t>(q2.size(), 32), base = q2.size() - m;
        for (size_t l = 0; l < m; l++) ready.push_back(q2[base + l]);
        q2.resize(base);
      } else if (q1.size() >= 32 || (last && !q1.empty())) {
        size_t m = std::min<size_t>(q1.size(), 32), base = q1.size() - m;
        std::vector<E2> parked;
        for (size_t l = 0; l < m; l++) {
          E2 o;
          o.idx = q1[base + l].idx;
          if (person_r::stage_b1<Env>(K, q1[base + l].p, &o.p)) ready.push_back(o);
          else parked.push_back(o);
        }
        q1.resize(base);
        q2.insert(q2.end(), parked.begin(), parked.end());
      } else if (!last) {
        for (int lane = 0; lane < 32; lane++) {
          int64_t i = tile * 32 + lane;
          if (i < sh->n) {
            env.src.g = sub[lane];
            E1 e;
            E2 r;
            e.idx = r.idx = i;
            if (person_r::stage_a(env, &e.p, &r.p)) ready.push_back(r);
            else q1.push_back(e);
          }
          sub[lane].jump(st.stride);
        }
        tile += GW;
      } else {
        break;
      }
      if (q1.size() >= 64 || q2.size() >= 64) return -99;  // the kernel's queues hold 64
      for (size_t l = 0; l < ready.size(); l++) person_r::run_life<false>(env, K, sh->first_person + ready[l].idx, ready[l].p, &err);
    }
  }
  if (err) return err;
  if (counts) memcpy(counts, cnt, sizeof cnt);
  if (rowlog) {
    sort_rows_by_person(rows);
    static const char *const cols[5] = {"endTime", "event", "id", "startTime", "state"};
    rows_to_table(rows, cols, rowlog);
  }
  if (report) dense_to_report_impl(dense.data(), geom, report);
  return 0;
}

int emu_person_run_staged(const msim_person_shard *sh, int grid, msim_table *rowlog, double counts[9], msim_report *report) {
  return emu_person_run_staged_waves(sh, grid, 0.0, rowlog, counts, report);
}

// the accumulator vector a shard contributes to the all-reduce: dense report cells, then the 9 counts
int emu_person_dense(const msim_person_shard *sh, int grid, double *acc, int64_t n_doubles) {
  if (!mrg_seed_ok(sh->seed)) return MSIM_ERR_SEED;
  StreamStart st;
  make_start(sh->seed, sh->first_person + 1, grid, &st);
  person_r::Consts K = person_r::make_consts();
  ReportGeom geom = ages_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate);
  const int cells = (sh->outputs & OUT_REPORT) ? geom.total : 0;
  if (n_doubles != cells + 9) return MSIM_ERR_ARG;
  memset(acc, 0, sizeof(double) * n_doubles);
  Rows rows;
  typedef HostEnv<MrgSource> Env;
  Env env;
  env.outputs = sh->outputs & ~OUT_ROWLOG;
  env.rows = &rows;
  env.geom = &geom;
  env.acc.d = acc;
  env.counts = acc + cells;
  int64_t n_tiles = (sh->n + BLOCK - 1) / BLOCK;
  for (int blk = 0; blk < grid; blk++)
    for (int tid = 0; tid < BLOCK; tid++) {
      Mrg32k3a sub = thread_start_state(st, blk * BLOCK + tid);
      for (int64_t tile = blk; tile < n_tiles; tile += grid) {
        int64_t idx = tile * BLOCK + tid;
        if (idx < sh->n) {
          env.src.g = sub;
          person_r::Person<Env> person(env, K, sh->first_person + idx);
          person.run();
          if (person.error_) return person.error_;
        }
        sub.jump(st.stride);
      }
    }
  return 0;
}

// model: 0 = callSimplePerson (row log), 1 = callSimplePerson2 (report), 2 = callIllnessDeath (report)
int emu_sequential(int model, int64_t n, double cure, double zsd, uint32_t state[625], msim_table *rowlog,
                   msim_report *report) {
  Rows rows;
  int rc;
  if (model == 0 || model == 1) {
    simple_example::Consts c = {0};
    ReportGeom geom = ages_geom(simple_example::N_STATE, simple_example::N_EVENT, 0.0);
    std::vector<double> dense(geom.total, 0.0);
    rc = sequential<simple_example::SimplePerson>(n, c, simple_example::MAX_DRAWS, model == 0 ? OUT_ROWLOG : OUT_REPORT,
                                                  geom, state, &rows, &dense);
    if (rc) return rc;
    static const char *const cols[5] = {"endtime", "event", "id", "startTime", "state"};
    if (rowlog) rows_to_table(rows, cols, rowlog);
    if (report) dense_to_report_impl(dense.data(), geom, report);
    return 0;
  }
  illness_death::Consts c = illness_consts(cure, zsd);
  ReportGeom geom = ages_geom(illness_death::N_STATE, illness_death::N_EVENT, 0.0);
  std::vector<double> dense(geom.total, 0.0);
  rc = sequential<illness_death::SimplePerson>(n, c, zsd == 0.0 ? 5 : 7, OUT_REPORT, geom, state, &rows, &dense);
  if (rc) return rc;
  if (report) dense_to_report_impl(dense.data(), geom, report);
  return 0;
}

int emu_calibration(const double seed[6], int64_t first_person, int64_t n, const double runpar[6], int grid,
                    msim_calibration *out) {
  int tie = 0;
  return emu_calibration_t<false>(seed, first_person, n, runpar, grid, out, &tie);
}
int emu_calibration_ordered(const double seed[6], int64_t first_person, int64_t n, const double runpar[6], int grid,
                            msim_calibration *out, int *tie_out) {
  *tie_out = 0;
  return emu_calibration_t<true>(seed, first_person, n, runpar, grid, out, tie_out);
}

// The calibration sweep's straight-line form (calib_shared.cuh): every person's draws once, then every parameter set
// over them by comparisons of the event times.  runpar = [n_sets][6], out = [n_sets]; *tie_out as above.  With
// force_tD > 0 every person's time of death is replaced by that value (tests: event times that coincide).
int emu_calibration_shared(const double seed[6], int64_t first_person, int64_t n, int n_sets, const double *runpar, int grid,
                           msim_calibration *out, int *tie_out, double force_tD) {
  *tie_out = 0;
  StreamStart st;
  make_start(seed, first_person + 1, grid, &st);
  std::vector<Mrg32k3a> subs((size_t)n);
  for (int blk = 0; blk < grid; blk++)
    for (int tid = 0; tid < BLOCK; tid++) {
      Mrg32k3a sub = thread_start_state(st, blk * BLOCK + tid);
      for (int64_t idx = (int64_t)blk * BLOCK + tid; idx < n; idx += (int64_t)grid * BLOCK) {
        subs[idx] = sub;
        sub.jump(st.stride);
      }
    }
  std::vector<calib::Draws> draws((size_t)n);
  std::vector<int> nC((size_t)n);
  bool tie = false;
  for (int64_t i = 0; i < n; i++) {
    MrgSource src;
    src.g = subs[i];
    draws[i] = calib::draw_person(src);
    if (force_tD > 0) draws[i].tD = force_tD;
    nC[i] = calib::counts_before(draws[i].tD, calib::MSIM_CALIB_RUNGS, &tie);
  }
  for (int s = 0; s < n_sets; s++) {
    const double *rp = runpar + 6 * s;
    if (!calib::straight_ok(rp)) return -7;
    const calib::SetPar P = {rp[0], rp[1], rp[2], rp[3], rp[4], rp[5] * rp[4]};
    msim_calibration *o = out + s;
    memset(o, 0, sizeof *o);
    for (int64_t i = 0; i < n; i++) {
      int c1, c2, c3;
      double clin;
      calib::straight_life(draws[i], nC[i], P, &c1, &c2, &c3, &clin, &tie);
      for (int k = 0; k < nC[i]; k++) o->occupancy[(k < c1 ? 0 : k < c2 ? 1 : k < c3 ? 2 : 3) * 10 + k] += 1;
      calib::time_at_risk_terms(clin, [o](int j, double v) {
        o->time_at_risk[j] += v;
        if (j + 1 > o->n_time_at_risk) o->n_time_at_risk = j + 1;
      });
    }
    for (int stg = 0; stg < 4; stg++) {
      double tot = 0;
      for (int k = 0; k < 10; k++) tot += o->occupancy[stg * 10 + k];
      o->present[stg] = tot > 0;
    }
  }
  *tie_out = tie;
  return 0;
}

// exp_rand (rmath.cuh) fed with the given uniforms, against R's algorithm as R writes it (Ahrens & Dieter 1972,
// nmath/sexp.c: the doubling loop).  Returns the number of results that differ in any bit.
struct ListSource {
  const double *u;
  int64_t pos;
  double unif_rand() { return u[pos++]; }
};
int64_t emu_exp_rand_selftest(const double *u, int64_t n, int draws_each) {
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    ListSource s1{u + i * draws_each, 0}, s2{u + i * draws_each, 0};
    RMath<ListSource> R(&s1);
    const double got = R.exp_rand();
    // R's own form
    const double *q = MSIM_EXP_Q;
    double a = 0., x = s2.unif_rand();
    while (x <= 0. || x >= 1.) x = s2.unif_rand();
    for (;;) {
      x += x;
      if (x > 1.) break;
      a += q[0];
    }
    x -= 1.;
    double want;
    if (x <= q[0]) {
      want = a + x;
    } else {
      int k = 0;
      double ustar = s2.unif_rand(), umin = ustar;
      do {
        ustar = s2.unif_rand();
        if (umin > ustar) umin = ustar;
        k++;
      } while (x > q[k]);
      want = a + umin * q[0];
    }
    if (memcmp(&got, &want, 8) != 0 || s1.pos != s2.pos) bad++;
  }
  return bad;
}

int emu_counts_before(double t, int *tie_out) {
  bool tie = false;
  const int c = calib::counts_before(t, calib::MSIM_CALIB_RUNGS, &tie);
  *tie_out = tie;
  return c;
}

// OrderedEvents against EventHeap on a random schedule/pop(/cancel) sequence (engine.cuh).  Times are drawn from
// `n_times` distinct values and priorities from `n_prios`, so small values force pairs that compare equal.  With
// cancel_every > 0 every so many operations cancel one event kind: the heap marks those entries A_Ignore and pops them
// later without effect (ssim.cc:131-142, 171-175), the queue removes them; what must agree is the sequence of LIVE
// pops.  Returns -1 if the queue failed to flag a pair comparing equal (checked by brute force over the pending set
// at every insert), otherwise the number of live pops that differed from the heap's; *tie_out = the queue's flag.
int emu_ordered_vs_heap(uint32_t seed, int ops, int n_times, int n_prios, int cancel_every, int *tie_out) {
  constexpr int CAP = 14;
  double ht[4 * CAP], qt[CAP];
  int hm[4 * CAP], qm[CAP];
  EventHeap<-4 * CAP> heap;  // lingering cancelled entries need room
  OrderedEvents<CAP> queue;
  heap.bind(ht, hm, 1);
  queue.bind(qt, qm, 1);
  heap.reset();
  queue.reset();
  std::vector<std::pair<double, int> > pending;  // live entries
  bool equal_pair = false;
  int differ = 0;
  uint32_t x = seed ? seed : 1u;
  auto next = [&x]() {
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    return x;
  };
  for (int i = 0; i < ops; i++) {
    if (cancel_every > 0 && i % cancel_every == cancel_every - 1) {
      const int kind = (int)(next() % 7);
      heap.for_each_meta([kind](int &m) {
        if (meta_type(m) == A_Event && meta_kind(m) == kind) m = (int)(((uint32_t)m & 0x3fffffffu) | ((uint32_t)A_Ignore << 30));
      });
      queue.remove_if([kind](int m) { return meta_type(m) == A_Event && meta_kind(m) == kind; });
      for (size_t k = pending.size(); k-- > 0;)
        if (meta_kind(pending[k].second) == kind) pending.erase(pending.begin() + (long)k);
      continue;
    }
    const bool push = queue.n == 0 || (queue.n < CAP && heap.n < 4 * CAP && next() % 5 < 3);
    if (push) {
      const double t = (double)(next() % (uint32_t)n_times) * 0.25;
      const int prio = (int)(next() % (uint32_t)n_prios) - 1;
      const int m = pack_meta(A_Event, (int)(next() % 7), prio);
      for (size_t k = 0; k < pending.size(); k++)
        if (pending[k].first == t && meta_prio(pending[k].second) == prio) equal_pair = true;
      pending.push_back(std::make_pair(t, m));
      if (!heap.insert(t, m) || !queue.insert(t, m)) return -2;
    } else {
      double t1 = 0, t2;
      int m1 = 0, m2;
      bool live = false;
      while (heap.n > 0 && !live) {  // the reference's loop: ignored actions are popped and skipped
        heap.pop_first(t1, m1);
        live = meta_type(m1) != A_Ignore;
      }
      if (!live) return -5;  // the queue says there is a live action, the heap has none
      queue.pop_first(t2, m2);
      if (t1 != t2 || m1 != m2) differ++;
      for (size_t k = 0; k < pending.size(); k++)
        if (pending[k].first == t2 && pending[k].second == m2) {
          pending.erase(pending.begin() + (long)k);
          break;
        }
    }
    if ((size_t)queue.n != pending.size()) return -3;
    if (equal_pair && !queue.tie) return -1;
  }
  *tie_out = queue.tie ? 1 : 0;
  if (!equal_pair && queue.tie) return -4;  // a false alarm costs a repeat, but there should be none
  return differ;
}

// callSim (Sick-Sicker, SummaryReport) with the device's decomposition: a first pass for every person's last
// set_uc() code, the carry scan, then the persons.  substreams = 0 reads R's Mersenne-Twister stream from `state`
// (advanced on return); substreams = 1 uses substream first_person+i+1 of the stream whose seed is `seed`.
int emu_sick_sicker(int64_t n, const double par17[17], int indivp, int substreams, int64_t first_person, const double seed[6],
                    uint32_t state[625], msim_report *out) {
  sick_sicker::Par P;
  P.r_HS1 = par17[0]; P.r_HD = par17[1]; P.r_S1H = par17[2]; P.r_S1S2 = par17[3]; P.r_S1D = par17[4]; P.r_S2D = par17[5];
  P.c_H = par17[6]; P.c_S1 = par17[7]; P.c_S2 = par17[8]; P.c_Trt = par17[9];
  P.u_H = par17[10]; P.u_S1 = par17[11]; P.u_S2 = par17[12]; P.u_Trt = par17[13];
  P.Trt = par17[16] != 0.0;
  SummaryGeom G;
  if (!make_summary_geom(&G, sick_sicker::N_STATE, sick_sicker::N_EVENT, 0.0, 31.0, par17[15], 1.0e100, par17[14], par17[14]))
    return MSIM_ERR_ARG;
  std::vector<double> dense(G.total, 0.0), indiv(2 * (size_t)std::max<int64_t>(indivp ? n : 1, 1), 0.0);
  std::vector<unsigned char> code_out((size_t)std::max<int64_t>(n, 1)), code_in((size_t)std::max<int64_t>(n, 1));
  std::vector<Mrg32k3a> subs;
  std::vector<int64_t> starts;
  std::vector<uint32_t> w;
  int64_t mti0 = 0, mt_end = 0, end_word = 0;
  if (substreams) {
    StreamStart st;
    make_start(seed, first_person + 1, 1, &st);
    subs.resize((size_t)n);
    Mrg32k3a b;
    memcpy(b.s, st.base, sizeof b.s);
    for (int64_t i = 0; i < n; i++) {
      subs[i] = b;
      b.jump(mrg_substream_jump());
    }
    for (int64_t i = 0; i < n; i++) {
      HostNullEnv<MrgSource> env;
      env.src.g = subs[i];
      sick_sicker::SickSicker<HostNullEnv<MrgSource> > person(env, P, 0);
      person.run();
      if (person.error_) return person.error_;
      code_out[i] = (unsigned char)env.code;
    }
  } else {
    mti0 = state[0];
    const int64_t n_batches = (mti0 + 250 * n + 251 + 623) / 624;
    mt_end = n_batches * 624;
    w = mt_words_emulated(state, n_batches);
    starts.resize((size_t)n);
    int64_t pos = mti0;
    for (int64_t i = 0; i < n; i++) {  // (the chain resolution itself is exercised by the other stream models)
      HostNullEnv<StreamSource> env;
      env.src.w = w.data();
      env.src.pos = pos;
      env.src.end = mt_end;
      env.src.overrun = 0;
      sick_sicker::SickSicker<HostNullEnv<StreamSource> > person(env, P, 0);
      person.run();
      if (env.src.overrun) return -6;
      if (person.error_) return person.error_;
      starts[i] = pos;
      code_out[i] = (unsigned char)env.code;
      pos = env.src.pos;
    }
    end_word = pos;
  }
  int carry = 0;
  if (substreams && first_person > 0) {  // a shard inherits the last code of the persons before it (host_summary.h)
    for (int64_t hi_p = first_person, wdw = 256; hi_p > 0 && carry == 0; wdw *= 16) {
      const int64_t lo_p = std::max<int64_t>(0, hi_p - wdw);
      StreamStart st2;
      make_start(seed, lo_p + 1, 1, &st2);
      Mrg32k3a b2;
      memcpy(b2.s, st2.base, sizeof b2.s);
      for (int64_t i = lo_p; i < hi_p; i++) {
        HostNullEnv<MrgSource> env;
        env.src.g = b2;
        sick_sicker::SickSicker<HostNullEnv<MrgSource> > person(env, P, 0);
        person.run();
        if (env.code) carry = env.code;
        b2.jump(mrg_substream_jump());
      }
      hi_p = lo_p;
    }
  }
  for (int64_t i = 0; i < n; i++) {
    code_in[i] = (unsigned char)carry;
    if (code_out[i]) carry = code_out[i];
  }
  for (int64_t i = 0; i < n; i++) {
    double tu, tc;
    if (substreams) {
      HostSummaryEnv<MrgSource> env;
      env.geom = &G;
      env.acc.d = dense.data();
      env.src.g = subs[i];
      sick_sicker::uc_of_code(P, code_in[i], &env.u, &env.c);
      env.tot_u = env.tot_c = 0;
      sick_sicker::SickSicker<HostSummaryEnv<MrgSource> > person(env, P, first_person + i);
      person.run();
      tu = env.tot_u;
      tc = env.tot_c;
    } else {
      HostSummaryEnv<StreamSource> env;
      env.geom = &G;
      env.acc.d = dense.data();
      env.src.w = w.data();
      env.src.pos = starts[i];
      env.src.end = mt_end;
      env.src.overrun = 0;
      sick_sicker::uc_of_code(P, code_in[i], &env.u, &env.c);
      env.tot_u = env.tot_c = 0;
      sick_sicker::SickSicker<HostSummaryEnv<StreamSource> > person(env, P, first_person + i);
      person.run();
      tu = env.tot_u;
      tc = env.tot_c;
    }
    if (indivp) {
      indiv[2 * i] = tu;
      indiv[2 * i + 1] = tc;
    } else {
      indiv[0] += tu;
      indiv[1] += tc;
    }
  }
  dense_to_summary_impl(dense.data(), G, indiv.data(), indivp ? n : 1, out);
  if (!substreams && n > 0) {
    int64_t batch = (end_word - 1) / 624;
    memcpy(state + 1, w.data() + batch * 624, sizeof(uint32_t) * 624);
    state[0] = (uint32_t)(end_word - batch * 624);
  }
  return 0;
}

// the device's gsm functions (gsm.cuh) on the host: mode 0 randU, 1 randU0, 2 eta
int emu_gsm(const msim_gsm *m, int mode, int64_t n, const double *u, const double *tentry, const int32_t *index, double scale,
            double *out) {
  static GsmModel M;
  memset(&M, 0, sizeof M);
  M.n_terms = m->n_terms;
  M.log_time = m->log_time != 0;
  M.n_index = m->n_index;
  M.tmin = m->tmin / m->inflate;
  M.tmax = m->tmax * m->inflate;
  M.t0 = m->t0;
  M.etap = m->etap;
  M.etap0 = m->etap0;
  for (int i = 0; i < m->n_terms; i++) {
    const msim_gsm_term &t = m->terms[i];
    GsmTerm &T = M.terms[i];
    const int df = (t.intercept != 0) + 3 + t.n_knots;
    const bool tr = t.q_cols < t.q_rows;
    const int rows = tr ? t.q_cols : t.q_rows, cols = tr ? t.q_rows : t.q_cols;
    if (cols != df) return -3;
    std::vector<double> q((size_t)rows * cols);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) q[(size_t)r * cols + c] = tr ? t.q_const[c * t.q_cols + r] : t.q_const[r * t.q_cols + c];
    GsmTermSpec spec = {t.n_knots, t.intercept != 0, rows, t.Boundary_knots[0], t.Boundary_knots[1], t.knots, q.data(), t.gamma};
    if (!gsm_term_build(T, spec)) return -3;
    M.x[i] = t.x;
  }
  for (int64_t i = 0; i < n; i++) {
    int idx = index ? index[i] : 0;
    out[i] = mode == 0 ? gsm_randU(M, u[i], tentry ? tentry[i] : 0.0, idx, scale)
             : mode == 1 ? gsm_randU0(M, u[i], idx, scale) : gsm_eta(M, idx, u[i], false);
  }
  return 0;
}

// the device's lookup functions (lookups.cuh) on the host; families and modes as in include/msim.h
int emu_rpexp(int mode, int n_t, const double *h, const double *t, int64_t n, const double *a, const double *b, double *out) {
  std::vector<double> Hi(n_t);
  rpexp_cumulate(n_t, h, t, Hi.data());
  Rpexp T = {n_t, Hi.data(), h, t};
  for (int64_t i = 0; i < n; i++) {
    int failed = 0;
    out[i] = mode == 0 ? rpexp_rand(T, a[i], b ? b[i] : 0.0, &failed)
             : mode == 1 ? rpexp_H(T, a[i]) : mode == 2 ? rpexp_h(T, a[i]) : mode == 3 ? rpexp_p(T, a[i], true) : rpexp_d(T, a[i]);
    if (failed) return -3;
  }
  return 0;
}
int emu_rpexp_gamma(int mode, int n_t, double theta, const double *mu, const double *t, int64_t n, const double *a, const double *Z,
                    const double *from, double *out) {
  std::vector<double> H0(n_t);
  rpexp_gamma_cumulate(n_t, theta, mu, t, H0.data());
  RpexpGamma T = {n_t, theta, H0.data(), mu, t};
  for (int64_t i = 0; i < n; i++)
    out[i] = mode == 0 ? rpexp_gamma_rand(T, a[i], Z[i], from ? from[i] : 0.0)
             : mode == 1 ? rpexp_gamma_H(T, a[i], Z[i]) : mode == 2 ? rpexp_gamma_h(T, a[i], Z[i]) : rpexp_gamma_p(T, a[i], Z[i], true);
  return 0;
}
int emu_interpolate(int mode, int n_pts, const double *x, const double *y, int64_t n, const double *xfind, double *out) {
  std::vector<double> slope(n_pts - 1);
  for (int i = 0; i < n_pts - 1; i++) slope[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
  Interp T = {n_pts, x, y, slope.data()};
  for (int64_t i = 0; i < n; i++)
    out[i] = mode == 0 ? interp_approx(T, xfind[i]) : mode == 1 ? interp_step(T, xfind[i])
             : mode == 2 ? interp_invert(T, xfind[i]) : interp_invert_decreasing(T, xfind[i]);
  return 0;
}

// CostReport: n point costs (state, time, cost) into a report over setPartition(start, finish, delta); out_rows gets
// (state, age, cost) per non-empty cell in key order, *current the individual's running total
int emu_cost_report(double rate, double start_age, double p_start, double p_finish, double p_delta, int64_t n, const int *state,
                    const double *time, const double *cost, double *out_rows, int64_t cap, double *current) {
  SummaryGeom G;
  if (!make_summary_geom(&G, 8, 1, p_start, p_finish, p_delta, 1.0e100, 0.0, rate)) return -1;
  std::vector<double> dense(G.total, 0.0);
  PlainAcc acc;
  acc.d = dense.data();
  double cur = 0.0;
  for (int64_t i = 0; i < n; i++) cur += cost_report_add(G, acc, state[i], time[i], cost[i], rate, start_age);
  *current = cur;
  int64_t k = 0;
  for (int s = 0; s < G.n_state; s++)
    for (int b = 0; b < G.n_bin; b++)
      if (dense[G.off_pcn + s * G.n_bin + b] > 0 && k < cap) {
        out_rows[3 * k] = s;
        out_rows[3 * k + 1] = G.part[b];
        out_rows[3 * k + 2] = dense[G.off_cost + s * G.n_bin + b];
        k++;
      }
  return (int)k;
}

void emu_mt_set_seed(uint32_t seed, uint32_t state[625]) {
  HostRngState r;
  r.mt_set_seed(seed);
  memcpy(state, r.mt, sizeof r.mt);
}

// first n uniforms from an MT state (three-wave regeneration), state not advanced
void emu_mt_uniforms(const uint32_t state[625], int64_t n, double *out) {
  int64_t mti0 = state[0];
  int64_t n_batches = (mti0 + n + 623) / 624;
  std::vector<uint32_t> w = mt_words_emulated(state, n_batches);
  for (int64_t i = 0; i < n; i++) out[i] = mt_word_to_unif(w[mti0 + i]);
}

int emu_advance_substream(double seed[6], int n) {
  HostStream r;
  uint32_t s[6];
  seed_to_u32(seed, s);
  r.set_all(s);
  r.advance_substream(0, n);
  for (int i = 0; i < 6; i++) seed[i] = (double)r.Cg.s[i];
  return 0;
}

void emu_table_free(msim_table *t) {
  free(t->data);
  t->data = 0;
}
void emu_report_free(msim_report *r) {
  emu_table_free(&r->pt);
  emu_table_free(&r->ut);
  emu_table_free(&r->events);
  emu_table_free(&r->prev);
  emu_table_free(&r->costs);
  emu_table_free(&r->indiv);
}


int64_t emu_mrg_selftest(uint64_t seed, int64_t n) {
  return mrg_selftest_component<MRG_C1>(seed, n) + mrg_selftest_component<MRG_C2>(seed + 1, n);
}


// csrc/fastmath.cuh on the host (the device runs the same source): mode 0 exp(x), 1 log(x), 2 pow(x, c)
void emu_fastmath(int mode, int64_t n, const double *x, double c, double *out) {
  for (int64_t i = 0; i < n; i++) out[i] = mode == 0 ? fm_exp(x[i]) : (mode == 1 ? fm_log(x[i]) : fm_pow(x[i], c));
}

}  // extern "C"
