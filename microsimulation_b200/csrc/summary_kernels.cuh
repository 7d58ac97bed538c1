// Kernels for models that report through SummaryReport (summary.cuh); the one
// such model the reference documents is Sick-Sicker (README.org:148-301).
//
// Two regimes, as for the other models: R's single Mersenne-Twister stream
// (the README's runs: set.seed(12345), draws consumed person after person — the
// stream-offset resolution of kernels.cuh applies, with up to 250 draws per
// person), or one RngStream substream per person.
//
// The report's utility and cost carry over from one person to the next in the
// reference (models.cuh, sick_sicker::uc_of_code).  A first pass therefore
// finds each person's LAST set_uc() code (sick_code_kernel, or the draw-count
// pass in the stream regime), carry_tile_kernel / carry_apply_kernel turn that into the code each
// person starts with, and only then does sick_kernel run the persons.
#pragma once
#include <new>

#include "kernels.cuh"
#include "summary.cuh"

namespace msim {

template <bool ORDERED>
struct SickTraitsT {
  typedef sick_sicker::Par Consts;
  template <class Env> using Model = sick_sicker::SickSicker<Env, ORDERED>;
  static constexpr int MAX_ROWS = 1;
};
typedef SickTraitsT<false> SickTraits;

struct SickParams {
  StreamStart st;             // substream regime
  const uint32_t *mt_words;   // stream regime
  int64_t mt_end;
  const int64_t *starts;
  int64_t first_person, n;
  sick_sicker::Par par;
  const SummaryGeom *geom;       // in global memory
  const unsigned char *code_in;  // [n] the set_uc() code in force when person i starts
  unsigned char *code_out;       // [n] (first pass) person i's last set_uc() code, 0 = none
  double *dense;                 // geom->total accumulators
  double *indiv;                 // [2n] (utilities, costs) per person if indivp, else [2]
  int indivp;
  int *error;
  int *tie;  // set if two pending actions of a person compared equal (ORDERED only)
};

// Lives of this model differ widely in length (1 to ~40 events), so with 32 persons side by side a warp waits for its
// longest life: 8 of 32 lanes busy (profiles/r01c_secondary_kernels.md).  The loop below lets every lane walk ITS OWN
// list of persons (lane l of the grid takes persons l, l + threads, l + 2 threads, ...: the same persons as before, so
// its substream state still advances by one fixed jump per person), one event per pass, and restarts lanes whose
// person has ended once SICK_REFILL_MIN of them wait (or nothing else is running).  Measured (1e6 persons, one
// substream each, kernels of the whole call): restart threshold 1 lane 5.68 ms, 4: 5.30, 8: 5.05, 16: 4.80,
// 32 (the warp's lanes start together): 3.84 — against 4.1 ms for 32 persons side by side through person.run().
// Finer restarts LOSE: the start-up code (substream jump, init(), three schedules) then runs in nearly every pass
// for a few lanes, and costs more than the waiting it removes; what the event-per-pass form does buy is a loop body
// the compiler keeps tight.  So the threshold stays at the full warp.
#ifndef MSIM_SICK_REFILL
#define MSIM_SICK_REFILL 32
#endif
constexpr int SICK_REFILL_MIN = MSIM_SICK_REFILL;

// first pass, substream regime: each person's last set_uc() code
template <bool ORDERED>
__global__ void __launch_bounds__(BLOCK) sick_code_kernel(SickParams p) {
  typedef NullEnv<MrgSource> Env;
  typedef sick_sicker::SickSicker<Env, ORDERED> Person;
  Mrg32k3a sub = thread_start_state(p.st, blockIdx.x * BLOCK + threadIdx.x);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  int64_t next = (int64_t)blockIdx.x * BLOCK + threadIdx.x, cur = 0;
  bool have = false;
  Env env;
  Person person(env, p.par, 0);
  for (;;) {
    __syncwarp();
    const bool want = !have && next < p.n;
    const unsigned wm = __ballot_sync(0xffffffffu, want), hm = __ballot_sync(0xffffffffu, have);
    if ((wm | hm) == 0) break;
    if (want && (__popc(wm) >= SICK_REFILL_MIN || hm == 0)) {
      cur = next;
      next += stride;
      env.code = 0;
      env.src.g = sub;
      sub.jump(p.st.stride);
      new (&person) Person(env, p.par, 0);
      person.begin();
      have = true;
    }
    __syncwarp();
    if (have) {
      have = person.alive() && person.step();
      if (!have) {
        p.code_out[cur] = (unsigned char)env.code;
        if (person.error_) atomicCAS(p.error, 0, person.error_);
        if (ORDERED && person.actions.tie) *p.tie = 1;
      }
    }
  }
}

// stream regime: person i's last code is the one found for its start position
__global__ void gather_code_kernel(const unsigned char *aux, const int64_t *starts, int64_t first_word, int64_t n,
                                   unsigned char *code_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    code_out[i] = aux[starts[i] - first_word];
}

// code_in[i] = last non-zero code_out[j], j < i (code0 if none): an exclusive scan under "last writer wins", in two
// launches over tiles of CARRY_TILE persons: every tile's last non-zero code (carry_tile_kernel), then per tile the
// carry from the tiles before it and the scan inside it (carry_apply_kernel).
constexpr int CARRY_TILE = 4096;
constexpr int CARRY_PER = CARRY_TILE / 256;  // consecutive persons per thread

// last non-zero among this thread's CARRY_PER consecutive codes, then among the block's threads up to and including t
__device__ __forceinline__ int carry_block_inclusive(int mine, int *s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o && incl == 0) incl = t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (incl == 0)
    for (int w = warp - 1; w >= 0 && incl == 0; w--) incl = s_w[w];
  return incl;
}

__global__ void __launch_bounds__(256) carry_tile_kernel(const unsigned char *code_out, int64_t n, unsigned char *tile_last) {
  __shared__ int s_w[8];
  const int64_t lo = (int64_t)blockIdx.x * CARRY_TILE + (int64_t)threadIdx.x * CARRY_PER;
  int mine = 0;
#pragma unroll
  for (int k = 0; k < CARRY_PER; k++) {
    const int v = (lo + k < n) ? code_out[lo + k] : 0;
    if (v) mine = v;
  }
  const int incl = carry_block_inclusive(mine, s_w);
  if (threadIdx.x == 255) tile_last[blockIdx.x] = (unsigned char)incl;
}

__global__ void __launch_bounds__(256) carry_apply_kernel(const unsigned char *code_out, int64_t n, int code0,
                                                          const unsigned char *tile_last, unsigned char *code_in) {
  __shared__ int s_w[8];
  __shared__ int s_before;
  if (threadIdx.x == 0) {
    int before = code0;  // last non-zero code of the tiles before this one
    for (int64_t j = (int64_t)blockIdx.x - 1; j >= 0; j--)
      if (tile_last[j]) {
        before = tile_last[j];
        break;
      }
    s_before = before;
  }
  const int64_t lo = (int64_t)blockIdx.x * CARRY_TILE + (int64_t)threadIdx.x * CARRY_PER;
  int v[CARRY_PER], mine = 0;
#pragma unroll
  for (int k = 0; k < CARRY_PER; k++) {
    v[k] = (lo + k < n) ? code_out[lo + k] : 0;
    if (v[k]) mine = v[k];
  }
  const int incl = carry_block_inclusive(mine, s_w);  // (contains the barrier that publishes s_before)
  // exclusive over threads: the inclusive value of the thread before this one
  int carry = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) == 0) {
    carry = 0;
    for (int w = (threadIdx.x >> 5) - 1; w >= 0 && carry == 0; w--) carry = s_w[w];
  }
  if (carry == 0) carry = s_before;
#pragma unroll
  for (int k = 0; k < CARRY_PER; k++) {
    if (lo + k < n) code_in[lo + k] = (unsigned char)carry;
    if (v[k]) carry = v[k];
  }
}

template <class SrcT>
struct SummaryEnv {
  typedef SrcT Src;
  Src src;
  const SummaryGeom *geom;  // shared-memory copy
  SharedAcc acc;
  double u, c;      // the report's current utility and cost
  double tot_u, tot_c;  // this person's discounted totals
  __device__ __forceinline__ void begin_person() {}
  __device__ __forceinline__ void row(double, double, double, int, int) {}
  __device__ __forceinline__ void report(int, int, double, double) {}
  __device__ __forceinline__ void set_uc(int, double uu, double cc) {
    u = uu;
    c = cc;
  }
  __device__ __forceinline__ void summary(int state, int kind, double lhs, double rhs) {
    double du, dc;
    summary_add(*geom, acc, state, kind, lhs, rhs, u, c, &du, &dc);
    tot_u += du;
    tot_c += dc;
  }
  // SummaryReport::addPointCost(state, cost, id) at the current time
  __device__ __forceinline__ void point_cost(int state, double now, double cost) { tot_c += summary_point_cost(*geom, acc, state, now, cost); }
};

template <bool STREAM, bool ORDERED>
__global__ void __launch_bounds__(BLOCK) sick_kernel(SickParams p) {
  typedef typename std::conditional<STREAM, StreamSource, MrgSource>::type Src;
  typedef SummaryEnv<Src> Env;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SummaryGeom *sg = reinterpret_cast<SummaryGeom *>(smem_raw);
  {
    const int *src = reinterpret_cast<const int *>(p.geom);
    int *dst = reinterpret_cast<int *>(sg);
    for (int k = threadIdx.x; k < (int)(sizeof(SummaryGeom) / 4); k += BLOCK) dst[k] = src[k];
  }
  __syncthreads();
  const int total = sg->total, n_f = sg->n_float;
  double *s_f = reinterpret_cast<double *>(smem_raw + ((sizeof(SummaryGeom) + 15) & ~(size_t)15));
  int *s_i = reinterpret_cast<int *>(s_f + n_f);
  Env env;
  env.geom = sg;
  env.acc.f = s_f;
  env.acc.i = s_i;
  env.acc.n_f = n_f;
  shared_acc_clear(env.acc, total);
  __syncthreads();

  Mrg32k3a sub;
  if constexpr (!STREAM) sub = thread_start_state(p.st, blockIdx.x * BLOCK + threadIdx.x);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  int err = 0;
  double all_u = 0.0, all_c = 0.0;
  // every lane walks its own list of persons, one event per pass (see sick_code_kernel)
  typedef sick_sicker::SickSicker<Env, ORDERED> Person;
  int64_t next = (int64_t)blockIdx.x * BLOCK + threadIdx.x, cur = 0;
  bool have = false;
  Person person(env, p.par, 0);
  for (;;) {
    __syncwarp();
    const bool want = !have && next < p.n;
    const unsigned wm = __ballot_sync(0xffffffffu, want), hm = __ballot_sync(0xffffffffu, have);
    if ((wm | hm) == 0) break;
    if (want && (__popc(wm) >= SICK_REFILL_MIN || hm == 0)) {
      cur = next;
      next += stride;
      if constexpr (STREAM) {
        env.src.w = p.mt_words;
        env.src.pos = p.starts[cur];
        env.src.end = p.mt_end;
        env.src.overrun = 0;
      } else {
        env.src.g = sub;
        sub.jump(p.st.stride);
      }
      sick_sicker::uc_of_code(p.par, p.code_in[cur], &env.u, &env.c);
      env.tot_u = env.tot_c = 0.0;
      new (&person) Person(env, p.par, p.first_person + cur);
      person.begin();
      have = true;
    }
    __syncwarp();
    if (have) {
      have = person.alive() && person.step();
      if (!have) {
        if (person.error_) err = person.error_;
        if (ORDERED && person.actions.tie) *p.tie = 1;
        if constexpr (STREAM) {
          if (env.src.overrun) err = -6;
        }
        if (p.indivp) {
          p.indiv[2 * cur] = env.tot_u;
          p.indiv[2 * cur + 1] = env.tot_c;
        } else {
          all_u += env.tot_u;
          all_c += env.tot_c;
        }
      }
    }
  }
  if (!p.indivp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      all_u += __shfl_xor_sync(0xffffffffu, all_u, o);
      all_c += __shfl_xor_sync(0xffffffffu, all_c, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&p.indiv[0], all_u);
      atomicAdd(&p.indiv[1], all_c);
    }
  }
  if (err) atomicCAS(p.error, 0, err);
  __syncthreads();
  shared_acc_flush(env.acc, total, p.dense);
}

// The cancer-management example (models.cuh, cancer_mgmt): SummaryReport::add for every interval and addPointCost at
// onset and at every management visit; one substream per person.
struct MgmtParams {
  StreamStart st;
  int64_t first_person, n;
  cancer_mgmt::Par par;
  const SummaryGeom *geom;  // in global memory
  double *dense;            // geom->total accumulators
  double *indiv;            // [2n] (utilities, costs) per person if indivp, else [2]
  int indivp;
  int *error;
};

__global__ void __launch_bounds__(BLOCK) mgmt_kernel(MgmtParams p) {
  typedef SummaryEnv<MrgSource> Env;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SummaryGeom *sg = reinterpret_cast<SummaryGeom *>(smem_raw);
  {
    const int *src = reinterpret_cast<const int *>(p.geom);
    int *dst = reinterpret_cast<int *>(sg);
    for (int k = threadIdx.x; k < (int)(sizeof(SummaryGeom) / 4); k += BLOCK) dst[k] = src[k];
  }
  __syncthreads();
  const int total = sg->total, n_f = sg->n_float;
  double *s_f = reinterpret_cast<double *>(smem_raw + ((sizeof(SummaryGeom) + 15) & ~(size_t)15));
  int *s_i = reinterpret_cast<int *>(s_f + n_f);
  Env env;
  env.geom = sg;
  env.acc.f = s_f;
  env.acc.i = s_i;
  env.acc.n_f = n_f;
  shared_acc_clear(env.acc, total);
  __syncthreads();
  Mrg32k3a sub = thread_start_state(p.st, blockIdx.x * BLOCK + threadIdx.x);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  int err = 0;
  double all_u = 0.0, all_c = 0.0;
  for (int64_t i0 = (int64_t)blockIdx.x * BLOCK + (threadIdx.x & ~31); i0 < p.n; i0 += stride) {
    const int64_t i = i0 + (threadIdx.x & 31);
    __syncwarp();  // every pass starts with the warp together
    if (i < p.n) {
      env.src.g = sub;
      env.u = 1.0;  // the report's utility and cost are never set by this model: the constructor's (microsimulation.h:1030-1031)
      env.c = 0.0;
      env.tot_u = env.tot_c = 0.0;
      cancer_mgmt::SimplePerson<Env> person(env, p.par, p.first_person + i);
      person.run();
      if (person.error_) err = person.error_;
      if (p.indivp) {
        p.indiv[2 * i] = env.tot_u;
        p.indiv[2 * i + 1] = env.tot_c;
      } else {
        all_u += env.tot_u;
        all_c += env.tot_c;
      }
    }
    __syncwarp();
    sub.jump(p.st.stride);
  }
  if (!p.indivp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      all_u += __shfl_xor_sync(0xffffffffu, all_u, o);
      all_c += __shfl_xor_sync(0xffffffffu, all_c, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&p.indiv[0], all_u);
      atomicAdd(&p.indiv[1], all_c);
    }
  }
  if (err) atomicCAS(p.error, 0, err);
  __syncthreads();
  shared_acc_flush(env.acc, total, p.dense);
}

// CostReport::add (microsimulation.h:1265-1278) for a batch of point costs, all booked to one individual: the cost
// cells of a SummaryGeom in global memory and the individual's running total `current`
__global__ void cost_report_kernel(const SummaryGeom *geom, int64_t n, const int *state, const double *time, const double *cost,
                                   double rate, double start_age, double *dense, double *current, int *error) {
  struct GlobalAcc {
    double *d;
    __device__ __forceinline__ void addf(int i, double v) { atomicAdd(&d[i], v); }
    __device__ __forceinline__ void addi(int i, int v) { atomicAdd(&d[i], (double)v); }
  } acc = {dense};
  double cur = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (state[i] < 0 || state[i] >= geom->n_state) {
      atomicCAS(error, 0, -3);
      continue;
    }
    cur += cost_report_add(*geom, acc, state[i], time[i], cost[i], rate, start_age);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cur += __shfl_xor_sync(0xffffffffu, cur, o);
  if ((threadIdx.x & 31) == 0 && cur != 0.0) atomicAdd(current, cur);
}

inline size_t sick_kernel_smem(const SummaryGeom &g) {
  return ((sizeof(SummaryGeom) + 15) & ~(size_t)15) + (size_t)g.n_float * 8 + (size_t)(g.total - g.n_float) * 4;
}

// parity hook for the discounting free functions: out[i] = discountedInterval(y[i], a[i], b[i], rate) (mode 0) or
// discountedPoint(y[i], a[i], rate) (mode 1)
__global__ void discount_kernel(int mode, int64_t n, const double *y, const double *a, const double *b, double rate, double *out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mode == 0 ? discountedInterval(y[i], a[i], b[i], rate) : discountedPoint(y[i], a[i], rate);
}

}  // namespace msim
