// CUDA kernels (sm_100a) for the cohort simulations and the RNG hooks.
//
// Work decomposition ("kernels"):
//   * one person per thread; a block works on "tiles" of BLOCK consecutive
//     persons, tile = it*gridDim.x + blockIdx.x (persistent grid sized to the
//     SM count).  In the substream regime a thread's next person is therefore
//     always gridDim.x*BLOCK substreams further on: ONE 3x3 mat-vec pair (the
//     `stride` jump) per person replaces the reference's per-person
//     ResetNextSubstream (RngStream.cpp:381-387);
//   * in the sequential-stream regime (R's single Mersenne-Twister stream,
//     consumed by person after person with a data-dependent number of draws)
//     the stream is generated once (mt_generate_kernel), the number of draws a
//     person would take is computed for EVERY stream position
//     (seq_len_kernel), the chain start[i+1] = start[i] + len[start[i]] is
//     resolved chunk-wise (chain_*_kernel), and then the n real persons run in
//     parallel from their resolved offsets;
//   * the life-history row log must come out in person order with a
//     data-dependent number of rows per person: warps pack their rows into a
//     provisional area at fixed offsets, and a bandwidth-bound second pass scans
//     the per-tile counts and expands them into the final columns (no spin
//     waits, no cross-warp synchronisation inside the simulation kernel);
//   * report tables and per-kind counts accumulate in shared memory and are
//     flushed once per block with `red.global.add.f64`.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "gsm.cuh"
#include "lookups.cuh"
#include "models.cuh"
#include "report.cuh"
#include "seqstream.cuh"

namespace msim {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr unsigned OUT_ROWLOG = 1u, OUT_REPORT = 2u, OUT_COUNTS = 4u, OUT_INDIV = 8u;
constexpr int N_COUNTS = 9;  // up to 8 per-kind counts + sum(endTime)

// ---- small device helpers ------------------------------------------------

// dense accumulators in shared memory: [0, n_f) doubles, [n_f, total) ints
struct SharedAcc {
  double *f;
  int *i;
  int n_f;
  __device__ __forceinline__ void addf(int idx, double v) { atomicAdd(&f[idx], v); }
  __device__ __forceinline__ void addi(int idx, int v) { atomicAdd(&i[idx - n_f], v); }
  __device__ __forceinline__ void addc(int idx) { atomicAdd(&i[idx - n_f], 1); }
};

__device__ __forceinline__ void shared_acc_clear(SharedAcc &a, int total) {
  for (int k = threadIdx.x; k < a.n_f; k += blockDim.x) a.f[k] = 0.0;
  for (int k = threadIdx.x; k < total - a.n_f; k += blockDim.x) a.i[k] = 0;
}
__device__ __forceinline__ void shared_acc_flush(const SharedAcc &a, int total, double *g) {
  for (int k = threadIdx.x; k < a.n_f; k += blockDim.x) {
    double v = a.f[k];
    if (v != 0.0) atomicAdd(&g[k], v);
  }
  for (int k = threadIdx.x; k < total - a.n_f; k += blockDim.x) {
    int v = a.i[k];
    if (v != 0) atomicAdd(&g[a.n_f + k], (double)v);
  }
}

// ---- K1: uniforms of consecutive substreams (parity hook) ----------------

struct UniformsParams {
  StreamStart st;
  int64_t n_substreams;
  int draws;
  double *out;  // [n_substreams][draws]
};

__global__ void __launch_bounds__(BLOCK) rng_uniforms_kernel(UniformsParams p) {
  unsigned g = blockIdx.x * BLOCK + threadIdx.x;
  Mrg32k3a sub = thread_start_state(p.st, g);
  int64_t stride = (int64_t)gridDim.x * BLOCK;
  for (int64_t j = g; j < p.n_substreams; j += stride) {
    Mrg32k3a cur = sub;
    for (int k = 0; k < p.draws; k++) p.out[j * p.draws + k] = cur.u01();
    sub.jump(p.st.stride);
  }
}

// ---- Mersenne-Twister stream generation ----------------------------------
// words[0..624) holds the current state; the kernels below write the words that follow it.  Seen as ONE sequence
// x[0], x[1], ... MT19937 is the recurrence
//     x[n] = x[n-227] ^ twist(x[n-624], x[n-623]),
// so 227 consecutive elements are independent of one another: each pass of mt_segment_kernel's loop produces 227 of
// them, reading only elements of earlier passes from a 1024-word circular buffer in shared memory (writes of a pass
// land at least 170 words away from anything it reads), with ONE barrier per pass.  That is all the parallelism the
// recurrence has by itself, and run as one block it used to be the bulk of the sequential-stream models' time.
//
// Long streams are therefore cut into segments of MT_SEGMENT words that run side by side.  The state a segment
// starts from — the 624 words x[pS .. pS+623] — comes from JUMP-AHEAD: from index 1 on the word sequence obeys a
// GF(2)-linear recurrence with a characteristic polynomial phi of degree 19937, so
//     x[pS + k] = XOR over the set bits i of g_p of x[1 + k + i],    g_p(t) = t^(pS-1) mod phi(t),
// a binary convolution of a precomputed polynomial (mt_jump.py writes the table at build time) with the first 20560
// words of the stream (mt_jump_kernel).  Every word is still exactly R's.
constexpr int64_t MT_SEGMENT = 65536;          // words per segment; must match mt_jump.py
constexpr int MT_POLY_WORDS = 624;             // 19937 coefficient bits
constexpr int MT_JUMP_INPUT = 1 + 19937 + 623;  // words [0, MT_JUMP_INPUT) must exist before a jump
constexpr int MT_JUMP_SPLIT = 8;               // blocks per jump, each taking a slice of the polynomial

// Block b generates segment p = b: [base + 624, end) from the state words[base .. base+623], where base = p * seg
// (block 0: seg0_base, the last 624 words generated so far) and end = (p + 1) * seg (last block: total).  seg is a
// multiple of MT_SEGMENT: the host picks it so that jumping and generating take about equally long.
__global__ void __launch_bounds__(256) mt_segment_kernel(uint32_t *words, int64_t total, int64_t seg0_base, int64_t seg) {
  __shared__ uint32_t x[1024];
  const int t = threadIdx.x;
  const int64_t p = blockIdx.x;
  const int64_t base = p ? p * seg : seg0_base;
  const int64_t end = (p == gridDim.x - 1) ? total : (p + 1) * seg;
  for (int k = t; k < 624; k += 256) x[(base + k) & 1023] = words[base + k];
  __syncthreads();
  for (int64_t n = base + 624; n < end; n += 227) {
    const int64_t i = n + t;
    if (t < 227 && i < end) {
      const uint32_t a = x[(i - 624) & 1023], b = x[(i - 623) & 1023], c = x[(i - 227) & 1023];
      const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
      const uint32_t v = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      x[i & 1023] = v;
      words[i] = v;
    }
    __syncthreads();
  }
}

// grid (segments - 1, MT_JUMP_SPLIT), 160 threads; segments are `mult` table segments long (polynomial p * mult - 1
// of the table).  Block (p - 1, q) XORs into words[p S + k], k < 624 (zeroed
// beforehand), the terms of polynomial words [q W, (q + 1) W), W = 624 / MT_JUMP_SPLIT: its window of the stream,
// x[1 + 32 q W .. + 32 W + 623], sits in shared memory.  Every thread walks the same set bits (no divergence) for FOUR
// outputs (k = t, t + 156, t + 312, t + 468), so the bit bookkeeping is shared and the four loads are independent.
__global__ void __launch_bounds__(160) mt_jump_kernel(uint32_t *words, const uint32_t *table, int mult) {
  constexpr int W = MT_POLY_WORDS / MT_JUMP_SPLIT;  // 78 polynomial words = 2496 coefficients per block
  constexpr int Q = 156;                            // 624 / 4
  __shared__ uint32_t win[32 * W + 624];
  __shared__ uint32_t g[W];
  const int t = threadIdx.x, q = blockIdx.y;
  const int64_t p = blockIdx.x + 1;
  const int first = 1 + 32 * q * W;  // stream index of win[0]
  for (int k = t; k < 32 * W + 624; k += 160) win[k] = (first + k < MT_JUMP_INPUT) ? words[first + k] : 0u;
  if (t < W) g[t] = table[(p * mult - 1) * MT_POLY_WORDS + q * W + t];
  __syncthreads();
  if (t < Q) {
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int w = 0; w < W; w++) {
      uint32_t bits = g[w];  // the same for every thread
      const uint32_t *row = win + 32 * w + t;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        a0 ^= row[b];
        a1 ^= row[b + Q];
        a2 ^= row[b + 2 * Q];
        a3 ^= row[b + 3 * Q];
      }
    }
    uint32_t *out = words + p * mult * MT_SEGMENT + t;
    if (a0) atomicXor(out, a0);
    if (a1) atomicXor(out + Q, a1);
    if (a2) atomicXor(out + 2 * Q, a2);
    if (a3) atomicXor(out + 3 * Q, a3);
  }
}

// MT parity hook: the uniforms R's unif_rand() would return
__global__ void mt_uniforms_kernel(const uint32_t *words, int64_t first_word, int64_t n, double *out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mt_word_to_unif(words[first_word + i]);
}

// ---- model traits ---------------------------------------------------------

struct PersonTraits {
  typedef person_r::Consts Consts;
  template <class Env> using Model = person_r::Person<Env>;
  static constexpr int MAX_ROWS = person_r::MAX_ROWS;
};
struct IllnessDeathTraits {
  typedef illness_death::Consts Consts;
  template <class Env> using Model = illness_death::SimplePerson<Env>;
  static constexpr int MAX_ROWS = illness_death::MAX_ROWS;
};
struct SimpleTraits {
  typedef simple_example::Consts Consts;
  template <class Env> using Model = simple_example::SimplePerson<Env>;
  static constexpr int MAX_ROWS = simple_example::MAX_ROWS;
};

// ---- cohort kernel: run n persons, emit row log / report / counts --------
//
// Row log.  Rows must come out in person order with a data-dependent number of
// rows per person.  The simulation kernel never synchronises across warps for
// this: each warp owns "warp tiles" of 32 consecutive persons, stages its rows
// in shared memory, packs them with a warp scan and stores them into a
// provisional area at a FIXED offset (tile * 32 * MAX_ROWS) together with the
// tile's row count.  A second, bandwidth-bound pass (rowlog_expand_kernel)
// scans the tile counts and expands the packed rows (start, end, meta = 20 B)
// into the final five f64 columns with coalesced stores.

template <class Consts>
struct CohortParams {
  // substream regime
  StreamStart st;
  // sequential-stream regime
  const uint32_t *mt_words;
  int64_t mt_end;         // number of valid words
  const int64_t *starts;  // word index of person i's first draw
  int64_t first_person, n;
  unsigned outputs;  // OUT_*
  Consts consts;
  // provisional packed rows: capacity 32*MAX_ROWS per warp tile
  double *prov_start, *prov_end;
  int *prov_meta;
  int *tile_count;  // rows per warp tile
  ReportGeom geom;
  double *dense;   // dense report accumulators (global, zeroed before launch)
  double *counts;  // N_COUNTS doubles
  int *error;      // first error code seen
};

template <class SrcT, int MAXR>
struct CohortEnv {
  typedef SrcT Src;
  Src src;
  unsigned outputs;
  // this thread's row slots in shared memory: slot k at [k*BLOCK]
  double *s_start, *s_end;
  int *s_meta;
  int nrows;
  const ReportGeom *geom;
  SharedAcc acc;
  int *s_kind;
  double sum_end;
  ReportCursor cursor;

  __device__ __forceinline__ void begin_person() { report_cursor_start(*geom, cursor); }
  __device__ __forceinline__ void row(double, double start, double end, int state, int kind) {
    if (outputs & OUT_COUNTS) {
      atomicAdd(&s_kind[kind], 1);
      sum_end += end;
    }
    if (outputs & OUT_ROWLOG) {
      if (nrows < MAXR) {
        s_start[nrows * BLOCK] = start;
        s_end[nrows * BLOCK] = end;
        s_meta[nrows * BLOCK] = (state << 8) | kind;
      }
      nrows++;
    }
  }
  __device__ __forceinline__ void report(int state, int kind, double lhs, double rhs) {
    if (outputs & OUT_REPORT) report_add(*geom, acc, cursor, state, kind, lhs, rhs);
  }
};

// sink-less environment: only the draws matter (seq_len_kernel)
template <class SrcT>
struct NullEnv {
  typedef SrcT Src;
  Src src;
  int code = 0;  // last set_uc() code (models with carried-over report state, sick_sicker)
  __device__ __forceinline__ void begin_person() {}
  __device__ __forceinline__ void row(double, double, double, int, int) {}
  __device__ __forceinline__ void report(int, int, double, double) {}
  __device__ __forceinline__ void summary(int, int, double, double) {}
  __device__ __forceinline__ void set_uc(int k, double, double) { code = k; }
  __device__ __forceinline__ void value(int, double) {}  // plugin.cuh
};

template <class Traits, bool STREAM>
__global__ void __launch_bounds__(BLOCK, 3) cohort_kernel(const __grid_constant__ CohortParams<typename Traits::Consts> p) {
  typedef typename std::conditional<STREAM, StreamSource, MrgSource>::type Src;
  typedef CohortEnv<Src, Traits::MAX_ROWS> Env;
  constexpr int MAXR = Traits::MAX_ROWS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // shared layout: [dense doubles | row slots start, end | row slots meta | dense ints]
  const int total = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  const int n_f = (p.outputs & OUT_REPORT) ? p.geom.off_dstart : 0;
  double *s_f = reinterpret_cast<double *>(smem_raw);
  constexpr int SLOTS = BLOCK * MAXR;
  double *slot_start = s_f + n_f;
  double *slot_end = slot_start + SLOTS;
  int *slot_meta = reinterpret_cast<int *>(slot_end + SLOTS);
  int *s_i = slot_meta + SLOTS;
  __shared__ int s_kind[8];
  __shared__ double s_sum[WARPS];
  __shared__ int s_error;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  Env env;
  env.outputs = p.outputs;
  env.geom = &p.geom;
  env.acc.f = s_f;
  env.acc.i = s_i;
  env.acc.n_f = n_f;
  env.s_kind = s_kind;
  env.sum_end = 0.0;
  env.s_start = slot_start + tid;
  env.s_end = slot_end + tid;
  env.s_meta = slot_meta + tid;
  if (p.outputs & OUT_REPORT) shared_acc_clear(env.acc, total);
  if (tid < 8) s_kind[tid] = 0;
  if (tid == 0) s_error = 0;
  __syncthreads();

  Mrg32k3a sub;
  if constexpr (!STREAM) sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t n_tiles = (p.n + BLOCK - 1) / BLOCK;  // block tiles; warp tile = tile*WARPS + warp
  int err = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t idx = tile * BLOCK + tid;
    env.nrows = 0;
    if (idx < p.n) {
      if constexpr (STREAM) {
        env.src.w = p.mt_words;
        env.src.pos = p.starts[idx];
        env.src.end = p.mt_end;
        env.src.overrun = 0;
      } else {
        env.src.g = sub;
      }
      typename Traits::template Model<Env> person(env, p.consts, p.first_person + idx);
      person.run();
      if (person.error_) err = person.error_;
      if (env.nrows > MAXR) err = -4;
      if constexpr (STREAM) {
        if (env.src.overrun) err = -6;
      }
    }
    __syncwarp();  // persons end after different numbers of events: re-form the warp before the jump and the next tile
    if constexpr (!STREAM) sub.jump(p.st.stride);

    if (p.outputs & OUT_ROWLOG) {
      // pack this warp's rows: exclusive scan of the per-person row counts
      const int cnt = env.nrows > MAXR ? MAXR : env.nrows;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      const int off = incl - cnt;
      const int64_t wt = tile * WARPS + warp;
      const int64_t base = wt * (32 * MAXR);
      for (int r = 0; r < cnt; r++) {
        p.prov_start[base + off + r] = env.s_start[r * BLOCK];
        p.prov_end[base + off + r] = env.s_end[r * BLOCK];
        p.prov_meta[base + off + r] = (lane << 16) | env.s_meta[r * BLOCK];
      }
      if (lane == 0) p.tile_count[wt] = tot;
    }
  }

  if (err) atomicCAS(&s_error, 0, err);
  __syncthreads();
  if (p.outputs & OUT_REPORT) shared_acc_flush(env.acc, total, p.dense);
  if (p.outputs & OUT_COUNTS) {
    double v = env.sum_end;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_sum[warp] = v;
    __syncthreads();
    if (tid < 8 && s_kind[tid]) atomicAdd(&p.counts[tid], (double)s_kind[tid]);
    if (tid == 0) {
      double t = 0;
      for (int w = 0; w < WARPS; w++) t += s_sum[w];
      atomicAdd(&p.counts[8], t);
    }
  }
  if (tid == 0 && s_error) atomicCAS(p.error, 0, s_error);
}

template <class Traits>
inline size_t cohort_kernel_smem(const ReportGeom &g, unsigned outputs) {
  size_t b = 0;
  if (outputs & OUT_REPORT) b += (size_t)g.off_dstart * 8 + (size_t)(g.total - g.off_dstart) * 4;
  b += (size_t)BLOCK * Traits::MAX_ROWS * (8 + 8 + 4);
  return b;
}

// ---- row log, second pass: tile counts -> offsets, packed rows -> columns --

constexpr int EXPAND_TILES = 256;  // warp tiles (8192 persons) per block of the expand pass: 122 blocks for 1e6 persons

struct ExpandParams {
  const double *prov_start, *prov_end;
  const int *prov_meta;
  const int *tile_count;
  int64_t n_warp_tiles;
  int tile_cap;                  // 32 * MAX_ROWS
  unsigned long long *partial;   // per-chunk row totals, then (in place) exclusive chunk offsets
  unsigned long long *n_rows;
  int64_t n_chunks;
  int64_t first_person;
  double *rowlog;   // 5 columns (end time, event, id, startTime, state), column stride `capacity`
  int64_t capacity;
};

__global__ void __launch_bounds__(256) rowlog_partial_kernel(ExpandParams p) {
  __shared__ unsigned long long s_w[8];
  const int64_t lo = (int64_t)blockIdx.x * EXPAND_TILES;
  unsigned long long v = 0;
  for (int k = threadIdx.x; k < EXPAND_TILES; k += 256) {
    int64_t t = lo + k;
    if (t < p.n_warp_tiles) v += (unsigned long long)p.tile_count[t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; w++) t += s_w[w];
    p.partial[blockIdx.x] = t;
  }
}

// one block: exclusive scan of the chunk totals in place; total row count out
__global__ void __launch_bounds__(1024) rowlog_chunk_scan_kernel(ExpandParams p) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t c0 = 0; c0 < p.n_chunks; c0 += 1024) {
    int64_t c = c0 + tid;
    unsigned long long v = (c < p.n_chunks) ? p.partial[c] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    unsigned long long carry = s_carry;
    if (c < p.n_chunks) p.partial[c] = carry + woff + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (tid == 0) *p.n_rows = s_carry;
}

__global__ void __launch_bounds__(256) rowlog_expand_kernel(ExpandParams p) {
  __shared__ unsigned long long s_off[EXPAND_TILES];  // exclusive row offset of each tile of this chunk
  __shared__ unsigned long long s_w[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t lo = (int64_t)blockIdx.x * EXPAND_TILES;
  // block exclusive scan over the chunk's tile counts: thread t owns tiles [8t, 8t+8)
  constexpr int PER = EXPAND_TILES / 256;
  int c[PER];
  unsigned long long v = 0;
#pragma unroll
  for (int k = 0; k < PER; k++) {
    int64_t t = lo + tid * PER + k;
    c[k] = (t < p.n_warp_tiles) ? p.tile_count[t] : 0;
    v += (unsigned long long)c[k];
  }
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned long long run = p.partial[blockIdx.x] + incl - v;
  for (int w = 0; w < warp; w++) run += s_w[w];
#pragma unroll
  for (int k = 0; k < PER; k++) {
    s_off[tid * PER + k] = run;
    run += (unsigned long long)c[k];
  }
  __syncthreads();
  // each warp expands tiles warp, warp+8, ... of the chunk
  for (int k = warp; k < EXPAND_TILES; k += 8) {
    const int64_t t = lo + k;
    if (t >= p.n_warp_tiles) break;
    const int cnt = p.tile_count[t];
    const unsigned long long off = s_off[k];
    const int64_t src = t * p.tile_cap;
    const double id0 = (double)(p.first_person + t * 32);
    for (int r = lane; r < cnt; r += 32) {
      unsigned long long g = off + r;
      if (g < (unsigned long long)p.capacity) {
        int m = p.prov_meta[src + r];
        p.rowlog[0 * p.capacity + g] = p.prov_end[src + r];
        p.rowlog[1 * p.capacity + g] = (double)(m & 0xff);
        p.rowlog[2 * p.capacity + g] = id0 + (double)(m >> 16);
        p.rowlog[3 * p.capacity + g] = p.prov_start[src + r];
        p.rowlog[4 * p.capacity + g] = (double)((m >> 8) & 0xff);
      }
    }
  }
}

// ---- sequential-stream regime: draws per start position, chain resolution --

// len[p] = number of uniforms a person whose first draw is word first_word+p
// would consume (255 = ran past the generated words)
template <class Traits>
__global__ void __launch_bounds__(BLOCK) seq_len_kernel(const uint32_t *words, int64_t mt_end, int64_t first_word,
                                                        int64_t n_pos, typename Traits::Consts consts,
                                                        unsigned char *len, unsigned char *aux) {
  typedef NullEnv<StreamSource> Env;
  for (int64_t p0 = (int64_t)blockIdx.x * BLOCK + (threadIdx.x & ~31); p0 < n_pos; p0 += (int64_t)gridDim.x * BLOCK) {
    const int64_t p = p0 + (threadIdx.x & 31);
    __syncwarp();  // every pass starts with the warp together
    if (p >= n_pos) continue;
    Env env;
    env.src.w = words;
    env.src.pos = first_word + p;
    env.src.end = mt_end;
    env.src.overrun = 0;
    typename Traits::template Model<Env> person(env, consts, 0);
    person.run();
    int64_t used = env.src.pos - (first_word + p);
    len[p] = (env.src.overrun || used > 250 || person.error_) ? 255 : (unsigned char)used;
    if (aux) aux[p] = (unsigned char)env.code;
  }
}

// For chunk c and entry offset e: walk p += len[p] from c*CHUNK+e to the chunk
// end; record where the chain enters the next chunk and how many persons
// started inside this one.
__global__ void __launch_bounds__(256) chain_chunk_kernel(const unsigned char *len, int64_t n_pos, int max_draws,
                                                          int stride, unsigned char *exit_off, int *count) {
  __shared__ unsigned char s_len[CHAIN_CHUNK];
  const int64_t c = blockIdx.x;
  const int64_t lo = c * CHAIN_CHUNK;
  for (int k = threadIdx.x; k < CHAIN_CHUNK; k += blockDim.x) s_len[k] = (lo + k < n_pos) ? len[lo + k] : 255;
  __syncthreads();
  int e = threadIdx.x;
  if (e < max_draws) {
    int cnt = 0;
    int ex = chain_walk(s_len, e, &cnt);
    exit_off[c * stride + e] = (unsigned char)ex;
    count[c * stride + e] = cnt;
  }
}

// Pass over chunks: entry offset and first person index of every chunk (entry = 255 marks "chain broken / not
// reached").  The dependency chain is sequential — chunk c+1 is entered where chunk c is left — but every link is a
// table lookup, so the one block does it in two levels: tables are staged in shared memory; groups of SCAN_GROUP
// chunks are composed for EVERY entry offset in parallel; one thread walks the groups; the groups are expanded in
// parallel with the exact per-chunk rule (chain_scan_step).  Narrow tables (models with <= 8 draws per person) take
// 512 chunks per batch; wide tables (stride 256) 32 chunks per batch and are walked chunk by chunk in shared memory.
constexpr int SCAN_BATCH = 512;
constexpr int SCAN_GROUP = 8;
constexpr int SCAN_WIDE_BATCH = 32;
__global__ void __launch_bounds__(256) chain_scan_kernel(const unsigned char *exit_off, const int *count,
                                                         int64_t n_chunks, int64_t n, int stride, unsigned char *entry,
                                                         int64_t *base) {
  // narrow: 512 x 8 exits (4 KB) + counts (16 KB); wide: 32 x 256 exits (8 KB) + counts (32 KB, the larger)
  __shared__ unsigned char s_exit[SCAN_WIDE_BATCH * 256];
  __shared__ int s_count[SCAN_WIDE_BATCH * 256];
  __shared__ unsigned char s_entry[SCAN_BATCH];
  __shared__ int64_t s_base[SCAN_BATCH];
  __shared__ unsigned char g_exit[(SCAN_BATCH / SCAN_GROUP) * CHAIN_MAX_DRAWS];
  __shared__ int g_count[(SCAN_BATCH / SCAN_GROUP) * CHAIN_MAX_DRAWS];
  __shared__ unsigned char g_entry[SCAN_BATCH / SCAN_GROUP];
  __shared__ int64_t g_base[SCAN_BATCH / SCAN_GROUP];
  __shared__ int s_e;
  __shared__ int64_t s_b;
  static_assert(SCAN_BATCH * CHAIN_MAX_DRAWS <= SCAN_WIDE_BATCH * 256, "shared tables are sized by the wide layout");
  if (threadIdx.x == 0) {
    s_e = 0;
    s_b = 0;
  }
  if (stride > CHAIN_MAX_DRAWS) {
    // wide tables: stage 32 chunks at a time, thread 0 follows the chain in shared memory
    for (int64_t c0 = 0; c0 < n_chunks; c0 += SCAN_WIDE_BATCH) {
      const int m = (int)((n_chunks - c0 < SCAN_WIDE_BATCH) ? (n_chunks - c0) : SCAN_WIDE_BATCH);
      __syncthreads();
      for (int k = threadIdx.x; k < m * stride; k += blockDim.x) {
        s_exit[k] = exit_off[c0 * stride + k];
        s_count[k] = count[c0 * stride + k];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int e = s_e;
        int64_t b = s_b;
        for (int c = 0; c < m; c++) {
          s_entry[c] = (unsigned char)e;
          s_base[c] = b;
          chain_scan_step(&s_exit[c * stride], &s_count[c * stride], n, &e, &b);
        }
        s_e = e;
        s_b = b;
      }
      __syncthreads();
      for (int k = threadIdx.x; k < m; k += blockDim.x) {
        entry[c0 + k] = s_entry[k];
        base[c0 + k] = s_base[k];
      }
    }
    return;
  }
  for (int64_t c0 = 0; c0 < n_chunks; c0 += SCAN_BATCH) {
    const int m = (int)((n_chunks - c0 < SCAN_BATCH) ? (n_chunks - c0) : SCAN_BATCH);
    const int n_groups = (m + SCAN_GROUP - 1) / SCAN_GROUP;
    __syncthreads();
    for (int k = threadIdx.x; k < m * CHAIN_MAX_DRAWS; k += blockDim.x) {
      s_exit[k] = exit_off[c0 * CHAIN_MAX_DRAWS + k];
      s_count[k] = count[c0 * CHAIN_MAX_DRAWS + k];
    }
    __syncthreads();
    // every group, from every entry offset: where the chain leaves it and how many persons start inside
    for (int k = threadIdx.x; k < n_groups * CHAIN_MAX_DRAWS; k += blockDim.x) {
      const int grp = k / CHAIN_MAX_DRAWS;
      int e = k % CHAIN_MAX_DRAWS, cnt = 0;
      const int c_end = min(m, (grp + 1) * SCAN_GROUP);
      for (int c = grp * SCAN_GROUP; c < c_end && e != 255; c++) {
        cnt += s_count[c * CHAIN_MAX_DRAWS + e];
        e = s_exit[c * CHAIN_MAX_DRAWS + e];
        if (e != 255 && e >= CHAIN_MAX_DRAWS) e = 255;  // cannot happen for a model within its draw bound
      }
      g_exit[k] = (unsigned char)e;
      g_count[k] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int e = s_e;
      int64_t b = s_b;
      for (int grp = 0; grp < n_groups; grp++) {
        g_entry[grp] = (unsigned char)e;
        g_base[grp] = b;
        // past person n, or broken: everything behind is 255 (the exact rule is applied chunk by chunk below)
        chain_scan_step(&g_exit[grp * CHAIN_MAX_DRAWS], &g_count[grp * CHAIN_MAX_DRAWS], n, &e, &b);
      }
      s_e = e;
      s_b = b;
    }
    __syncthreads();
    for (int grp = threadIdx.x; grp < n_groups; grp += blockDim.x) {
      int e = g_entry[grp];
      int64_t b = g_base[grp];
      const int c_end = min(m, (grp + 1) * SCAN_GROUP);
      for (int c = grp * SCAN_GROUP; c < c_end; c++) {
        s_entry[c] = (unsigned char)e;
        s_base[c] = b;
        chain_scan_step(&s_exit[c * CHAIN_MAX_DRAWS], &s_count[c * CHAIN_MAX_DRAWS], n, &e, &b);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
      entry[c0 + k] = s_entry[k];
      base[c0 + k] = s_base[k];
    }
  }
}

// Wide tables, many chunks: one more level across blocks.  chain_group_kernel composes CHAIN_GROUP consecutive chunks
// for every entry offset (one block per group, one thread per offset), chain_scan_kernel then runs over the GROUPS, and
// chain_expand_kernel walks each group's chunks from the group's resolved entry with the exact per-chunk rule.
constexpr int CHAIN_GROUP = 32;
__global__ void __launch_bounds__(256) chain_group_kernel(const unsigned char *exit_off, const int *count, int64_t n_chunks,
                                                          int stride, unsigned char *g_exit, int *g_count) {
  const int64_t grp = blockIdx.x;
  int e = threadIdx.x, cnt = 0;
  if (e >= stride) return;
  const int64_t c_end = min(n_chunks, (grp + 1) * CHAIN_GROUP);
  for (int64_t c = grp * CHAIN_GROUP; c < c_end && e != 255; c++) {
    cnt += count[c * stride + e];
    e = exit_off[c * stride + e];
  }
  g_exit[grp * stride + threadIdx.x] = (unsigned char)e;
  g_count[grp * stride + threadIdx.x] = cnt;
}
__global__ void __launch_bounds__(64) chain_expand_kernel(const unsigned char *exit_off, const int *count, int64_t n_chunks,
                                                          int64_t n_groups, int64_t n, int stride, const unsigned char *g_entry,
                                                          const int64_t *g_base, unsigned char *entry, int64_t *base) {
  const int64_t grp = (int64_t)blockIdx.x * 64 + threadIdx.x;
  if (grp >= n_groups) return;
  int e = g_entry[grp];
  int64_t b = g_base[grp];
  const int64_t c_end = min(n_chunks, (grp + 1) * CHAIN_GROUP);
  for (int64_t c = grp * CHAIN_GROUP; c < c_end; c++) {
    entry[c] = (unsigned char)e;
    base[c] = b;
    chain_scan_step(exit_off + c * stride, count + c * stride, n, &e, &b);
  }
}

// Second walk from the resolved entry offsets: starts[i] (as word indices) for
// i < n, and the word index one past the last draw of person n-1.
__global__ void __launch_bounds__(64) chain_emit_kernel(const unsigned char *len, int64_t n_pos, int64_t first_word,
                                                        const unsigned char *entry, const int64_t *base, int64_t n,
                                                        int64_t *starts, int64_t *end_word, int *error) {
  __shared__ unsigned char s_len[CHAIN_CHUNK];
  const int64_t c = blockIdx.x;
  const int64_t lo = c * CHAIN_CHUNK;
  int e = entry[c];
  int64_t b = base[c];
  if (e == 255 || b >= n) return;  // block-uniform
  for (int k = threadIdx.x; k < CHAIN_CHUNK; k += blockDim.x) s_len[k] = (lo + k < n_pos) ? len[lo + k] : 255;
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (!chain_emit(s_len, e, b, n, first_word + lo, starts, end_word)) atomicCAS(error, 0, -6);
}

// ---- K3: calibration sweep ------------------------------------------------

struct CalibParams {
  StreamStart st;  // stride = M^(gridDim.x*BLOCK)
  int64_t n;
  int n_sets;
  const double *runpar;  // [n_sets][6]
  double *out;           // [n_sets][48]: occupancy[4][10], TimeAtRisk[4], then for k < 4 how many blocks added to TimeAtRisk[k]
  int set0;              // blockIdx.y = 0 is parameter set set0 (a launch may cover a range of the sets)
  int *error;
  int *tie;              // set if two pending actions of a person compared equal (ORDERED only)
  double *partial;       // [n_sets][gridDim.x][4]: every block's TimeAtRisk sums, added up in block order afterwards
};

struct CalibEnv {
  typedef MrgSource Src;
  Src src;
  int *s_occ;     // [4][10]
  double *heap_t;  // this thread's event-heap slots in shared memory: slot k at heap_t[k * heap_stride]
  int *heap_m;
  int heap_stride;
  double tar[4];  // per-thread TimeAtRisk partial sums
  int tar_n;      // 1 + largest TimeAtRisk index this thread touched
  __device__ __forceinline__ void occupancy(int stage, int cind) {
    if (stage < 4) atomicAdd(&s_occ[stage * 10 + cind], 1);
  }
  double died_clin;  // clinTime of the person that died in the event loop just run, < 0 if nobody did
  __device__ __forceinline__ void at_death(double clinTime) { died_clin = clinTime; }
  // after the event loop, all lanes together
  __device__ __forceinline__ void settle() {
    if (died_clin >= 0.0) {
      const double c = died_clin;
      tar[0] += fmin(20.0, c);
      int reached = 1;
      if (!(c < 20.0)) {
        tar[1] += fmin(40.0, c);
        reached = 2;
        if (!(c < 40.0)) {
          tar[2] += fmin(60.0, c);
          reached = 3;
          if (!(c < 60.0)) {
            tar[3] += fmin(80.0, c);
            reached = 4;
          }
        }
      }
      if (reached > tar_n) tar_n = reached;
    }
  }
};

// grid = (blocks_x, n_sets): every parameter set sees the same substreams
// (common random numbers, R/calibSim.R:12-14 with one seed for all sets).
// ORDERED: the pending actions as an ordered array (engine.cuh, OrderedEvents); results are valid unless *p.tie is set
template <bool ORDERED>
__global__ void __launch_bounds__(BLOCK) calib_kernel(CalibParams p) {
  __shared__ int s_occ[40];
  __shared__ double s_heap_t[calib::HEAP_SLOTS * BLOCK];  // 28 KB + 14 KB: every thread's heap slots
  __shared__ int s_heap_m[calib::HEAP_SLOTS * BLOCK];
  __shared__ double s_tar[WARPS][4];
  __shared__ int s_tar_n;
  __shared__ int s_error;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int set = p.set0 + blockIdx.y;
  if (tid < 40) s_occ[tid] = 0;
  if (tid == 0) {
    s_tar_n = 0;
    s_error = 0;
  }
  __syncthreads();

  calib::Par par;
  const double *rp = p.runpar + 6 * set;
  par.Lam1 = rp[0];
  par.sigm1 = rp[1];
  par.p2 = rp[2];
  par.lam2 = rp[3];
  par.mu3 = rp[4];
  par.tau3 = rp[5];

  CalibEnv env;
  env.s_occ = s_occ;
  env.heap_t = s_heap_t + tid;
  env.heap_m = s_heap_m + tid;
  env.heap_stride = BLOCK;
  env.tar[0] = env.tar[1] = env.tar[2] = env.tar[3] = 0.0;
  env.tar_n = 0;

  Mrg32k3a sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  // the trip count is the warp's (lane 0 has the smallest index), and every pass starts with the warp together:
  // without the __syncwarp, lanes whose person ends early run on into the substream jump and the next person on
  // their own and the warp never re-forms (6 of 32 lanes active, measured)
  const int64_t warp_first = (int64_t)blockIdx.x * BLOCK + (tid & ~31);
  bool tie = false;
  for (int64_t base = warp_first; base < p.n; base += stride) {
    const int64_t idx = base + lane;
    __syncwarp();
    if (idx < p.n) {
      env.src.g = sub;
      env.died_clin = -1.0;
      calib::CalibPerson<CalibEnv, ORDERED> person(env, par);
      person.run();
      if (person.error_) atomicCAS(&s_error, 0, person.error_);
      if (person.actions.tie) tie = true;
      env.settle();  // = calib::time_at_risk_terms(), written out
    }
    __syncwarp();
    sub.jump(p.st.stride);
  }

#pragma unroll
  for (int k = 0; k < 4; k++) {
    double v = env.tar[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_tar[warp][k] = v;
  }
  atomicMax(&s_tar_n, env.tar_n);
  __syncthreads();
  double *out = p.out + 48 * set;
  if (tid < 40 && s_occ[tid]) atomicAdd(&out[tid], (double)s_occ[tid]);
  if (tid < 4) {
    double t = 0;
    for (int w = 0; w < WARPS; w++) t += s_tar[w][tid];
    p.partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 + tid] = t;  // no atomics: the sum must not depend on timing
    // TimeAtRisk has as many entries as the largest index any person reached (calibperson-r.cc:124-135); kept as
    // counts per index so that shards merge by addition like everything else in the vector
    if (tid < s_tar_n) atomicAdd(&out[44 + tid], 1.0);
  }
  if (tid == 0 && s_error) atomicCAS(p.error, 0, s_error);
  if (ORDERED && tie) *p.tie = 1;
}

// TimeAtRisk of every set: the blocks' sums in block order (one block per set, threads 0..3)
__global__ void calib_fold_kernel(const double *partial, int blocks, int set0, double *out) {
  if (threadIdx.x < 4) {
    const double *q = partial + (size_t)blockIdx.x * blocks * 4 + threadIdx.x;
    double t = 0.0;
    for (int b = 0; b < blocks; b++) t += q[4 * b];
    out[48 * (size_t)(set0 + blockIdx.x) + 40 + threadIdx.x] = t;
  }
}

// ---- survival-time utilities and lookup tables: one evaluation per thread --
struct LookupParams {
  int family, mode;  // family 0 Rpexp, 1 Rpexp_gamma, 2 NumericInterpolate, 3 Table; modes as in include/msim.h
  Rpexp pexp;
  RpexpGamma pgam;
  Interp interp;
  DenseTable table;
  int64_t n;
  const double *a, *b, *c;  // inputs (b, c may be null: 0); for tables a = keys [n][n_axes]
  double *out;
  int *error;
};
__global__ void __launch_bounds__(256) lookup_kernel(LookupParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    double r = 0.0;
    if (p.family == 0) {
      const double x = p.a[i];
      if (p.mode == 0) {
        int failed = 0;
        r = rpexp_rand(p.pexp, x, p.b ? p.b[i] : 0.0, &failed);
        if (failed) atomicCAS(p.error, 0, -3);
      } else {
        r = p.mode == 1 ? rpexp_H(p.pexp, x) : p.mode == 2 ? rpexp_h(p.pexp, x) : p.mode == 3 ? rpexp_p(p.pexp, x, true) : rpexp_d(p.pexp, x);
      }
    } else if (p.family == 1) {
      const double x = p.a[i], Z = p.b[i];
      r = p.mode == 0 ? rpexp_gamma_rand(p.pgam, x, Z, p.c ? p.c[i] : 0.0)
          : p.mode == 1 ? rpexp_gamma_H(p.pgam, x, Z) : p.mode == 2 ? rpexp_gamma_h(p.pgam, x, Z) : rpexp_gamma_p(p.pgam, x, Z, true);
    } else if (p.family == 2) {
      const double x = p.a[i];
      r = p.mode == 0 ? interp_approx(p.interp, x) : p.mode == 1 ? interp_step(p.interp, x)
          : p.mode == 2 ? interp_invert(p.interp, x) : interp_invert_decreasing(p.interp, x);
    } else {
      r = table_lookup(p.table, p.a + i * p.table.n_axes);
    }
    p.out[i] = r;
  }
}

// ---- generalised survival model: one inversion per thread ----------------
// mode 0: randU(u, tentry, index), 1: randU0(u, index), 2: eta(y = u, index)
__global__ void __launch_bounds__(128) gsm_kernel(const GsmModel *M_global, int mode, int64_t n, const double *u, const double *tentry,
                                                  const int *index, double scale, double *out) {
  // the model (knots, projection matrices, coefficients: ~10 KB) is read at every spline evaluation: keep it on chip
  __shared__ GsmModel M;
  {
    const int *src = reinterpret_cast<const int *>(M_global);
    int *dst = reinterpret_cast<int *>(&M);
    for (int k = threadIdx.x; k < (int)(sizeof(GsmModel) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (mode == 2) {  // eta only: one evaluation each
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      out[i] = gsm_eta(M, index ? index[i] : 0, u[i], false);
    return;
  }
  // randU / randU0.  A search takes 2 + (4 to 40) evaluations of the spline, so with one search per lane per pass a
  // warp waits for its slowest search (5.6 threads per warp instruction, profiles/r01c_secondary_kernels.md).  Here
  // every lane carries the state of ITS search (gsm.cuh, Brent: the state machine asks for a function value) and
  // every pass of the loop is ONE spline evaluation on every lane, whatever stage the lane's search is in — the
  // target's own evaluations (tentry, randU0's t0) included; a lane whose search has ended starts its next one.
  int64_t next = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, cur = 0;
  Brent s;
  double target = 0.0, target0 = 0.0, y = 0.0, uu = 0.0, e_first = 0.0, ymin = 0.0, ymax = 0.0;
  int idx = 0, stage = 0;  // 0: idle, 1: eta(ymin) for randU's target, 2/3: eta(y0, true/false) for randU0's, 4: searching
  bool base0 = false;
  for (;;) {
    __syncwarp();
    if (stage == 0 && next < n) {
      cur = next;
      next += stride;
      idx = index ? index[cur] : 0;
      uu = u[cur];
      const double te = (mode == 0 && tentry) ? tentry[cur] : 0.0;
      ymin = te == 0.0 ? (M.log_time ? log(M.tmin / scale) : M.tmin / scale) : (M.log_time ? log(te) : te);
      ymax = M.log_time ? log(M.tmax * scale) : M.tmax * scale;
      target0 = 0.0;
      if (mode == 0 && te != 0.0) {  // gsm::randU with an entry time (gsm.cpp:40-43)
        stage = 1;
        y = ymin;
        base0 = false;
      } else if (mode == 1) {  // gsm::randU0 (gsm.cpp:52-55)
        stage = 2;
        y = M.log_time ? log(M.t0) : M.t0;
        base0 = true;
      } else {
        target = gsm_link(uu);
        brent_begin(s, ymin, ymax, 1.0e-8, 100);
        stage = 4;
      }
    }
    if (__ballot_sync(0xffffffffu, stage != 0) == 0) break;
    if (stage != 0) {
      if (stage == 4) {  // the search's function (gsm.cpp:29-31): which baseline depends on where it is evaluated
        y = s.x;
        base0 = !((M.log_time ? exp(y) : y) < M.t0);
      }
      const double e = gsm_eta(M, idx, y, base0);
      if (stage == 4) {
        brent_feed(s, e - (base0 ? target0 : target));
        if (s.done) {
          out[cur] = M.log_time ? exp(s.root) : s.root;
          stage = 0;
        }
      } else if (stage == 1) {
        target = gsm_link(uu * gsm_linkinv(e));
        brent_begin(s, ymin, ymax, 1.0e-8, 100);
        stage = 4;
      } else if (stage == 2) {
        e_first = e;  // eta(y0, true)
        base0 = false;
        stage = 3;
      } else {  // stage 3: eta(y0, false)
        target = gsm_link(uu);
        target0 = gsm_link(uu * gsm_linkinv(e_first) / gsm_linkinv(e));
        brent_begin(s, ymin, ymax, 1.0e-8, 100);
        stage = 4;
      }
    }
  }
}

}  // namespace msim
