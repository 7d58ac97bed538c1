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
//     data-dependent number of rows per person: each tile does a block scan of
//     its row counts and obtains its global row offset with a decoupled
//     look-back over per-tile status words (single pass, no second simulation);
//     rows are compacted in shared memory and stored with coalesced, linear
//     writes, column by column;
//   * report tables and per-kind counts accumulate in shared memory and are
//     flushed once per block with `red.global.add.f64`.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "models.cuh"
#include "report.cuh"
#include "seqstream.cuh"

namespace msim {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr unsigned OUT_ROWLOG = 1u, OUT_REPORT = 2u, OUT_COUNTS = 4u;
constexpr int N_COUNTS = 9;  // up to 8 per-kind counts + sum(endTime)

// ---- small device helpers ------------------------------------------------

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// dense accumulators in shared memory: [0, n_f) doubles, [n_f, total) ints
struct SharedAcc {
  double *f;
  int *i;
  int n_f;
  __device__ __forceinline__ void addf(int idx, double v) { atomicAdd(&f[idx], v); }
  __device__ __forceinline__ void addi(int idx, int v) { atomicAdd(&i[idx - n_f], v); }
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
// One block.  words[0..624) holds the current state; writes n_batches further
// regenerations behind it (three barrier-separated waves each, seqstream.cuh).
__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t *words, int n_batches) {
  __shared__ uint32_t mt[624];
  const int t = threadIdx.x;
  for (int k = t; k < 624; k += 256) mt[k] = words[k];
  __syncthreads();
  for (int b = 1; b <= n_batches; b++) {
#pragma unroll
    for (int wave = 0; wave < 3; wave++) {
      int kk = 0;
      uint32_t v = 0;
      bool act = mt_wave_value(mt, wave, t, &kk, &v);
      __syncthreads();
      if (act) mt[kk] = v;
      __syncthreads();
    }
    uint32_t *dst = words + (size_t)b * 624;
    for (int k = t; k < 624; k += 256) dst[k] = mt[k];
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

template <class Consts>
struct CohortParams {
  // substream regime
  StreamStart st;
  // sequential-stream regime
  const uint32_t *mt_words;
  int64_t mt_end;           // number of valid words
  const int64_t *starts;    // word index of person i's first draw
  int64_t first_person, n;
  unsigned outputs;  // OUT_*
  Consts consts;
  // row log: 5 columns (end time, event, id, startTime, state), stride `capacity`
  double *rowlog;
  int64_t capacity;
  unsigned long long *tile_status;  // zeroed before launch
  unsigned long long *n_rows;       // total rows (may exceed capacity)
  ReportGeom geom;
  double *dense;   // dense report accumulators (global, zeroed before launch)
  double *counts;  // N_COUNTS doubles
  int *error;      // first error code seen
};

constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_MASK = (1ull << 62) - 1;

template <class SrcT, int MAXR>
struct CohortEnv {
  typedef SrcT Src;
  Src src;
  unsigned outputs;
  // per-thread row scratch (at most MAXR rows per person)
  double r_start[MAXR], r_end[MAXR];
  int r_meta[MAXR];
  int nrows;
  const ReportGeom *geom;
  SharedAcc acc;
  int *s_kind;
  double sum_end;

  __device__ __forceinline__ void row(double, double start, double end, int state, int kind) {
    if (outputs & OUT_COUNTS) {
      atomicAdd(&s_kind[kind], 1);
      sum_end += end;
    }
    if (outputs & OUT_ROWLOG) {
      if (nrows < MAXR) {
        r_start[nrows] = start;
        r_end[nrows] = end;
        r_meta[nrows] = (state << 8) | kind;
      }
      nrows++;
    }
  }
  __device__ __forceinline__ void report(int state, int kind, double lhs, double rhs) {
    if (outputs & OUT_REPORT) report_add(*geom, acc, state, kind, lhs, rhs);
  }
};

// sink-less environment: only the draws matter (seq_len_kernel)
template <class SrcT>
struct NullEnv {
  typedef SrcT Src;
  Src src;
  __device__ __forceinline__ void row(double, double, double, int, int) {}
  __device__ __forceinline__ void report(int, int, double, double) {}
};

// Exclusive offset of this tile in the global row log (decoupled look-back).
// Called by warp 0 only; `total` is the tile's own row count.
__device__ __forceinline__ unsigned long long tile_lookback(unsigned long long *status, int64_t tile,
                                                            unsigned long long total) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(&status[0], ST_PREFIX | total);
    return 0;
  }
  if (lane == 0) st_relaxed_u64(&status[tile], ST_AGG | total);
  unsigned long long excl = 0;
  int64_t j = tile - 1;  // lane L inspects tile j - L
  for (;;) {
    int64_t mine = j - lane;
    unsigned long long v = ST_PREFIX;  // tiles before 0 act as an empty prefix
    if (mine >= 0) {
      do {
        v = ld_relaxed_u64(&status[mine]);
      } while ((v >> 62) == 0);
    }
    unsigned has_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    int first = has_prefix ? (__ffs(has_prefix) - 1) : 32;  // nearest predecessor holding a full prefix
    unsigned long long contrib = (lane <= first) ? (v & ST_MASK) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    excl += contrib;
    if (has_prefix) break;
    j -= 32;
  }
  if (lane == 0) st_relaxed_u64(&status[tile], ST_PREFIX | (excl + total));
  return excl;
}

template <class Traits, bool STREAM>
__global__ void __launch_bounds__(BLOCK, 2) cohort_kernel(CohortParams<typename Traits::Consts> p) {
  typedef typename std::conditional<STREAM, StreamSource, MrgSource>::type Src;
  typedef CohortEnv<Src, Traits::MAX_ROWS> Env;
  constexpr int MAXR = Traits::MAX_ROWS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // shared layout: [dense doubles | compact rows (start, end, meta) | dense ints]
  const int total = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  const int n_f = (p.outputs & OUT_REPORT) ? p.geom.off_touch : 0;
  double *s_f = reinterpret_cast<double *>(smem_raw);
  constexpr int CROWS = BLOCK * MAXR;
  double *c_start = s_f + n_f;
  double *c_end = c_start + CROWS;
  int *c_meta = reinterpret_cast<int *>(c_end + CROWS);
  int *s_i = c_meta + CROWS;
  __shared__ int s_kind[8];
  __shared__ double s_sum[WARPS];
  __shared__ int s_warp_tot[WARPS];
  __shared__ unsigned long long s_tile_base;
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
  if (p.outputs & OUT_REPORT) shared_acc_clear(env.acc, total);
  if (tid < 8) s_kind[tid] = 0;
  if (tid == 0) s_error = 0;
  __syncthreads();

  Mrg32k3a sub;
  if constexpr (!STREAM) sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t n_tiles = (p.n + BLOCK - 1) / BLOCK;

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
      int err = person.error_;
      if (env.nrows > MAXR) err = -4;
      if constexpr (STREAM) {
        if (env.src.overrun) err = -6;
      }
      if (err) atomicCAS(&s_error, 0, err);
    }
    if constexpr (!STREAM) sub.jump(p.st.stride);

    if (p.outputs & OUT_ROWLOG) {
      // block exclusive scan of the per-person row counts
      int cnt = env.nrows > MAXR ? MAXR : env.nrows;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_warp_tot[warp] = incl;
      __syncthreads();  // also orders the previous tile's reads of c_* before this tile's writes
      int woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        int t = s_warp_tot[w];
        if (w < warp) woff += t;
        tot += t;
      }
      const int off = woff + incl - cnt;
      for (int k = 0; k < cnt; k++) {
        c_start[off + k] = env.r_start[k];
        c_end[off + k] = env.r_end[k];
        c_meta[off + k] = (tid << 16) | env.r_meta[k];
      }
      if (warp == 0) {
        unsigned long long base = tile_lookback(p.tile_status, tile, (unsigned long long)tot);
        if (lane == 0) {
          s_tile_base = base;
          if (tile == n_tiles - 1) *p.n_rows = base + (unsigned long long)tot;
        }
      }
      __syncthreads();
      const unsigned long long base = s_tile_base;
      const double id0 = (double)(p.first_person + tile * BLOCK);
      for (int r = tid; r < tot; r += BLOCK) {
        unsigned long long g = base + r;
        if (g < (unsigned long long)p.capacity) {
          int m = c_meta[r];
          p.rowlog[0 * p.capacity + g] = c_end[r];
          p.rowlog[1 * p.capacity + g] = (double)(m & 0xff);
          p.rowlog[2 * p.capacity + g] = id0 + (double)(m >> 16);
          p.rowlog[3 * p.capacity + g] = c_start[r];
          p.rowlog[4 * p.capacity + g] = (double)((m >> 8) & 0xff);
        }
      }
      // the next tile's first __syncthreads() protects c_* against reuse
    }
  }

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
  if (outputs & OUT_REPORT) b += (size_t)g.off_touch * 8 + (size_t)(g.total - g.off_touch) * 4;
  b += (size_t)BLOCK * Traits::MAX_ROWS * (8 + 8 + 4);
  return b;
}

// ---- sequential-stream regime: draws per start position, chain resolution --

// len[p] = number of uniforms a person whose first draw is word first_word+p
// would consume (255 = ran past the generated words)
template <class Traits>
__global__ void __launch_bounds__(BLOCK) seq_len_kernel(const uint32_t *words, int64_t mt_end, int64_t first_word,
                                                        int64_t n_pos, typename Traits::Consts consts,
                                                        unsigned char *len) {
  typedef NullEnv<StreamSource> Env;
  for (int64_t p = (int64_t)blockIdx.x * BLOCK + threadIdx.x; p < n_pos; p += (int64_t)gridDim.x * BLOCK) {
    Env env;
    env.src.w = words;
    env.src.pos = first_word + p;
    env.src.end = mt_end;
    env.src.overrun = 0;
    typename Traits::template Model<Env> person(env, consts, 0);
    person.run();
    int64_t used = env.src.pos - (first_word + p);
    len[p] = (env.src.overrun || used > 250 || person.error_) ? 255 : (unsigned char)used;
  }
}

// For chunk c and entry offset e: walk p += len[p] from c*CHUNK+e to the chunk
// end; record where the chain enters the next chunk and how many persons
// started inside this one.
__global__ void __launch_bounds__(64) chain_chunk_kernel(const unsigned char *len, int64_t n_pos, int max_draws,
                                                         unsigned char *exit_off, int *count) {
  __shared__ unsigned char s_len[CHAIN_CHUNK];
  const int64_t c = blockIdx.x;
  const int64_t lo = c * CHAIN_CHUNK;
  for (int k = threadIdx.x; k < CHAIN_CHUNK; k += blockDim.x) s_len[k] = (lo + k < n_pos) ? len[lo + k] : 255;
  __syncthreads();
  int e = threadIdx.x;
  if (e < max_draws) {
    int cnt = 0;
    int ex = chain_walk(s_len, e, &cnt);
    exit_off[c * CHAIN_MAX_DRAWS + e] = (unsigned char)ex;
    count[c * CHAIN_MAX_DRAWS + e] = cnt;
  }
}

// Serial pass over chunks: entry offset and first person index of every chunk.
// The dependency chain is inherently sequential but tiny; the block stages
// SCAN_BATCH chunks' tables in shared memory and thread 0 walks them there.
// entry = 255 marks "chain broken / not reached".
constexpr int SCAN_BATCH = 512;
__global__ void __launch_bounds__(256) chain_scan_kernel(const unsigned char *exit_off, const int *count,
                                                         int64_t n_chunks, int64_t n, unsigned char *entry,
                                                         int64_t *base) {
  __shared__ unsigned char s_exit[SCAN_BATCH * CHAIN_MAX_DRAWS];
  __shared__ int s_count[SCAN_BATCH * CHAIN_MAX_DRAWS];
  __shared__ unsigned char s_entry[SCAN_BATCH];
  __shared__ int64_t s_base[SCAN_BATCH];
  __shared__ int s_e;
  __shared__ int64_t s_b;
  if (threadIdx.x == 0) {
    s_e = 0;
    s_b = 0;
  }
  for (int64_t c0 = 0; c0 < n_chunks; c0 += SCAN_BATCH) {
    int m = (int)((n_chunks - c0 < SCAN_BATCH) ? (n_chunks - c0) : SCAN_BATCH);
    for (int k = threadIdx.x; k < m * CHAIN_MAX_DRAWS; k += blockDim.x) {
      s_exit[k] = exit_off[c0 * CHAIN_MAX_DRAWS + k];
      s_count[k] = count[c0 * CHAIN_MAX_DRAWS + k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int e = s_e;
      int64_t b = s_b;
      for (int c = 0; c < m; c++) {
        s_entry[c] = (unsigned char)e;
        s_base[c] = b;
        chain_scan_step(&s_exit[c * CHAIN_MAX_DRAWS], &s_count[c * CHAIN_MAX_DRAWS], n, &e, &b);
      }
      s_e = e;
      s_b = b;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
      entry[c0 + k] = s_entry[k];
      base[c0 + k] = s_base[k];
    }
    __syncthreads();
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
  double *out;           // [n_sets][48]: occupancy[4][10], TimeAtRisk[4], n_time_at_risk, pad
  int *error;
};

struct CalibEnv {
  typedef MrgSource Src;
  Src src;
  int *s_occ;     // [4][10]
  double tar[4];  // per-thread TimeAtRisk partial sums
  int tar_n;      // 1 + largest TimeAtRisk index this thread touched
  __device__ __forceinline__ void occupancy(int stage, int cind) {
    if (stage < 4) atomicAdd(&s_occ[stage * 10 + cind], 1);
  }
  __device__ __forceinline__ void time_at_risk(int i, double v) {
    if (i == 0) tar[0] += v;
    else if (i == 1) tar[1] += v;
    else if (i == 2) tar[2] += v;
    else tar[3] += v;
    if (i + 1 > tar_n) tar_n = i + 1;
  }
};

// grid = (blocks_x, n_sets): every parameter set sees the same substreams
// (common random numbers, R/calibSim.R:12-14 with one seed for all sets).
__global__ void __launch_bounds__(BLOCK) calib_kernel(CalibParams p) {
  __shared__ int s_occ[40];
  __shared__ double s_tar[WARPS][4];
  __shared__ int s_tar_n;
  __shared__ int s_error;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int set = blockIdx.y;
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
  env.tar[0] = env.tar[1] = env.tar[2] = env.tar[3] = 0.0;
  env.tar_n = 0;

  Mrg32k3a sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  for (int64_t idx = (int64_t)blockIdx.x * BLOCK + tid; idx < p.n; idx += stride) {
    env.src.g = sub;
    calib::CalibPerson<CalibEnv> person(env, par);
    person.run();
    if (person.error_) atomicCAS(&s_error, 0, person.error_);
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
    if (t != 0.0) atomicAdd(&out[40 + tid], t);
  }
  if (tid == 0) {
    // max over blocks of a non-negative integer-valued double (bit order = value order)
    unsigned long long *cell = reinterpret_cast<unsigned long long *>(&out[44]);
    atomicMax(cell, (unsigned long long)__double_as_longlong((double)s_tar_n));
    if (s_error) atomicCAS(p.error, 0, s_error);
  }
}

}  // namespace msim
