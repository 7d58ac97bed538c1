// Calibration sweep (BASELINE configs[4], R/calibSim.R:10-40) with the random draws SHARED between parameter sets.
//
// Every parameter set of a sweep is simulated from the same substreams (common random numbers: callCalibrationSimulation
// re-seeds with one seed for all sets, R/calibSim.R:12-14, calibperson-r.cc:180-193).  The model (calibperson-r.cc:95-171)
// takes its draws in an order that does not depend on the parameters:
//
//   init():         runif (diseasepot), norm_rand (lam1), exp_rand (time to precursor), runif (Gumbel time of death)
//   toPrecursor:    exp_rand (dwell time to pre-clinical)         — only if diseasepot and the person is still alive
//   toPreClinical:  norm_rand (dwell time to clinical)            — only after toPrecursor's draw
//
// (rnorm(mu, sigma) = mu + sigma * norm_rand() and rexp(scale) = scale * exp_rand(): the parameters scale the draws, they
// do not change how many uniforms are taken — except rnorm with sigma == 0, which draws nothing; the host sends sweeps
// with such a set to the event-queue kernels.)  So person i's six variates are the same numbers in every set, and a
// block computes them ONCE per person (phase 1: one person per thread, full warps) and then runs all its parameter
// sets over them (phase 2: one SET per thread, the person's variates read from shared memory by broadcast).
//
// Phase 2 needs no event queue either.  The model's events are one chain — toPrecursor -> toPreClinical -> toClinical,
// each scheduled by its predecessor's handler — racing one toDeath and crossing a fixed ladder of Count events at
// 10, 20, .., 100.  With all event times distinct, Sim::run_simulation's pop order IS the order of the times, so the
// stage at every Count and the clinical time at death follow from comparisons of t1 <= t2 <= t3 with tD and the
// ladder (straight_life below).  Where two of those times are EQUAL the pop order is heap.h's business (it depends
// on the heap's shape): such a person raises the tie flag and the host repeats the sweep with the binary-heap kernel
// (calib_kernel<false>), exactly as the ordered-array kernel does.  Times are computed with the engine's own
// roundings (scheduleAt: clock + (t - clock), microsimulation.h:347 + ssim.cc:97).
//
// Occupancy is kept as histograms of c1 <= c2 <= c3 <= nC — the number of Count events a person sees before each
// transition and before death (nC does not depend on the parameters) — in per-thread columns of shared memory (no
// atomics), turned into the report's occupancy[stage][decade] when the block ends:
//   DiseaseFree[k] = #(k < c1), Precursor[k] = #(c1 <= k < c2), PreClinical[k] = #(c2 <= k < c3), Clinical[k] = #(c3 <= k < nC).
// Most persons never leave Precursor, so c2 and c3 are stored as differences from nC.
//
// Same output vector as calib_kernel (kernels.cuh): [set][48].
#pragma once
#include "models.cuh"
#ifdef __CUDACC__
#include "kernels.cuh"
#endif

namespace msim {
namespace calib {

// what one person takes from its substream, in stream order (tD: the Gumbel time of death, already transformed)
struct Draws {
  double u0, z1, e1, tD, e2, z2;
};

template <class Src>
MSIM_DEV Draws draw_person(Src &src) {
  RMath<Src> R(&src);
  Draws d;
  d.u0 = R.runif(0, 1);
  d.z1 = R.norm_rand();
  d.e1 = R.exp_rand();
  const double x = R.runif(0, 1);
  d.tD = 65 - 15 * fm_log(-fm_log(x));  // calibperson-r.cc:105-106
  d.e2 = R.exp_rand();
  d.z2 = R.norm_rand();
  return d;
}

// the Count ladder with a rung below every time and one above: RUNGS[a] < t <= RUNGS[a + 1] brackets t >= 0
#ifdef __CUDACC__
__constant__
#else
static
#endif
    const double MSIM_CALIB_RUNGS[12] = {-1.0, 10.0, 20.0, 30.0, 40.0, 50.0, 60.0, 70.0, 80.0, 90.0, 100.0, INFINITY};

// number of Count events (at 10, 20, .., 100) strictly before time t >= 0; *tie is set if t coincides with one (or
// is infinite, which no event time of a valid sweep is).  `rungs`: MSIM_CALIB_RUNGS or a copy of it (the kernel keeps
// one in shared memory: lanes look up different rungs, which the constant cache would serve one address at a time)
MSIM_DEV int counts_before(double t, const double *rungs, bool *tie) {
  // floor(t / 10) give or take one, within 0..10
#ifdef __CUDA_ARCH__
  int a = __double2int_rz(t * 0.1);  // saturates; NaN gives 0
  a = a < 0 ? 0 : (a > 10 ? 10 : a);
#else
  const double q = t * 0.1;
  int a = q >= 10.0 ? 10 : (q > 0.0 ? (int)q : 0);
#endif
  const double lo = rungs[a], hi = rungs[a + 1];  // one exact comparison either side settles it
  *tie = *tie | (t == lo) | (t == hi);
  a += (int)(t > hi) - (int)(t <= lo);
  return a;
}

struct SetPar {
  double Lam1, sigm1, p2, lam2, mu3, sd3;  // sd3 = tau3 * mu3, the sigma handed to rnorm (calibperson-r.cc:150)
};

// Sweeps the straight-line form can run: every draw is taken in every set (no rnorm with sigma == 0) and no event
// is scheduled in the past (lam2 >= 0); anything else goes through the event queue.
inline bool straight_ok(const double par[6]) {
  for (int k = 0; k < 6; k++)
    if (!(par[k] - par[k] == 0.0)) return false;  // NaN or infinite
  return par[1] != 0.0 && par[5] * par[4] != 0.0 && par[3] >= 0.0;
}

// One person under one parameter set.  c1, c2, c3: Count events seen before toPrecursor / toPreClinical / toClinical
// (capped at nC, the number seen before death); clin: clinTime as the toDeath handler leaves it (calibperson-r.cc:120).
MSIM_DEV void straight_life(const Draws &d, int nC, const SetPar &P, int *c1, int *c2, int *c3, double *clin, bool *tie) {
  const double lam1 = fm_exp(P.Lam1 + P.sigm1 * d.z1);
  const double t1 = lam1 * d.e1;  // scheduled from time 0: 0 + (t1 - 0) == t1
  if (t1 == d.tD) *tie = true;
  const int m1 = counts_before(t1, MSIM_CALIB_RUNGS, tie);
  *c1 = m1 < nC ? m1 : nC;
  *c2 = nC;
  *c3 = nC;
  *clin = d.tD;
  if (d.u0 < P.p2 && t1 < d.tD) {
    const double dwell = t1 + P.lam2 * d.e2;
    const double t2 = t1 + (dwell - t1);
    if (t2 == d.tD) *tie = true;
    const int m2 = counts_before(t2, MSIM_CALIB_RUNGS, tie);
    *c2 = m2 < nC ? m2 : nC;
    if (t2 < d.tD) {
      const double dwell3 = t2 + fm_exp(P.mu3 + P.sd3 * d.z2);
      const double t3 = t2 + (dwell3 - t2);
      if (t3 == d.tD) *tie = true;
      const int m3 = counts_before(t3, MSIM_CALIB_RUNGS, tie);
      *c3 = m3 < nC ? m3 : nC;
      if (t3 < d.tD) *clin = t3;
    }
  }
}

}  // namespace calib

#ifdef __CUDACC__

constexpr int CS_BINS = 11;  // c in 0..10

struct CalibSharedParams {
  StreamStart st;  // stride = M^(gridDim.x*BLOCK)
  int64_t n;
  const double *runpar;  // [n_sets][6]
  double *out;           // [n_sets][48], as calib_kernel's
  int set0;              // first parameter set of this launch
  int n_here;            // sets of this launch; blockIdx.y covers sets [set0 + y*spb, set0 + min((y+1)*spb, n_here))
  int spb;               // sets per block (<= BLOCK)
  int *tie;
  double *partial;       // [n_here][gridDim.x][4]
};

// dynamic shared memory: 6 x BLOCK doubles (draws), 8 x BLOCK doubles (TimeAtRisk: every thread's phase-1 sums and
// phase-2 corrections, kept here rather than in registers for the whole run), BLOCK ints (nC), 3 x BLOCK ints
// (persons leaving the risk set early), 3 x CS_BINS x BLOCK ints (histograms, one column per thread)
constexpr size_t CS_SMEM = sizeof(double) * 14 * BLOCK + sizeof(int) * 4 * BLOCK + sizeof(int) * 3 * CS_BINS * BLOCK;

// a counter in a column of shared memory that only this thread touches: one instruction, no read-back
__device__ __forceinline__ void cs_bump(unsigned shared_addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(shared_addr), "r"(v) : "memory");
}

// TimeAtRisk (calibperson-r.cc:124-135): term k of a person with clinical time c, and whether it is added at all
__device__ __forceinline__ double cs_term(int k, double c) {
  return k == 0 ? fmin(20.0, c) : (c >= 20.0 * k ? fmin(20.0 * (k + 1), c) : 0.0);
}

#ifndef MSIM_CS_MINB
#define MSIM_CS_MINB 3
#endif
__global__ void __launch_bounds__(BLOCK, MSIM_CS_MINB) calib_shared_kernel(CalibSharedParams p) {
  extern __shared__ __align__(16) unsigned char cs_raw[];
  double *s_d = reinterpret_cast<double *>(cs_raw);  // [6][BLOCK]
  double *s_base = s_d + 6 * BLOCK;                  // [4][BLOCK]
  double *s_dev = s_base + 4 * BLOCK;                // [4][BLOCK]
  int *s_nC = reinterpret_cast<int *>(s_dev + 4 * BLOCK);
  int *s_drop = s_nC + BLOCK;   // [3][BLOCK]
  int *s_h = s_drop + 3 * BLOCK;  // [3][CS_BINS][BLOCK]
  __shared__ int s_histN[CS_BINS];
  __shared__ double s_rungs[12];
  __shared__ int s_tie;
  __shared__ double s_base_sum[4];  // TimeAtRisk of the block's persons had nobody become clinical (no parameter in it)
  __shared__ int s_reach[3];    // persons of the block that die at or after 20, 40, 60
  const int tid = threadIdx.x;
  const int set_first = blockIdx.y * p.spb;
  const int n_sets_blk = min(p.spb, p.n_here - set_first);  // >= 1
  const int G = BLOCK / n_sets_blk;                         // person groups
  const bool active = tid < G * n_sets_blk;
  const int s = tid % n_sets_blk, g = tid / n_sets_blk;
  for (int k = tid; k < 3 * CS_BINS * BLOCK; k += BLOCK) s_h[k] = 0;
  for (int k = 0; k < 8; k++) s_base[k * BLOCK + tid] = 0.0;  // s_base and s_dev
  for (int k = 0; k < 3; k++) s_drop[k * BLOCK + tid] = 0;
  if (tid < CS_BINS) s_histN[tid] = 0;
  if (tid < 12) s_rungs[tid] = calib::MSIM_CALIB_RUNGS[tid];
  if (tid < 3) s_reach[tid] = 0;
  if (tid == 0) s_tie = 0;

  calib::SetPar P;
  {
    const double *rp = p.runpar + 6 * (size_t)(p.set0 + set_first + (active ? s : 0));
    P.Lam1 = rp[0];
    P.sigm1 = rp[1];
    P.p2 = rp[2];
    P.lam2 = rp[3];
    P.mu3 = rp[4];
    P.sd3 = rp[5] * rp[4];
  }
  // TimeAtRisk: phase 1 adds every person's terms as if nobody became clinical (no parameter in them: s_base, this
  // thread's persons); phase 2 corrects for the persons of this thread's set who did, and so leave the risk set at t3
  // instead of tD (s_dev, s_drop)
  bool tie = false;
  // this thread's histogram columns, as shared-window addresses; bin b of a column is b * BLOCK ints on
  const unsigned h1 = (unsigned)__cvta_generic_to_shared(s_h + tid), d2 = h1 + 4u * CS_BINS * BLOCK, d3 = d2 + 4u * CS_BINS * BLOCK;
  constexpr unsigned BIN = 4u * BLOCK;

  MrgSource src;
  Mrg32k3a sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  for (int64_t base = (int64_t)blockIdx.x * BLOCK; base < p.n; base += stride) {
    __syncthreads();  // the previous tile has been read (and the zeroing above is done)
    // ---- phase 1: this thread's person
    {
      const int64_t idx = base + tid;
      double tD = -1.0;
      if (idx < p.n) {
        src.g = sub;
        const calib::Draws d = calib::draw_person(src);
        bool t0 = false;
        const int nC = calib::counts_before(d.tD, s_rungs, &t0);
        if (t0) tie = true;
        s_d[0 * BLOCK + tid] = d.u0;
        s_d[1 * BLOCK + tid] = d.z1;
        s_d[2 * BLOCK + tid] = d.e1;
        s_d[4 * BLOCK + tid] = d.e2;
        s_d[5 * BLOCK + tid] = d.z2;
        s_nC[tid] = nC;
        atomicAdd(&s_histN[nC], 1);
        tD = d.tD;
#pragma unroll
        for (int k = 0; k < 4; k++) s_base[k * BLOCK + tid] += cs_term(k, tD);
        if (tD >= 20.0) atomicAdd(&s_reach[tD >= 60.0 ? 2 : tD >= 40.0 ? 1 : 0], 1);
      }
      s_d[3 * BLOCK + tid] = tD;
      sub.jump(p.st.stride);
    }
    __syncthreads();
    // ---- phase 2: this thread's parameter set over the tile's persons (every G-th, from g)
    if (active) {
      for (int j = g; j < BLOCK; j += G) {
        const double tD = s_d[3 * BLOCK + j];
        if (tD < 0.0) break;  // past the last person (persons of a tile are in index order)
        const double u0 = s_d[0 * BLOCK + j], z1 = s_d[1 * BLOCK + j], e1 = s_d[2 * BLOCK + j];
        const int nC = s_nC[j];
        // calib::straight_life, with the branch for persons who go on to pre-clinical disease written out so that the
        // other draws are only read there
        const double t1 = fm_exp_bank(P.Lam1 + P.sigm1 * z1) * e1;
        tie = tie | (t1 == tD);
        cs_bump(h1 + (unsigned)min(calib::counts_before(t1, s_rungs, &tie), nC) * BIN, 1);
        if (u0 < P.p2 && t1 < tD) {
          const double dwell = t1 + P.lam2 * s_d[4 * BLOCK + j];
          const double t2 = t1 + (dwell - t1);
          if (t2 == tD) tie = true;
          const int c2 = min(calib::counts_before(t2, s_rungs, &tie), nC);
          int c3 = nC;
          if (t2 < tD) {
            const double dwell3 = t2 + fm_exp_bank(P.mu3 + P.sd3 * s_d[5 * BLOCK + j]);
            const double t3 = t2 + (dwell3 - t2);
            if (t3 == tD) tie = true;
            c3 = min(calib::counts_before(t3, s_rungs, &tie), nC);
            if (t3 < tD) {  // clinical before death: clinTime = t3
#pragma unroll
              for (int k = 0; k < 4; k++) s_dev[k * BLOCK + tid] += cs_term(k, t3) - cs_term(k, tD);
#pragma unroll
              for (int k = 0; k < 3; k++) s_drop[k * BLOCK + tid] += (int)(tD >= 20.0 * (k + 1)) - (int)(t3 >= 20.0 * (k + 1));
            }
          }
          if (c2 != nC) {
            cs_bump(d2 + (unsigned)c2 * BIN, 1);
            cs_bump(d2 + (unsigned)nC * BIN, -1);
          }
          if (c3 != nC) {
            cs_bump(d3 + (unsigned)c3 * BIN, 1);
            cs_bump(d3 + (unsigned)nC * BIN, -1);
          }
        }
      }
    }
  }
  if (tie) s_tie = 1;
  __syncthreads();

  // ---- the block's results.  The phase-1 sums are added in thread order, so the result does not depend on timing
  if (tid < 4) {
    double t = 0.0;
    for (int k = 0; k < BLOCK; k++) t += s_base[tid * BLOCK + k];
    s_base_sum[tid] = t;
  }
  __syncthreads();
  if (tid < n_sets_blk) {
    const int set_local = set_first + tid;  // within this launch
    double *out = p.out + 48 * (size_t)(p.set0 + set_local);
    // histograms: sum the groups' columns, then cumulate
    int F1 = 0, D2 = 0, D3 = 0, FN = 0, N = 0;
    for (int b = 0; b < CS_BINS; b++) N += s_histN[b];
    for (int k = 0; k < 10; k++) {
      int a1 = 0, a2 = 0, a3 = 0;
      for (int gg = 0; gg < G; gg++) {
        const int col = gg * n_sets_blk + tid;
        a1 += s_h[k * BLOCK + col];
        a2 += s_h[(CS_BINS + k) * BLOCK + col];
        a3 += s_h[(2 * CS_BINS + k) * BLOCK + col];
      }
      F1 += a1;  // #(c1 <= k)
      D2 += a2;  // #(c2 <= k) - #(nC <= k)
      D3 += a3;
      FN += s_histN[k];
      const int occ[4] = {N - F1, F1 - (FN + D2), D2 - D3, D3};  // F2 = FN + D2, F3 = FN + D3; Clinical = F3 - FN
      for (int st = 0; st < 4; st++)
        if (occ[st]) atomicAdd(&out[st * 10 + k], (double)occ[st]);
    }
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    int drop[3] = {0, 0, 0};
    for (int gg = 0; gg < G; gg++) {
      const int col = gg * n_sets_blk + tid;
      for (int k = 0; k < 4; k++) t[k] += s_dev[k * BLOCK + col];
      for (int k = 0; k < 3; k++) drop[k] += s_drop[k * BLOCK + col];
    }
    double *q = p.partial + ((size_t)set_local * gridDim.x + blockIdx.x) * 4;
    for (int k = 0; k < 4; k++) q[k] = s_base_sum[k] + t[k];
    // the list has as many entries as the largest index any person reached: persons with clinTime >= 20, 40, 60
    const int at60 = s_reach[2] - drop[2], at40 = s_reach[2] + s_reach[1] - drop[1], at20 = s_reach[2] + s_reach[1] + s_reach[0] - drop[0];
    const int reached = N > 0 ? 1 + (at20 > 0) + (at40 > 0) + (at60 > 0) : 0;
    for (int k = 0; k < reached; k++) atomicAdd(&out[44 + k], 1.0);
  }
  if (tid == 0 && s_tie) *p.tie = 1;
}

#endif  // __CUDACC__

}  // namespace msim
