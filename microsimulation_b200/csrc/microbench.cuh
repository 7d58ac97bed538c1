// Instruction weights of the person-r hot path's building blocks, for the issue roofline's FLOOR (SURVEY.md 8(d):
// W = J + D (c_rng + c_samp) + E c_evt + B c_bin with weights "measured once on straight-line code").
//
// One kernel per building block: every thread of a full warp runs the block `reps` times on its own state, no lane
// ever diverges from its warp (except inside exp_rand, whose trip count is data dependent by nature), and ncu's
// smsp__inst_executed.sum of the launch, minus that of the empty kernel with the same loop, divided by warps x reps
// is the block's cost in warp instructions per warp of 32 persons (scripts/inst_floor.py).  The blocks are the
// product's own device functions, compiled with the product's flags.
#pragma once
#include "engine.cuh"
#include "models.cuh"
#include "report.cuh"

namespace msim {

#ifdef __CUDACC__

enum MicroOp { MB_EMPTY = 0, MB_JUMP, MB_U01, MB_WEIBULL, MB_EXP_RAND, MB_HEAP, MB_REPORT, MB_EXP, MB_N_OPS };

struct MicroParams {
  StreamStart st;
  ReportGeom geom;
  int reps;
  double *sink;  // one double per thread (keeps the work alive)
  int *counters; // [geom.total] scratch for the report block
};

struct MicroAcc {  // the cheapest accumulator with real memory operations: plain adds on thread-private cells
  double *f;
  int *i;
  __device__ __forceinline__ void addf(int idx, double v) { f[idx & 7] += v; }
  __device__ __forceinline__ void addi(int idx, int v) { i[idx & 7] += v; }
  __device__ __forceinline__ void addc(int idx) { i[idx & 7] += 1; }
};

template <int OP>
__global__ void __launch_bounds__(256) microbench_kernel(const __grid_constant__ MicroParams p) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  Mrg32k3a g = thread_start_state(p.st, gtid);
  MrgSource src;
  src.g = g;
  RMath<MrgSource> R(&src);
  double acc = 0.0, t = 1.0 + (double)(g.s[0] & 1023u) * 0.0625;  // a per-thread value the compiler cannot know
  double cf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int ci[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  MicroAcc macc = {cf, ci};
  ReportCursor cur;
  report_cursor_start(p.geom, cur);
  EventHeap<4> heap;
  heap.reset();
  heap.insert(1.0e9, pack_meta(A_Event, 0, 0));  // a standing entry: the measured insert + pop always sees 2 entries
  for (int r = 0; r < p.reps; r++) {
    if (OP == MB_JUMP) {
      src.g.jump(p.st.stride);
    } else if (OP == MB_U01) {
      acc += src.g.u01();
    } else if (OP == MB_WEIBULL) {  // rweibull's arithmetic on a given uniform: scale * pow(-log(u), 1/shape)
      t = 0.25 + 0.5 * (t - floor(t));  // u in (0.25, 0.75), data dependent
      acc += 64.0218 * fm_pow(-fm_log(t), 0.09513);
      t += acc;
    } else if (OP == MB_EXP_RAND) {
      acc += R.exp_rand();
    } else if (OP == MB_HEAP) {  // scheduleAt + pop of the minimum, as one handled event costs the queue
      double tt;
      int m;
      heap.insert(t, pack_meta(A_Event, 1, 0));
      heap.pop_first(tt, m);
      t = tt + 1.0 + (double)(m & 1);
      acc += tt;
    } else if (OP == MB_REPORT) {  // EventReport::add for one interval [lhs, rhs] (report.cuh)
      const double lhs = cur.t, rhs = lhs + 0.37 + t * 0.001;
      report_add(p.geom, macc, cur, (int)(gtid & 3), (int)(gtid & 7), lhs, rhs);
      if (rhs > 99.0) report_cursor_start(p.geom, cur);
    } else if (OP == MB_EXP) {
      t = -(t - floor(t)) * 3.0;
      acc += fm_exp(t);
      t = acc;
    }
  }
  for (int k = 0; k < 8; k++) acc += cf[k] + ci[k];
  for (int k = 0; k < 6; k++) acc += src.g.s[k];
  p.sink[gtid] = acc + t;
}

#endif  // __CUDACC__

}  // namespace msim
