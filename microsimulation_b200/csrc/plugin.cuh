// Out-of-tree models: the device mirror of the reference's plugin mechanism.
//
// In the reference a user model is a host C++ class derived from ssim::cProcess with init() and
// handleMessage(const cMessage*) (inst/include/microsimulation.h:325-447), typically compiled on the fly with
// Rcpp::sourceCpp against the package's headers (inst/doc/examples.org:254-312, README.org:148-301).  Here it is a
// .cu file compiled with nvcc against THIS header:
//
//     #include "plugin.cuh"
//     struct MyConsts { double rate; };                       // parameters, a plain struct
//     template <class Env>
//     struct MyModel : msim::cProcess<MyModel<Env>, typename Env::Src, 8> {   // 8 = most actions ever pending
//       MSIM_PLUGIN_MODEL(MyModel, MyConsts, 8)                // constructor (env, K = parameters, id) + the names handler bodies use
//       MSIM_DEV void init() { ... scheduleAt(now() + R.rexp(30.0), toOnset); ... }
//       MSIM_DEV void handleMessage(int kind) { ... env.report(state, kind, previous(), now()); env.value(0, x); ... }
//       MSIM_DEV void stop() { ... }                           // optional (Process::stop hook)
//     };
//     struct MyTraits {
//       typedef MyConsts Consts;
//       template <class Env> using Model = MyModel<Env>;
//       static constexpr int N_VALUES = 2;                     // per-person outputs, env.value(k, v), k < N_VALUES
//       static const char *const *value_names() { static const char *const n[] = {"QALY", "costs"}; return n; }
//     };
//     MSIM_PLUGIN(my_model, MyTraits)     // exports my_model_run / my_model_free / my_model_last_error / ...
//
// and is driven through the C structs of include/msim_plugin.h.  Handler bodies port statement for statement; what
// differs from the reference is listed in INTEGRATION.md (event kinds are integers, parameters a struct, outputs go
// through `env`).  Both RNG regimes of the built-in models are available: one RngStream substream per person, or R's
// Mersenne-Twister stream consumed person after person (a model may draw up to 250 uniforms per person there).
//
// The plugin is a shared library of its own, with its own device context (stream, buffers); it shares no state with
// libmsim_b200.so.  Compile (see examples/plugins/Makefile):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden
//          -shared -I<repo>/microsimulation_b200/csrc -I<repo>/include -o libmy_model.so my_model.cu
#pragma once
#include "../../include/msim_plugin.h"
#include "host_runtime.h"
#include "host_cohort.h"

namespace msim {

constexpr unsigned OUT_VALUES = 8u;

template <class Consts>
struct PluginParams {
  StreamStart st;            // substream regime
  const uint32_t *mt_words;  // stream regime
  int64_t mt_end;
  const int64_t *starts;
  int64_t first_person, n;
  unsigned outputs;  // OUT_REPORT | OUT_VALUES
  Consts consts;
  ReportGeom geom;
  double *dense;   // geom.total report cells (global, zeroed before launch)
  double *values;  // [n][N_VALUES]
  int *error;
};

// What a plugin model sees as `env`
template <class SrcT, int NV>
struct PluginEnv {
  typedef SrcT Src;
  Src src;
  unsigned outputs;
  const ReportGeom *geom;
  SharedAcc acc;
  ReportCursor cursor;
  double *my_values;  // this person's N_VALUES cells
  int code = 0;
  __device__ __forceinline__ void begin_person() { report_cursor_start(*geom, cursor); }
  // EventReport::add(state, event, lhs, rhs)
  __device__ __forceinline__ void report(int state, int kind, double lhs, double rhs) {
    if (outputs & OUT_REPORT) report_add(*geom, acc, cursor, state, kind, lhs, rhs);
  }
  // per-person output k (what the reference's models push_back into their own report vectors)
  __device__ __forceinline__ void value(int k, double v) {
    if (NV > 0 && (outputs & OUT_VALUES) && k >= 0 && k < NV) my_values[k] = v;
  }
  __device__ __forceinline__ void row(double, double, double, int, int) {}
};
// (the draw-counting pass of the stream regime runs the model over kernels.cuh's NullEnv: same hooks, no sinks)

template <class Traits, bool STREAM>
__global__ void __launch_bounds__(BLOCK) plugin_kernel(const __grid_constant__ PluginParams<typename Traits::Consts> p) {
  typedef typename std::conditional<STREAM, StreamSource, MrgSource>::type Src;
  typedef PluginEnv<Src, Traits::N_VALUES> Env;
  extern __shared__ __align__(16) unsigned char plugin_smem[];
  __shared__ int s_error;
  const int total = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  const int n_f = (p.outputs & OUT_REPORT) ? p.geom.off_dstart : 0;
  const int tid = threadIdx.x, lane = tid & 31;
  Env env;
  env.outputs = p.outputs;
  env.geom = &p.geom;
  env.acc.f = reinterpret_cast<double *>(plugin_smem);
  env.acc.i = reinterpret_cast<int *>(env.acc.f + n_f);
  env.acc.n_f = n_f;
  if (p.outputs & OUT_REPORT) shared_acc_clear(env.acc, total);
  if (tid == 0) s_error = 0;
  __syncthreads();

  Mrg32k3a sub;
  if constexpr (!STREAM) sub = thread_start_state(p.st, blockIdx.x * BLOCK + tid);
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  int err = 0;
  // one person per lane; the warp is re-formed at the head of every pass (persons end after different numbers of
  // events), with a trip count that is the warp's
  for (int64_t i0 = (int64_t)blockIdx.x * BLOCK + (tid & ~31); i0 < p.n; i0 += stride) {
    const int64_t i = i0 + lane;
    __syncwarp();
    if (i < p.n) {
      if constexpr (STREAM) {
        env.src.w = p.mt_words;
        env.src.pos = p.starts[i];
        env.src.end = p.mt_end;
        env.src.overrun = 0;
      } else {
        env.src.g = sub;
      }
      env.my_values = p.values + (size_t)i * (Traits::N_VALUES > 0 ? Traits::N_VALUES : 1);
      typename Traits::template Model<Env> person(env, p.consts, p.first_person + i);
      person.run();
      if (person.error_) err = person.error_;
      if constexpr (STREAM) {
        if (env.src.overrun) err = -6;
      }
    }
    __syncwarp();
    if constexpr (!STREAM) sub.jump(p.st.stride);
  }
  if (err) atomicCAS(&s_error, 0, err);
  __syncthreads();
  if (p.outputs & OUT_REPORT) shared_acc_flush(env.acc, total, p.dense);
  if (tid == 0 && s_error) atomicCAS(p.error, 0, s_error);
}

}  // namespace msim

namespace msim_host {
using namespace msim;

struct PluginBufs {
  DevBuf values;
};
static PluginBufs g_plugin;

inline void plugin_zero(msim_plugin_result *r) { memset(r, 0, sizeof *r); }

inline void plugin_free(msim_plugin_result *r) {
  if (!r) return;
  msim_table *tabs[] = {&r->values, &r->report.pt, &r->report.ut, &r->report.events, &r->report.prev, &r->report.costs,
                        &r->report.indiv};
  for (msim_table *t : tabs) {
    free(t->data);
    t->data = nullptr;
    t->nrow = 0;
  }
}

template <class Traits>
int plugin_run(const msim_plugin_call *c, msim_plugin_result *r) {
  typedef typename Traits::Consts Consts;
  constexpr int NV = Traits::N_VALUES;
  int rc;
  if (!c || !r) return fail(MSIM_ERR_ARG, "null argument");
  plugin_zero(r);
  if (c->n < 0 || c->first_person < 0) return fail(MSIM_ERR_ARG, "n and first_person must be >= 0");
  if (c->regime != MSIM_PLUGIN_SUBSTREAMS && c->regime != MSIM_PLUGIN_MT_STREAM) return fail(MSIM_ERR_ARG, "unknown RNG regime");
  if (c->consts_bytes != sizeof(Consts) || (sizeof(Consts) > 1 && !c->consts))
    return fail(MSIM_ERR_ARG, "consts_bytes does not match the model's parameter struct");
  if (c->outputs & ~(unsigned)(MSIM_OUT_REPORT | MSIM_PLUGIN_OUT_VALUES)) return fail(MSIM_ERR_ARG, "unknown output flag");
  if ((c->outputs & MSIM_OUT_REPORT) && (c->n_state < 1 || c->n_state > 64 || c->n_event < 1 || c->n_event > 64))
    return fail(MSIM_ERR_ARG, "n_state and n_event must be in 1..64");
  if (g.device != c->device && g.inited) return fail(MSIM_ERR_STATE, "this plugin is already bound to another device");
  g.device = c->device;
  if ((rc = ensure_init())) return rc;
  const int64_t n = c->n;

  PluginParams<Consts> p;
  memset(&p, 0, sizeof p);
  if (c->consts) memcpy(&p.consts, c->consts, sizeof(Consts));
  p.first_person = c->first_person;
  p.n = n;
  p.outputs = ((c->outputs & MSIM_OUT_REPORT) ? OUT_REPORT : 0u) | (((c->outputs & MSIM_PLUGIN_OUT_VALUES) && NV > 0) ? OUT_VALUES : 0u);
  if (p.outputs & OUT_REPORT) p.geom = ages_geom(c->n_state, c->n_event, c->discount_rate);
  const int cells = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  const size_t smem = (p.outputs & OUT_REPORT) ? (size_t)p.geom.off_dstart * 8 + (size_t)(p.geom.total - p.geom.off_dstart) * 4 : 0;
  const size_t n_val = (size_t)std::max<int64_t>(n, 1) * (size_t)(NV > 0 ? NV : 1);
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  if ((rc = ensure(g.dense, sizeof(double) * (size_t)std::max(cells, 1)))) return rc;
  if ((rc = ensure(g_plugin.values, sizeof(double) * n_val))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * (size_t)std::max(cells, 1), g.stream));
  CK(cudaMemsetAsync(g_plugin.values.p, 0, sizeof(double) * n_val, g.stream));
  p.dense = (double *)g.dense.p;
  p.values = (double *)g_plugin.values.p;
  p.error = small_view().error;

  const int64_t tiles = (n + BLOCK - 1) / BLOCK;
  int per_sm = 0;
  int64_t end_word = 0;
  if (c->regime == MSIM_PLUGIN_SUBSTREAMS) {
    if (!mrg_seed_ok(c->seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
    auto kern = plugin_kernel<Traits, false>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
    if (per_sm < 1) return fail(MSIM_ERR_CUDA, "the model's kernel does not fit on an SM");
    int64_t gmax = std::min<int64_t>((int64_t)per_sm * g.sm_count, JUMP_TAB * (int64_t)JUMP_TAB / BLOCK);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(gmax, tiles));
    uint32_t state[6];
    seed_to_u32(c->seed, state);
    stream_start(state, c->first_person, (int64_t)grid * BLOCK, &p.st);
    CK(cudaEventRecord(g.ev0, g.stream));
    if (n > 0) {
      kern<<<grid, BLOCK, smem, g.stream>>>(p);
      CK(cudaGetLastError());
      g.launches++;
    }
  } else {
    if (!c->mt_state) return fail(MSIM_ERR_ARG, "mt_state is required in the Mersenne-Twister regime");
    if (c->mt_state[0] < 1 || c->mt_state[0] > 624) return fail(MSIM_ERR_ARG, "MT position must be in 1..624");
    if (c->max_draws < 1 || c->max_draws > 250) return fail(MSIM_ERR_ARG, "max_draws must be in 1..250");
    memcpy(g.rng.mt, c->mt_state, sizeof g.rng.mt);
    g.rng.mt_seeded = true;
    auto kern = plugin_kernel<Traits, true>;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem));
    if (per_sm < 1) return fail(MSIM_ERR_CUDA, "the model's kernel does not fit on an SM");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)per_sm * g.sm_count, tiles));
    CK(cudaEventRecord(g.ev0, g.stream));
    if (n > 0) {
      // plan for the expected draws per person first, then for the stated maximum
      const double mean = c->mean_draws > 0 ? c->mean_draws : (double)c->max_draws;
      int64_t n_pos = std::min<int64_t>((int64_t)(1.3 * mean * (double)n) + 8192, (int64_t)c->max_draws * n);
      bool short_plan = false;
      int64_t mt_end = 0;
      if ((rc = stream_offsets_attempt<Traits>(n, p.consts, n_pos, c->max_draws, false, &short_plan, &end_word, &mt_end))) return rc;
      if (short_plan && n_pos < (int64_t)c->max_draws * n) {
        n_pos = (int64_t)c->max_draws * n;
        if ((rc = stream_offsets_attempt<Traits>(n, p.consts, n_pos, c->max_draws, false, &short_plan, &end_word, &mt_end))) return rc;
      }
      if (short_plan) return fail(MSIM_ERR_STREAM, "a person drew more uniforms than max_draws allows");
      CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
      p.mt_words = (const uint32_t *)g.mt_words.p;
      p.mt_end = mt_end;
      p.starts = (const int64_t *)g.starts.p;
      kern<<<grid, BLOCK, smem, g.stream>>>(p);
      CK(cudaGetLastError());
      g.launches++;
    }
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  if ((rc = finish_run(nullptr))) return rc;

  if (c->regime == MSIM_PLUGIN_MT_STREAM && n > 0 && end_word > 0) {
    // leave R's generator where the reference would: after the last person's last draw
    const int64_t batch = (end_word - 1) / 624;
    CK(cudaMemcpy(&g.rng.mt[1], (const uint32_t *)g.mt_words.p + batch * 624, sizeof(uint32_t) * 624, cudaMemcpyDeviceToHost));
    g.rng.mt[0] = (uint32_t)(end_word - batch * 624);
    memcpy(c->mt_state, g.rng.mt, sizeof g.rng.mt);
  }
  if (p.outputs & OUT_REPORT) {
    std::vector<double> d((size_t)cells);
    CK(cudaMemcpy(d.data(), g.dense.p, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
    dense_to_report_impl(d.data(), p.geom, &r->report);
  }
  if (p.outputs & OUT_VALUES) {
    std::vector<double> rows((size_t)n * NV);
    if (n > 0) CK(cudaMemcpy(rows.data(), g_plugin.values.p, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost));
    r->values.nrow = n;
    r->values.ncol = NV;
    r->values.colnames = Traits::value_names();
    r->values.data = (double *)malloc(sizeof(double) * std::max<size_t>(rows.size(), 1));
    if (!r->values.data) return fail(MSIM_ERR_CUDA, "out of host memory");
    for (int64_t i = 0; i < n; i++)
      for (int k = 0; k < NV; k++) r->values.data[(size_t)k * n + i] = rows[(size_t)i * NV + k];  // column-major, like every msim_table
  }
  r->kernel_ms = g.last_ms;
  r->launches = g.launches;
  return MSIM_OK;
}

}  // namespace msim_host

// inside a model: the constructor (environment, parameters, person index) and the names handler bodies use
#define MSIM_PLUGIN_MODEL(NAME, CONSTS, CAP)                                                                 \
  typedef msim::cProcess<NAME<Env>, typename Env::Src, CAP> Base;                                            \
  using Base::now;                                                                                           \
  using Base::previous;                                                                                      \
  using Base::previousEventTime;                                                                             \
  using Base::R;                                                                                             \
  using Base::scheduleAt;                                                                                    \
  using Base::RemoveKind;                                                                                    \
  using Base::CancelKind;                                                                                    \
  using Base::CancelEvents;                                                                                  \
  using Base::cancel_events;                                                                                 \
  using Base::cancel_kind;                                                                                   \
  using Base::stop_process;                                                                                  \
  using Base::stop_simulation;                                                                               \
  Env &env;                                                                                                  \
  const CONSTS &K;                                                                                           \
  int64_t id;                                                                                                \
  MSIM_DEV NAME(Env &e, const CONSTS &k, int64_t i) : Base(&e.src), env(e), K(k), id(i) { e.begin_person(); }

#define MSIM_PLUGIN_EXPORT extern "C" __attribute__((visibility("default")))

// the library's entry points (include/msim_plugin.h describes them)
#define MSIM_PLUGIN(NAME, TRAITS)                                                                             \
  MSIM_PLUGIN_EXPORT int NAME##_run(const msim_plugin_call *call, msim_plugin_result *result) {              \
    return msim_host::plugin_run<TRAITS>(call, result);                                                      \
  }                                                                                                           \
  MSIM_PLUGIN_EXPORT void NAME##_free(msim_plugin_result *result) { msim_host::plugin_free(result); }        \
  MSIM_PLUGIN_EXPORT const char *NAME##_last_error(void) { return msim_host::g_err.c_str(); }                \
  MSIM_PLUGIN_EXPORT int NAME##_n_values(void) { return TRAITS::N_VALUES; }                                  \
  MSIM_PLUGIN_EXPORT unsigned long NAME##_consts_bytes(void) { return sizeof(TRAITS::Consts); }
