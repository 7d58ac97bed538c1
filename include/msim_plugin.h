/* msim_plugin.h — C ABI of an out-of-tree model library built with microsimulation_b200/csrc/plugin.cuh.
 *
 * The reference's plugin mechanism is a host C++ class derived from ssim::cProcess (init() + handleMessage(),
 * inst/include/microsimulation.h:325-447) that the user compiles against the package headers, usually with
 * Rcpp::sourceCpp, and drives from an exported function of their own (inst/doc/examples.org:254-312: the loop
 * `for i < n: create_process; run_simulation; clear` and `wrap(report)`).  A model written against plugin.cuh is a .cu
 * file with the same handler bodies; MSIM_PLUGIN(name, Traits) exports, with C linkage:
 *
 *   int         <name>_run(const msim_plugin_call *call, msim_plugin_result *result);   0 or an MSIM_ERR_* code
 *   void        <name>_free(msim_plugin_result *result);                                  frees what _run allocated
 *   const char *<name>_last_error(void);
 *   int         <name>_n_values(void);        per-person outputs the model declares
 *   unsigned long <name>_consts_bytes(void);  sizeof the model's parameter struct
 *
 * There is no CPU path: without a CUDA device _run returns MSIM_ERR_NO_DEVICE.
 */
#ifndef MSIM_PLUGIN_H
#define MSIM_PLUGIN_H
#include "msim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MSIM_PLUGIN_SUBSTREAMS 0 /* RNGkind("user"): person i draws from RngStream substream first_person + i + 1 of the stream that starts at `seed` */
#define MSIM_PLUGIN_MT_STREAM 1  /* R's Mersenne-Twister stream, consumed person after person (set.seed before the loop) */

#define MSIM_PLUGIN_OUT_VALUES 8u /* with MSIM_OUT_REPORT (2u): which outputs to produce */

typedef struct msim_plugin_call {
  int32_t device;        /* CUDA device the plugin binds to at its first call */
  int32_t regime;        /* MSIM_PLUGIN_SUBSTREAMS or MSIM_PLUGIN_MT_STREAM */
  int64_t n;             /* persons */
  int64_t first_person;  /* substream regime: index of the first person (shards of one cohort) */
  double seed[6];        /* substream regime: the stream's start state (the package seed) */
  uint32_t *mt_state;    /* stream regime: R's 625 words (position, mt[624]); advanced past the last draw on return */
  int32_t max_draws;     /* stream regime: most uniforms one person may draw, 1..250 (more is MSIM_ERR_STREAM) */
  double mean_draws;     /* stream regime: expected uniforms per person (planning hint; 0 = plan for max_draws) */
  const void *consts;    /* the model's parameter struct */
  uint64_t consts_bytes; /* its size, checked against the model's */
  uint32_t outputs;      /* MSIM_OUT_REPORT | MSIM_PLUGIN_OUT_VALUES */
  int32_t n_state, n_event; /* EventReport dimensions (ages 0..100 by 1, as the built-in models) */
  double discount_rate;  /* EventReport utility discounting */
} msim_plugin_call;

typedef struct msim_plugin_result {
  msim_table values;  /* n rows x n_values columns (column-major), the model's env.value(k, v) */
  msim_report report; /* EventReport tables (pt, ut, events, prev) */
  float kernel_ms;    /* device time of the run (CUDA events) */
  int64_t launches;   /* kernels this plugin has launched so far */
} msim_plugin_result;

#ifdef __cplusplus
}
#endif
#endif
