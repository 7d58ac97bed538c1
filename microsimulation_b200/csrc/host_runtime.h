// Host-side runtime of the C ABI: the per-process context (device, stream, buffers that grow on demand),
// error reporting, page-locked host memory.  Everything in host_*.h lives in namespace msim_host; capi.cu holds
// the extern "C" entry points.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/msim.h"
#include "host_state.h"
#include "report_host.h"
#include "kernels.cuh"
#include "person_staged.cuh"
#include "summary_kernels.cuh"

namespace msim_host {
using namespace msim;

// ---------------------------------------------------------------- context

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

constexpr int STAGED_KERNELS = 26;  // instantiations of person_staged_kernel the host chooses from (host_person.h)

struct Ctx {
  bool inited = false;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  MrgJump *d_lo = nullptr, *d_hi = nullptr;
  uint32_t *d_mt_table = nullptr;  // jump-ahead polynomials of MT19937 (mt_jump.bin), null if the file is absent
  int mt_table_polys = 0;
  bool mt_parallel = true;         // msim_set_mt_parallel: tests switch the jump-ahead path off to compare
  DevBuf rowlog, prov_start, prov_end, prov_meta, tile_count, partial, dense, small, mt_words, len, starts, chain_exit, chain_count, chain_entry, chain_base, chain_gexit, chain_gcount, chain_gentry, chain_gbase, calib_out, calib_partial,
      calib_par, uniforms, first_end, first_tag, ov_fill, ov_tag, ov_start, ov_end, aux, code_in, code_out, sgeom,
      indiv, gsm_model, gsm_data, gsm_in, gsm_out, slices, fold_sync, carry_tiles;
  double last_small[16] = {0};  // the flag block as finish_run read it last
  // merge across the GPUs of a node inside the person-r launch (msim_peer_*, person_staged.cuh)
  struct Peers {
    int rank = 0, world = 0;
    int mode = 0;                    // 0 off, 1 merged when the launch ends, 2 merged one launch later (msim_peer_merge)
    void *local = nullptr;           // this rank's mailbox (cudaMalloc)
    void *mapped[8] = {nullptr};     // every rank's mailbox in this process's address space (own = local)
    unsigned epoch = 0;              // launches that sent a vector
    unsigned collected = 0;          // mode 2: the last epoch added up into `out`
    double *out = nullptr;           // mode 2: the merged vector (dense layout)
    int out_doubles = 0;
    } peers;
  PeerParams peers_last;        // as the last merging launch got them (msim_peer_flush)
  std::vector<int> peers_map;   // entry of the travelling vector -> cell of dense, as uploaded to peers_map_dev
  DevBuf peers_map_dev;
  int person_kernel = 0;  // 0 = staged schedule (person_staged.cuh), 1 = generic cohort kernel
  // share of the tiles the first wave of blocks takes, per kernel instantiation of the staged schedule (host_person.h)
  double wave_share[STAGED_KERNELS];  // (0.5 each to begin with: Ctx())
  Ctx() { for (double &f : wave_share) f = 0.5; }
  double wave_share_fixed = 0;  // msim_set_wave_share: > 0 switches the adaptation off
  int last_wave_slot = -1;      // of the last staged launch (-1: one wave)
  double last_wave_share = 0;
  bool last_wave_adapt = false;
  bool last_staged = false;
  double plan_rows = 1.25, plan_extra_rows = 0.26;  // row-log buffer planning (msim_set_rowlog_plan)
  int64_t last_ov_cap = 0, last_ov_rows = 0;
  HostRngState rng;
  // last cohort run
  int64_t last_n_rows = 0, last_capacity = 0;
  ReportGeom last_geom;
  unsigned last_outputs = 0;
  bool last_simple_names = false;
  int64_t want_capacity = 0;
  int64_t last_n = 0, last_first = 0;
  int last_max_rows = 0;
  msim_report_options report_opts;   // msim_set_report_options: the EventReport attached to person-r (and plugin) runs
  bool have_report_opts = false;
  int last_calib_sets = 0;  // parameter sets of the last calibration run (rows of calib_out)
  int last_cells = 0;  // dense report cells of the last run (0 when no report was requested)
  bool have_last = false;
  float last_ms = 0.f;
  bool timing_pending = false;
  int64_t launches = 0;
  // pinned host cache (one buffer) so steady-state row-log fetches do not re-pin
  void *pin_cache = nullptr;
  size_t pin_cache_bytes = 0;
  std::unordered_map<void *, size_t> pinned;  // live pinned allocations -> bytes
};

static Ctx g;  // one context per process (this header is included by capi.cu only)
static thread_local std::string g_err;

inline int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(MSIM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
  } while (0)

inline int ensure(DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return MSIM_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.bytes = 0;
  size_t want = bytes < 256 ? 256 : bytes;
  CK(cudaMalloc(&b.p, want));
  b.bytes = want;
  return MSIM_OK;
}

// mt_jump.bin (written by microsimulation_b200/mt_jump.py at build time) sits next to the shared library
inline void load_mt_jump_table() {
  Dl_info info;
  if (!dladdr((const void *)&load_mt_jump_table, &info) || !info.dli_fname) return;
  std::string path(info.dli_fname);
  const size_t slash = path.rfind('/');
  path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/mt_jump.bin";
  if (const char *env = getenv("MSIM_MT_JUMP_TABLE")) path = env;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return;
  char magic[4];
  uint32_t hdr[3];
  std::vector<uint32_t> tab;
  bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "MTJ1", 4) == 0 && fread(hdr, 4, 3, f) == 3 &&
            hdr[0] == (uint32_t)MT_SEGMENT && hdr[2] == (uint32_t)MT_POLY_WORDS && hdr[1] >= 1 && hdr[1] < 65536;
  if (ok) {
    tab.resize((size_t)hdr[1] * MT_POLY_WORDS);
    ok = fread(tab.data(), 4, tab.size(), f) == tab.size();
  }
  fclose(f);
  if (!ok) return;
  if (cudaMalloc(&g.d_mt_table, tab.size() * 4) != cudaSuccess) {
    g.d_mt_table = nullptr;
    return;
  }
  if (cudaMemcpy(g.d_mt_table, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(g.d_mt_table);
    g.d_mt_table = nullptr;
    return;
  }
  g.mt_table_polys = (int)hdr[1];
}

// The stream behind the state in words[0..624): words [624, total).  Long streams run as parallel segments started by
// jump-ahead (kernels.cuh); short ones, or all of them without the table, as one block.
inline int mt_generate(uint32_t *words, int64_t total) {
  if (total <= 624) return MSIM_OK;
  int64_t segments = 1, mult = 1;
  if (g.d_mt_table && g.mt_parallel && total >= 3 * MT_SEGMENT) {
    // a jump costs about as much as generating 1/19 of a table segment: with P segments of mult table segments each,
    // time ~ P + 19 * mult, so mult ~ sqrt(table segments / 19); longer segments also reach further with one table
    const double tab = (double)total / (double)MT_SEGMENT;
    mult = std::max<int64_t>(1, (int64_t)(sqrt(tab / 19.0) + 0.5));
    while ((total + mult * MT_SEGMENT - 1) / (mult * MT_SEGMENT) - 1 > g.mt_table_polys / mult && mult < 64) mult++;
    segments = std::min<int64_t>((total + mult * MT_SEGMENT - 1) / (mult * MT_SEGMENT), (int64_t)(g.mt_table_polys / mult) + 1);
  }
  if (segments <= 1) {
    mt_segment_kernel<<<1, 256, 0, g.stream>>>(words, total, 0, total);
    CK(cudaGetLastError());
    g.launches++;
    return MSIM_OK;
  }
  const int64_t seg = mult * MT_SEGMENT;
  // the words a jump reads, then the start states of segments 1.., then all segments side by side
  CK(cudaMemset2DAsync(words + seg, sizeof(uint32_t) * seg, 0, sizeof(uint32_t) * 624, (size_t)(segments - 1), g.stream));
  mt_segment_kernel<<<1, 256, 0, g.stream>>>(words, MT_JUMP_INPUT, 0, MT_JUMP_INPUT);
  CK(cudaGetLastError());
  mt_jump_kernel<<<dim3((unsigned)(segments - 1), MT_JUMP_SPLIT), 160, 0, g.stream>>>(words, g.d_mt_table, (int)mult);
  CK(cudaGetLastError());
  mt_segment_kernel<<<(unsigned)segments, 256, 0, g.stream>>>(words, total, MT_JUMP_INPUT - 624, seg);
  CK(cudaGetLastError());
  g.launches += 3;
  return MSIM_OK;
}

inline int ensure_init() {
  if (g.inited) return MSIM_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(MSIM_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
  if (g.device >= ndev) return fail(MSIM_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(g.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, g.device));
  g.sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
  g.stream = g.own_stream;
  CK(cudaEventCreate(&g.ev0));
  CK(cudaEventCreate(&g.ev1));
  // per-thread start tables: lo[j] = M^j, hi[j] = M^(JUMP_TAB*j)
  std::vector<MrgJump> lo(JUMP_TAB), hi(JUMP_TAB);
  const MrgJump &M = mrg_substream_jump();
  lo[0] = mrg_jump_identity();
  for (int j = 1; j < JUMP_TAB; j++) lo[j] = mrg_jump_mul(M, lo[j - 1]);
  MrgJump Mt = mrg_jump_mul(M, lo[JUMP_TAB - 1]);  // M^JUMP_TAB
  hi[0] = mrg_jump_identity();
  for (int j = 1; j < JUMP_TAB; j++) hi[j] = mrg_jump_mul(Mt, hi[j - 1]);
  CK(cudaMalloc(&g.d_lo, sizeof(MrgJump) * JUMP_TAB));
  CK(cudaMalloc(&g.d_hi, sizeof(MrgJump) * JUMP_TAB));
  CK(cudaMemcpy(g.d_lo, lo.data(), sizeof(MrgJump) * JUMP_TAB, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_hi, hi.data(), sizeof(MrgJump) * JUMP_TAB, cudaMemcpyHostToDevice));
  load_mt_jump_table();  // optional: without it the stream is generated by one block
  g.inited = true;
  return MSIM_OK;
}

// substream state of person `first_person` of the stream that starts at `state`
inline void stream_start(const uint32_t state[6], int64_t first_person, int64_t n_threads, StreamStart *st) {
  Mrg32k3a b;
  memcpy(b.s, state, sizeof b.s);
  b.jump(mrg_jump_pow(mrg_substream_jump(), (uint64_t)first_person + 1));
  memcpy(st->base, b.s, sizeof b.s);
  st->lo = g.d_lo;
  st->hi = g.d_hi;
  st->stride = mrg_jump_pow(mrg_substream_jump(), (uint64_t)n_threads);
}

// ---------------------------------------------------------------- host memory

inline int pinned_alloc(size_t bytes, double **out) {
  if (bytes == 0) bytes = 8;
  if (g.pin_cache && g.pin_cache_bytes >= bytes) {
    *out = (double *)g.pin_cache;
    g.pinned[g.pin_cache] = g.pin_cache_bytes;
    g.pin_cache = nullptr;
    g.pin_cache_bytes = 0;
    return MSIM_OK;
  }
  void *p = nullptr;
  CK(cudaMallocHost(&p, bytes));
  g.pinned[p] = bytes;
  *out = (double *)p;
  return MSIM_OK;
}

inline void host_free(double *p) {
  if (!p) return;
  auto it = g.pinned.find(p);
  if (it == g.pinned.end()) {
    free(p);
    return;
  }
  size_t bytes = it->second;
  g.pinned.erase(it);
  if (bytes > g.pin_cache_bytes) {  // keep the largest buffer for the next call
    if (g.pin_cache) cudaFreeHost(g.pin_cache);
    g.pin_cache = p;
    g.pin_cache_bytes = bytes;
  } else {
    cudaFreeHost(p);
  }
}

// The geometry of the EventReport attached to a run: the built models' partition {0, 1, ..., 100, 1e6}
// (illness-death.cpp:81-85) unless msim_set_report_options chose another.
inline int person_geom(int n_state, int n_event, double discount_rate, ReportGeom *out) {
  if (!g.have_report_opts) {
    *out = ages_geom(n_state, n_event, discount_rate);
    return MSIM_OK;
  }
  const msim_report_options &o = g.report_opts;
  if (!make_report_geom_partition(out, n_state, n_event, o.start, o.finish, o.delta, o.max_time, discount_rate, o.start_report_age,
                                  o.output_utilities))
    return fail(MSIM_ERR_ARG, "report options: the partition must ascend (delta > 0, finish >= start, maxTime > finish) and hold at most 127 points");
  return MSIM_OK;
}

inline ReportGeom geom_from_dims(const msim_dense_dims *dims, double discount_rate) {
  ReportGeom G;
  if (person_geom(dims->n_state, dims->n_event, discount_rate, &G) != MSIM_OK) G = ages_geom(dims->n_state, dims->n_event, discount_rate);
  return G;
}

// `small` buffer layout (8-byte cells): [9] n_rows (u64), [10] error (int), [11] end_word (i64), [12] overflow rows (u64)
struct SmallView {
  unsigned long long *wave_clock;  // [3] (person_staged.cuh, StagedParams::wave_clock)
  unsigned *wave_sync;             // [2] (StagedParams::wave_sync)
  unsigned long long *n_rows;
  int *error;
  int64_t *end_word;
  unsigned long long *ov_count;
};
inline SmallView small_view() {
  SmallView v;
  double *b = (double *)g.small.p;
  v.wave_clock = (unsigned long long *)b;
  v.wave_sync = (unsigned *)(b + 3);
  v.n_rows = (unsigned long long *)(b + 9);
  v.error = (int *)(b + 10);
  v.end_word = (int64_t *)(b + 11);
  v.ov_count = (unsigned long long *)(b + 12);
  return v;
}

}  // namespace msim_host
