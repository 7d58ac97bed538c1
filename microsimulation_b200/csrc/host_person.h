// Host side of the person-r entry points: the staged kernel (person_staged.cuh) with its row-log passes and
// report slices, or the generic cohort kernel (msim_set_person_schedule).
#pragma once
#include "host_cohort.h"

namespace msim_host {
using namespace msim;

// ---- person-r

// the bandwidth-bound passes that turn first rows + further rows into the final columns
int staged_rowpasses(int64_t n, int64_t first_person, int64_t capacity, bool count_first) {
  RowPassParams r;
  memset(&r, 0, sizeof r);
  SmallView sv = small_view();
  r.first_end = (const double *)g.first_end.p;
  r.first_tag = (const unsigned short *)g.first_tag.p;
  r.ov_tag = (const unsigned *)g.ov_tag.p;
  r.ov_start = (const double *)g.ov_start.p;
  r.ov_end = (const double *)g.ov_end.p;
  r.ov_fill = (const unsigned *)g.ov_fill.p;
  r.max_fill = sv.ov_count;
  r.ov_cap = (int)g.last_ov_cap;
  r.n = n;
  r.first_person = first_person;
  r.chunk_rows = (unsigned long long *)g.partial.p;
  r.n_rows = sv.n_rows;
  r.n_chunks = (n + ROW_CHUNK - 1) / ROW_CHUNK;
  r.rowlog = (double *)g.rowlog.p;
  r.capacity = capacity;
  if (r.n_chunks == 0) return MSIM_OK;
  if (count_first) {
    rowpass_count_kernel<<<(unsigned)r.n_chunks, 256, 0, g.stream>>>(r);
    CK(cudaGetLastError());
    rowpass_scan_kernel<<<1, 1024, 0, g.stream>>>(r);
    CK(cudaGetLastError());
    g.launches += 2;
  }
  rowpass_final_kernel<<<(unsigned)r.n_chunks, 256, 0, g.stream>>>(r);
  CK(cudaGetLastError());
  g.launches++;
  return MSIM_OK;
}

// ---- mailboxes for the in-launch merge across GPUs (person_staged.cuh, staged_peer_merge)
constexpr int PEER_CAP = 8192;  // values per slot, 16 bytes each (person-r's reachable report cells + counts: 2763)

inline size_t peer_mailbox_bytes(int world) { return 16 * 2 * (size_t)world * PEER_CAP; }

int peer_close() {
  for (int r = 0; r < 8; r++) {
    if (g.peers.mapped[r] && g.peers.mapped[r] != g.peers.local) cudaIpcCloseMemHandle(g.peers.mapped[r]);
    g.peers.mapped[r] = nullptr;
  }
  if (g.peers.local) cudaFree(g.peers.local);
  if (g.peers.out) cudaFree(g.peers.out);
  g.peers = Ctx::Peers();
  return MSIM_OK;
}

int peer_export(int world, void *handle_out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (world < 1 || world > PEER_MAX || !handle_out) return fail(MSIM_ERR_ARG, "peer merge: 1 to 8 ranks on one node");
  peer_close();
  CK(cudaMalloc(&g.peers.local, peer_mailbox_bytes(world)));
  CK(cudaMemset(g.peers.local, 0, peer_mailbox_bytes(world)));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, g.peers.local));
  static_assert(sizeof h == MSIM_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof h);
  g.peers.world = world;
  return MSIM_OK;
}

int peer_import(int rank, int world, const void *handles) {
  if (!g.peers.local || world != g.peers.world) return fail(MSIM_ERR_STATE, "peer merge: msim_peer_export first, with the same number of ranks");
  if (rank < 0 || rank >= world || !handles) return fail(MSIM_ERR_ARG, "peer merge: bad rank");
  for (int r = 0; r < world; r++) {
    if (r == rank) {
      g.peers.mapped[r] = g.peers.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * MSIM_PEER_HANDLE_BYTES, sizeof h);
    CK(cudaIpcOpenMemHandle(&g.peers.mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  g.peers.rank = rank;
  g.peers.epoch = 0;
  return MSIM_OK;
}

// mode 2: the vectors of the last launch, which no later launch has added up yet
int peer_flush() {
  if (g.peers.mode != 2 || g.peers.world < 2 || g.peers.epoch == g.peers.collected) return MSIM_OK;
  PeerCollectParams q;
  memset(&q, 0, sizeof q);
  q.peer = g.peers_last;
  q.peer.epoch = g.peers.epoch;
  q.error = small_view().error;
  peer_collect_kernel<<<1, 256, 0, g.stream>>>(q);
  CK(cudaGetLastError());
  g.launches++;
  g.peers.collected = g.peers.epoch;
  return MSIM_OK;
}

int person_enqueue_staged(const msim_person_shard *sh, int64_t capacity, int64_t ov_cap_override) {
  int rc;
  if (sh->n >= (int64_t)4294967296ll) return fail(MSIM_ERR_ARG, "a shard holds at most 2^32 - 1 persons");
  StagedParams p;
  memset(&p, 0, sizeof p);
  p.first_person = sh->first_person;
  p.n = sh->n;
  p.outputs = sh->outputs & 15u;
  if ((p.outputs & OUT_INDIV) && !(p.outputs & OUT_REPORT)) return fail(MSIM_ERR_ARG, "MSIM_OUT_INDIV needs MSIM_OUT_REPORT");
  p.consts = person_r::make_consts();
  if ((rc = person_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate, &p.geom))) return rc;
  if (!make_report_compact(p.geom, &person_r::REACHABLE_PAIRS[0][0], person_r::N_REACHABLE_PAIRS, &p.compact))
    return fail(MSIM_ERR_ARG, "report geometry too large for the block tables");
  size_t smem = sizeof(WarpQueues) * SWARPS;
  if (p.outputs & OUT_ROWLOG) smem += sizeof(WarpRows) * SWARPS;
  if (p.outputs & OUT_REPORT) smem += staged_counter_bytes(p.geom, p.compact);  // the counters; FP sums go to per-block global slices
  typedef void (*StagedKernel)(const StagedParams);
  // one instantiation per combination of outputs (person_staged.cuh); the general one, which reads the outputs at
  // run time, only where individual totals are kept
  static const StagedKernel kernels[8] = {person_staged_kernel<0>, person_staged_kernel<1>, person_staged_kernel<2>,
                                          person_staged_kernel<3>, person_staged_kernel<4>, person_staged_kernel<5>,
                                          person_staged_kernel<6>, person_staged_kernel<7>};
  // with the merge across GPUs in the launch's tail (msim_peer_merge): instantiations of their own
  static const StagedKernel peer_kernels[4] = {person_staged_kernel<2, true>, person_staged_kernel<3, true>,
                                               person_staged_kernel<6, true>, person_staged_kernel<7, true>};
  // with a report on the built models' partition {0, 1, ..., 100, 1e6} (the default): instantiations without the
  // search for other partitions in their code
  static const StagedKernel unit_kernels[4] = {person_staged_kernel<2, false, true>, person_staged_kernel<3, false, true>,
                                               person_staged_kernel<6, false, true>, person_staged_kernel<7, false, true>};
  static const StagedKernel unit_peer_kernels[4] = {person_staged_kernel<2, true, true>, person_staged_kernel<3, true, true>,
                                                    person_staged_kernel<6, true, true>, person_staged_kernel<7, true, true>};
  const bool with_peers = g.peers.mode && g.peers.world > 1 && (p.outputs & OUT_REPORT);
  if (with_peers && (p.outputs & OUT_INDIV)) return fail(MSIM_ERR_ARG, "peer merge: not with MSIM_OUT_INDIV (individual totals stay on their rank)");
  const bool unit = (p.outputs & OUT_REPORT) && !(p.outputs & OUT_INDIV) && p.geom.unit;
  const unsigned rep_idx = ((p.outputs & 4u) >> 1) | (p.outputs & 1u);  // among the four combinations with a report
  const StagedKernel kernel = with_peers ? (unit ? unit_peer_kernels : peer_kernels)[rep_idx]
                              : (p.outputs & OUT_INDIV) ? person_staged_kernel<OUT_RUNTIME>
                              : unit ? unit_kernels[rep_idx] : kernels[p.outputs & 7u];
  // blocks per SM for this (kernel, shared memory) pair: asked once, the answer does not change
  static size_t occ_smem[STAGED_KERNELS] = {0};
  static int occ_per_sm[STAGED_KERNELS] = {0};
  const unsigned slot = (with_peers ? 9u + rep_idx : (p.outputs & OUT_INDIV) ? 8u : (p.outputs & 7u)) + (unit ? 13u : 0u);
  int per_sm = occ_per_sm[slot];
  if (per_sm == 0 || occ_smem[slot] != smem) {
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // as much of the SM's memory as shared memory as there is: the kernel reads no global data worth caching
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SBLOCK, smem));
    if (per_sm < 1) return fail(MSIM_ERR_CUDA, "staged kernel does not fit on an SM");
    occ_per_sm[slot] = per_sm;
    occ_smem[slot] = smem;
  }
  const int64_t tiles = (sh->n + 31) / 32;
  int64_t gmax = (int64_t)per_sm * g.sm_count;
  if (gmax > JUMP_TAB * (int64_t)JUMP_TAB / SBLOCK) gmax = JUMP_TAB * (int64_t)JUMP_TAB / SBLOCK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(gmax, (tiles + SWARPS - 1) / SWARPS));
  // a lane tallies event kinds in 16-bit fields (person_staged.cuh): at most 65535 persons per lane
  // (persons are handed to lanes by queue position, not evenly: keep a factor of two in hand)
  if (sh->n / ((int64_t)grid * SBLOCK) >= 32000) return fail(MSIM_ERR_ARG, "shard too large for this device: split it");
  uint32_t state[6];
  seed_to_u32(sh->seed, state);
  // Two blocks per SM: the blocks that arrive first on their SMs (the first sm_count of the grid) are favoured by the
  // warp schedulers, so they get the larger share of the tiles — as much larger as makes both waves of blocks finish
  // together (wave_feedback below; person_staged.cuh).  Results do not depend on the share.
  const int64_t tiles_per_warp = tiles / ((int64_t)grid * SWARPS);
  const bool two_waves = per_sm == 2 && grid == 2 * g.sm_count && tiles_per_warp >= WAVE_MIN_TILES;
  p.wave_blocks = two_waves ? g.sm_count : grid;
  p.wave_split = tiles;
  g.last_wave_slot = -1;
  g.last_wave_share = 0;
  if (two_waves) {
    const double share = g.wave_share_fixed > 0 ? g.wave_share_fixed : g.wave_share[slot];
    p.wave_split = std::min<int64_t>(tiles, std::max<int64_t>(0, (int64_t)(share * (double)tiles + 0.5)));
    Mrg32k3a b;
    memcpy(b.s, state, sizeof b.s);
    b.jump(mrg_jump_pow(mrg_substream_jump(), (uint64_t)sh->first_person + 1 + 32ull * (uint64_t)p.wave_split));
    memcpy(p.base_b, b.s, sizeof b.s);
    g.last_wave_slot = (int)slot;
    g.last_wave_share = (double)p.wave_split / (double)tiles;
    g.last_wave_adapt = g.wave_share_fixed <= 0 && tiles_per_warp >= WAVE_ADAPT_TILES;
  }
  stream_start(state, sh->first_person, (int64_t)p.wave_blocks * SBLOCK, &p.st);
  if ((rc = ensure(g.small, 16 * sizeof(double)))) return rc;
  CK(cudaMemsetAsync(g.small.p, 0, 16 * sizeof(double), g.stream));
  SmallView sv = small_view();
  p.error = sv.error;
  p.wave_clock = two_waves ? sv.wave_clock : nullptr;
  p.wave_sync = sv.wave_sync;
  const int cells = (p.outputs & OUT_REPORT) ? p.geom.total : 0;
  if ((rc = ensure(g.dense, sizeof(double) * (cells + N_COUNTS)))) return rc;
  CK(cudaMemsetAsync(g.dense.p, 0, sizeof(double) * (cells + N_COUNTS), g.stream));
  p.dense = (double *)g.dense.p;
  p.counts = p.dense + cells;
  g.last_cells = cells;
  int64_t ov_cap = 0;
  if (p.outputs & OUT_ROWLOG) {
    const size_t n1 = (size_t)std::max<int64_t>(sh->n, 1);
    // rows beyond a person's first: 0.232 per person at the reference's parameters, at most 4; each chunk of
    // ROW_CHUNK persons owns a region planned with a wide margin (a chunk that needs more repeats the run)
    const int64_t n_chunks = (sh->n + ROW_CHUNK - 1) / ROW_CHUNK;
    ov_cap = ov_cap_override > 0 ? ov_cap_override
                                 : std::min<int64_t>(std::max<int64_t>((int64_t)(g.plan_extra_rows * 1.6 * ROW_CHUNK) + 96, ROW_OV_MIN),
                                                     (int64_t)ROW_CHUNK * (person_r::MAX_ROWS - 1));
    const size_t ov_n = (size_t)std::max<int64_t>(n_chunks, 1) * (size_t)ov_cap;
    if ((rc = ensure(g.first_end, sizeof(double) * n1))) return rc;
    if ((rc = ensure(g.first_tag, sizeof(unsigned short) * n1))) return rc;
    if ((rc = ensure(g.ov_fill, sizeof(unsigned) * (size_t)std::max<int64_t>(n_chunks, 1)))) return rc;
    if ((rc = ensure(g.ov_tag, sizeof(unsigned) * ov_n))) return rc;
    if ((rc = ensure(g.ov_start, sizeof(double) * ov_n))) return rc;
    if ((rc = ensure(g.ov_end, sizeof(double) * ov_n))) return rc;
    CK(cudaMemsetAsync(g.ov_fill.p, 0, sizeof(unsigned) * (size_t)std::max<int64_t>(n_chunks, 1), g.stream));
    if ((rc = ensure(g.partial, sizeof(unsigned long long) * (size_t)((sh->n + ROW_CHUNK - 1) / ROW_CHUNK + 1)))) return rc;
    if ((rc = ensure(g.rowlog, sizeof(double) * 5 * (size_t)std::max<int64_t>(capacity, 1)))) return rc;
    p.first_end = (double *)g.first_end.p;
    p.first_tag = (unsigned short *)g.first_tag.p;
    p.ov_tag = (unsigned *)g.ov_tag.p;
    p.ov_start = (double *)g.ov_start.p;
    p.ov_end = (double *)g.ov_end.p;
    p.ov_fill = (unsigned *)g.ov_fill.p;
    p.ov_cap = (int)ov_cap;
  }
  g.last_ov_cap = ov_cap;
  if (p.outputs & OUT_INDIV) {
    if ((rc = ensure(g.indiv, sizeof(double) * (size_t)std::max<int64_t>(sh->n, 1)))) return rc;
    p.indiv = (double *)g.indiv.p;
  }
  if (p.outputs & OUT_REPORT) {
    // per-block slices of the FP sums (pt, ut): zeroed by their blocks, added up by the launch itself
    const size_t bytes = sizeof(double) * (size_t)grid * staged_slice_doubles(p.geom, p.compact);
    if ((rc = ensure(g.slices, bytes))) return rc;
    p.slices = (double *)g.slices.p;
    const size_t sync_bytes = sizeof(unsigned) * (size_t)((grid + FOLD_GROUP - 1) / FOLD_GROUP + 2);
    if (g.fold_sync.bytes < sync_bytes || !g.fold_sync.p) {  // arrival counters: zero once, every launch leaves them zero
      if ((rc = ensure(g.fold_sync, sync_bytes))) return rc;
      CK(cudaMemsetAsync(g.fold_sync.p, 0, g.fold_sync.bytes, g.stream));
    }
    p.fold_sync = (unsigned *)g.fold_sync.p;
  }
  if (with_peers) {
    // the launch also merges the dense vector with the other ranks' (person_staged.cuh, staged_peer_merge): every
    // rank must make the same sequence of such calls
    if (p.n == 0) return fail(MSIM_ERR_ARG, "peer merge: every rank needs at least one person");
    const int n_compact = (4 * p.compact.n_states + p.compact.n_pairs) * p.geom.n_bin + N_COUNTS;  // person_staged.cuh, peer_cell
    if (n_compact > PEER_CAP) return fail(MSIM_ERR_ARG, "peer merge: report too large for the mailbox");
    if (g.peers.mode == 2 && g.peers.out_doubles != cells + N_COUNTS) {
      if (g.peers.epoch != g.peers.collected) return fail(MSIM_ERR_STATE, "peer merge: msim_peer_flush before a run with another report geometry");
      if (g.peers.out) cudaFree(g.peers.out);
      g.peers.out = nullptr;
      CK(cudaMalloc((void **)&g.peers.out, sizeof(double) * (cells + N_COUNTS)));
      CK(cudaMemsetAsync(g.peers.out, 0, sizeof(double) * (cells + N_COUNTS), g.stream));  // cells the model cannot reach stay 0
      g.peers.out_doubles = cells + N_COUNTS;
    }
    {  // which cell of dense each entry of the travelling vector is (person_staged.cuh, peer_cell_of)
      std::vector<int> map((size_t)n_compact);
      for (int c = 0; c < n_compact; c++) map[c] = peer_cell_of(p.geom, p.compact, c);
      if (map != g.peers_map) {
        if (g.peers.epoch != g.peers.collected) return fail(MSIM_ERR_STATE, "peer merge: msim_peer_flush before a run with another report geometry");
        if ((rc = ensure(g.peers_map_dev, sizeof(int) * map.size()))) return rc;
        CK(cudaMemcpyAsync(g.peers_map_dev.p, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice, g.stream));
        CK(cudaStreamSynchronize(g.stream));  // `map` is pageable host memory
        g.peers_map = map;
      }
      p.peer.cell = (const int *)g.peers_map_dev.p;
    }
    for (int r = 0; r < g.peers.world; r++) p.peer.mail[r] = (unsigned long long *)g.peers.mapped[r];
    p.peer.rank = g.peers.rank;
    p.peer.world = g.peers.world;
    p.peer.cap = PEER_CAP;
    p.peer.total = n_compact;
    p.peer.mode = g.peers.mode;
    p.peer.out = g.peers.out;
    p.peer.epoch = ++g.peers.epoch;
    if (g.peers.mode == 2) g.peers.collected = p.peer.epoch - 1;  // this launch adds up the previous one's vectors
    else g.peers.collected = p.peer.epoch;
    g.peers_last = p.peer;
  }
  CK(cudaEventRecord(g.ev0, g.stream));
  if (p.n > 0) {
    kernel<<<grid, SBLOCK, smem, g.stream>>>(p);
    CK(cudaGetLastError());
    g.launches++;
    if (p.outputs & OUT_ROWLOG) {
      if ((rc = staged_rowpasses(p.n, p.first_person, capacity, true))) return rc;
    }
  }
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  g.last_geom = p.geom;
  g.last_outputs = p.outputs;
  g.last_capacity = capacity;
  g.last_n = p.n;
  g.last_first = p.first_person;
  g.last_max_rows = person_r::MAX_ROWS;
  g.last_staged = true;
  g.have_last = true;
  return MSIM_OK;
}

int person_enqueue(const msim_person_shard *sh, int64_t capacity_override, int64_t ov_cap_override = 0) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if (!sh || sh->n < 0 || sh->first_person < 0) return fail(MSIM_ERR_ARG, "bad shard");
  if (!mrg_seed_ok(sh->seed)) return fail(MSIM_ERR_SEED, "illegal MRG32k3a seed");
  const int64_t capacity = capacity_override > 0 ? capacity_override : rowlog_capacity_guess(sh->n, person_r::MAX_ROWS, g.plan_rows);
  g.want_capacity = capacity;
  g.last_simple_names = false;
  if (g.person_kernel != 1) return person_enqueue_staged(sh, capacity, ov_cap_override);
  CohortParams<person_r::Consts> p;
  memset(&p, 0, sizeof p);
  p.first_person = sh->first_person;
  p.n = sh->n;
  if (sh->outputs & OUT_INDIV) return fail(MSIM_ERR_ARG, "individual totals (MSIM_OUT_INDIV) are produced by the staged schedule only");
  p.outputs = sh->outputs & 7u;
  p.consts = person_r::make_consts();
  if ((rc = person_geom(person_r::N_STATE, person_r::N_EVENT, sh->discount_rate, &p.geom))) return rc;
  uint32_t state[6];
  seed_to_u32(sh->seed, state);
  int grid;
  size_t smem;
  if ((rc = cohort_prepare<PersonTraits, false>(p, state, &grid, &smem))) return rc;
  CK(cudaEventRecord(g.ev0, g.stream));
  if ((rc = cohort_fire<PersonTraits, false>(p, grid, smem, capacity))) return rc;
  CK(cudaEventRecord(g.ev1, g.stream));
  g.timing_pending = true;
  return MSIM_OK;
}

int person_run(const msim_person_shard *sh) {
  int rc;
  if ((rc = person_enqueue(sh, 0))) return rc;
  int64_t rows = 0;
  if ((rc = finish_run(&rows))) return rc;
  if ((sh->outputs & OUT_ROWLOG) && g.last_staged && g.last_ov_rows > g.last_ov_cap) {
    // rare: more rows beyond the first than planned for — run again with the exact size
    if ((rc = person_enqueue(sh, rows > g.last_capacity ? rows : 0, g.last_ov_rows))) return rc;
    if ((rc = finish_run(&rows))) return rc;
  }
  if ((sh->outputs & OUT_ROWLOG) && rows > g.last_capacity) {  // rare: redo the expand pass with the exact size
    if ((rc = reexpand(rows))) return rc;
  }
  return MSIM_OK;
}

}  // namespace msim_host
