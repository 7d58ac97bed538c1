"""Multi-GPU sharding of the substream-regime models (one process per GPU).

Persons are independent and person i is a pure function of (seed, i), so a
cohort shards by contiguous person ranges with no data-path exchange; rank r
jumps straight to its first substream (A^(2^76))^(first+1) on the host.  The
only collective is ONE all-reduce(sum, f64) over the accumulator vector (dense
report cells + per-kind counts): every cell is a sum and integer counts are
exact in f64, so the merged tables equal a single-GPU run's (FP sums to
rounding).  Row logs stay sharded, in person order within and across ranks.

The reference's own parallel wrappers (R/calibSim.R:16-39, README.org:347-352)
give chunk k a different RngStream *stream*, which changes the random numbers;
this keeps the mc.cores = 1 numbering instead (end).
"""
import ctypes as C

import numpy as np

from . import api
from ._lib import MSIM_OUT_COUNTS, MSIM_OUT_REPORT, MSIM_OUT_ROWLOG, DenseDims  # noqa: F401


def shard_range(n, rank, world_size):
    """Contiguous [first, first+count) of n persons for `rank`; earlier ranks take the remainder."""
    base, rem = divmod(int(n), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class _CudaBuffer:
    """Zero-copy view of a device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def person_dims(with_report):
    cells = 0
    d = DenseDims(8, 8, 102, 0)
    if with_report:
        sb = 8 * 102
        cells = 4 * sb + 8 * 8 * 102  # pt, ut, dstart, noedge + events (csrc/report.cuh)
    d.n_doubles = cells + 9
    return d


def _gpu_compute(seed, first, count, outputs, discount_rate):
    """Runs the shard on this process's GPU; returns a torch CUDA tensor aliasing the library's accumulators."""
    import torch
    res = api.person_run_device(seed, first, count, outputs, discount_rate)
    ptr, n = api.dense_device_ptr()
    t = torch.as_tensor(_CudaBuffer(ptr, n), device="cuda")
    return t, res.dims, res.n_rows


_peer_group_ready = False


def enable_peer_merge(group=None, mode=1):
    """Sets up the merge INSIDE the person-r launch (csrc/person_staged.cuh, staged_peer_merge): every rank allocates a
    mailbox, the CUDA IPC handles go round once (all_gather_object), every rank maps the others' mailboxes over
    NVLink.  After this, person-r runs with a report leave the sum over all ranks in their dense vector — no
    all-reduce to enqueue.  mode 2: the sum of run k is formed during run k + 1 or by api.peer_flush(), and is read from
    api.peer_result_ptr() (api.peer_merge switches between the modes afterwards).  One node, at most 8 ranks, one
    process per GPU; returns False (and changes nothing) where that does not hold or peer memory cannot be mapped."""
    global _peer_group_ready
    import torch.distributed as dist
    if not dist.is_initialized():
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world < 2 or world > 8:
        return False
    ok = True
    handle = b""
    try:
        handle = api.peer_export(world)
    except RuntimeError:
        ok = False
    handles = [None] * world
    dist.all_gather_object(handles, handle if ok else None, group=group)
    if any(h is None for h in handles):
        ok = False
    if ok:
        try:
            api.peer_import(rank, world, handles)
        except RuntimeError:
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)  # also the barrier: every mailbox is zeroed and mapped from here on
    if not all(flags):
        try:
            api.peer_close()
        except RuntimeError:
            pass
        return False
    api.peer_merge(mode)
    _peer_group_ready = True
    return True


def disable_peer_merge():
    global _peer_group_ready
    if _peer_group_ready:
        api.peer_flush()
        api.peer_merge(0)
    _peer_group_ready = False


def person_simulation_sharded(n, seed=(12345,) * 6, outputs=MSIM_OUT_REPORT | MSIM_OUT_COUNTS, discount_rate=0.0,
                              first_person=0, group=None, compute=None):
    """callPersonSimulation over all ranks of `group`.

    Returns (report tables or None, counts[9], rows on this rank).  `compute`
    (tests only) replaces the GPU shard runner with another function of the
    same signature returning (tensor, dims, n_rows).
    """
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_range(n, rank, world)
    acc, dims, n_rows = (compute or _gpu_compute)(seed, first_person + first, count, outputs, discount_rate)
    merged_in_launch = _peer_group_ready and compute is None and bool(outputs & MSIM_OUT_REPORT)
    if world > 1 and not merged_in_launch:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)  # the single collective
    host = acc.detach().cpu().numpy()
    counts = host[-9:].copy()
    report = api.dense_to_report(host, dims, discount_rate) if outputs & MSIM_OUT_REPORT else None
    return report, counts, n_rows


def calibration_grid(world_size, n_sets):
    """How `world_size` ranks split a sweep: (groups of sets, groups of persons).  Sets are split first (every GPU
    keeps whole sets and a full grid of blocks per set), persons only by what is left over — the 2-D split of
    SURVEY.md 8(e)."""
    gs = 1
    for d in range(1, world_size + 1):
        if world_size % d == 0 and d <= max(1, n_sets):
            gs = d
    return gs, world_size // gs


def calibration_shard(rank, world_size, n_sets, n):
    """This rank's (set_lo, set_hi, first_person, persons)."""
    gs, gp = calibration_grid(world_size, n_sets)
    set_lo, n_here = shard_range(n_sets, rank % gs, gs)
    first, count = shard_range(n, rank // gs, gp)
    return set_lo, set_lo + n_here, first, count


def _gpu_calibration(seed, runpars, set_lo, set_hi, first, count):
    """Runs the shard on this process's GPU; returns a torch CUDA tensor aliasing the library's result vector."""
    import torch
    ptr, nd = api.calibration_sweep_device(seed, runpars, count, first, set_lo, set_hi)
    return torch.as_tensor(_CudaBuffer(ptr, nd), device="cuda")


def calibration_sweep_sharded(seed, runpars, n, group=None, compute=None, unpack=True):
    """BASELINE configs[4] (R/calibSim.R:10-40): n_sets parameter vectors x n persons over all ranks of `group`,
    every set on the same substreams (common random numbers) and — unlike mclapply chunks, which take a new stream
    each — with the same random numbers as a single-GPU run.  Ranks split the sets first and the persons second;
    every cell of the result vector (48 doubles per set) is a sum, so ONE all-reduce on the device merges both
    kinds of shard.  `compute` (tests only) replaces the GPU shard runner: f(seed, runpars, set_lo, set_hi,
    first_person, persons) -> tensor of 48 * n_sets doubles."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    runpars = np.ascontiguousarray(runpars, dtype=np.float64).reshape(-1, 6)
    set_lo, set_hi, first, count = calibration_shard(rank, world, runpars.shape[0], n)
    vec = (compute or _gpu_calibration)(seed, runpars, set_lo, set_hi, first, count)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)  # the single collective
    if not unpack:
        return vec
    host = vec.detach().cpu().numpy()
    return api.calibration_unpack(host)
