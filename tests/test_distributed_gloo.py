"""world_size = 2 over gloo on CPU: shard planning, the single all-reduce merge and
host finalisation of microsimulation_b200.distributed.  The per-shard compute is the
host emulation of the device engine (no GPU here); on the GPU box the same code path
runs with the CUDA shard runner (tests/test_gpu_parity.py, bench.py --gpus N)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions():
    from microsimulation_b200.distributed import shard_range
    for n in (0, 1, 7, 100, 10 ** 8 + 3):
        for w in (1, 2, 3, 8):
            nxt = 0
            for r in range(w):
                f, c = shard_range(n, r, w)
                assert f == nxt and c >= 0
                nxt = f + c
            assert nxt == n


def test_calibration_shards_cover_sets_and_persons_once():
    """The 2-D split of the sweep: every (set, person) belongs to exactly one rank; sets are split before persons."""
    from microsimulation_b200.distributed import calibration_grid, calibration_shard
    assert calibration_grid(8, 256) == (8, 1) and calibration_grid(8, 4) == (4, 2) and calibration_grid(8, 1) == (1, 8)
    assert calibration_grid(4, 3) == (2, 2) and calibration_grid(1, 256) == (1, 1) and calibration_grid(6, 5) == (3, 2)
    for world, n_sets, n in ((8, 256, 10 ** 6), (8, 4, 1001), (4, 3, 77), (2, 1, 501), (6, 5, 13), (3, 2, 0)):
        seen = np.zeros((n_sets, n), dtype=np.int32)
        for r in range(world):
            lo, hi, first, count = calibration_shard(r, world, n_sets, n)
            assert 0 <= lo <= hi <= n_sets and first >= 0 and first + count <= n
            seen[lo:hi, first:first + count] += 1
        assert np.all(seen == 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from microsimulation_b200 import distributed as D
    from oracle.pyoracle import PersonShard
    emu = C.CDLL(os.path.join(ROOT, "tests", "host_emulation", "libmsim_emu.so"))

    def compute(seed, first, count, outputs, discount_rate):
        from microsimulation_b200.api import _seed6
        sh = PersonShard((C.c_double * 6)(*_seed6(seed)), first, count, outputs, discount_rate)
        dims = D.person_dims(bool(outputs & 2))
        dense = np.zeros(dims.n_doubles)
        assert emu.emu_person_dense(C.byref(sh), C.c_int(3 + rank), dense.ctypes.data_as(C.POINTER(C.c_double)),
                                    C.c_int64(dims.n_doubles)) == 0
        return torch.from_numpy(dense), dims, -1

    rep, counts, _ = D.person_simulation_sharded(n, 12345, outputs=6, discount_rate=0.03, compute=compute)

    def calib(seed, runpars, set_lo, set_hi, first, count):
        """one vector of 48 doubles per set, every cell a sum (include/msim.h, msim_calibration_sweep_device)"""
        from oracle.pyoracle import Calibration
        from microsimulation_b200.api import _seed6
        vec = np.zeros(48 * len(runpars))
        for s_ in range(set_lo, set_hi):
            c = Calibration()
            assert emu.emu_calibration((C.c_double * 6)(*_seed6(seed)), C.c_int64(first), C.c_int64(count),
                                       (C.c_double * 6)(*runpars[s_]), C.c_int(2), C.byref(c)) == 0
            vec[48 * s_:48 * s_ + 40] = c.occupancy[:]
            vec[48 * s_ + 40:48 * s_ + 40 + c.n_time_at_risk] = c.time_at_risk[:c.n_time_at_risk]
            vec[48 * s_ + 44:48 * s_ + 44 + c.n_time_at_risk] = 1.0
        return torch.from_numpy(vec)

    pars = [[4, 0.5, 0.05, 10, 3, 0.5], [4, 0.0, 0.3, 8, 3, 0.2], [3.5, 0.4, 0.1, 9, 3, 0.4]]
    cal = D.calibration_sweep_sharded(10, pars, 501, compute=calib)        # 3 sets over 2 ranks: split by sets
    cal1 = D.calibration_sweep_sharded(10, pars[:1], 501, compute=calib)   # 1 set over 2 ranks: split by persons
    cal = cal + cal1
    if rank == 0:
        q.put((rep, counts, cal))
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_run(oracle, emu):
    n = 6001  # odd: ranks get 3001 / 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    rep, counts, cal = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from conftest import assert_report_close
    ref = oracle.person_run(12345, 0, n, rowlog=False, report=True, discount_rate=0.03)
    assert np.array_equal(counts[:8], ref["counts"][:8])
    assert abs(counts[8] - ref["counts"][8]) <= 1e-12 * ref["counts"][8]
    assert_report_close(rep, ref["report"], rtol=1e-12)
    pars = [[4, 0.5, 0.05, 10, 3, 0.5], [4, 0.0, 0.3, 8, 3, 0.2], [3.5, 0.4, 0.1, 9, 3, 0.4]]
    refc = oracle.calibration_sweep(10, 0, 501, pars + pars[:1])
    assert len(cal) == 4
    for got, want in zip(cal, refc):
        assert set(got) == set(want)
        for k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-12)
