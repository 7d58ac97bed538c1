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

    def calib(seed, runpars, count, first):
        from oracle.pyoracle import Calibration
        from microsimulation_b200.api import _seed6
        out = []
        for par in runpars:
            c = Calibration()
            assert emu.emu_calibration((C.c_double * 6)(*_seed6(seed)), C.c_int64(first), C.c_int64(count),
                                       (C.c_double * 6)(*par), C.c_int(2), C.byref(c)) == 0
            out.append(c.to_dict())
        return out

    cal = D.calibration_sweep_sharded(10, [[4, 0.5, 0.05, 10, 3, 0.5], [4, 0.0, 0.3, 8, 3, 0.2]], 501, compute=calib)
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
    refc = oracle.calibration_sweep(10, 0, 501, [[4, 0.5, 0.05, 10, 3, 0.5], [4, 0.0, 0.3, 8, 3, 0.2]])
    for got, want in zip(cal, refc):
        assert set(got) == set(want)
        for k in want:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-12)
