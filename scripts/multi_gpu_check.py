"""torchrun script: sharded person-r over all ranks (NCCL) equals the single-GPU run on rank 0's device.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import microsimulation_b200 as m
import microsimulation_b200.api as api
from microsimulation_b200.distributed import person_simulation_sharded, calibration_sweep_sharded

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m.init(local)
api.set_stream(torch.cuda.current_stream().cuda_stream)
n = 3_000_001
rep, counts, rows = person_simulation_sharded(n, discount_rate=0.03)
# the same with the merge inside the launch (NVLink peer memory) instead of the NCCL all-reduce: same tables, the same
# bits on every rank, many launches in a row (mailbox slots alternate), ragged shards
from microsimulation_b200 import distributed as D
if D.enable_peer_merge():
    for n2 in (n, 1000 * world + 3, n, 5 * world, n):
        rep2, counts2, _ = person_simulation_sharded(n2, discount_rate=0.03)
    assert np.array_equal(counts2[:8], counts[:8]), (counts2, counts)
    for tb in ("events", "prev"):
        for k in rep[tb]:
            assert np.array_equal(rep2[tb][k], rep[tb][k]), (tb, k)
    for tb, k in (("pt", "pt"), ("ut", "utility")):
        np.testing.assert_allclose(rep2[tb][k], rep[tb][k], rtol=1e-12)
    ptr, nd = api.dense_device_ptr()
    mine = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda").clone()
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    assert all(torch.equal(e.view(torch.int64), mine.view(torch.int64)) for e in every)
    for _ in range(200):  # back to back, asynchronously: a rank may run one launch ahead of its peers
        api.person_enqueue(12345, rank * 1000, 1000, api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS, 0.03)
    api.sync()
    ptr, nd = api.dense_device_ptr()
    v = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda").clone()
    assert float(v[-9] + v[-8]) == 1000.0 * world, float(v[-9] + v[-8])
    # mode 2: run k adds up what the ranks sent in run k - 1; a flush adds up the last run's
    api.peer_merge(2)
    sizes = [5000 + 17 * world, 300 * world, 40_001, 2 * world]
    want = []
    for n2 in sizes:  # what every run's merged vector must be, from the NCCL path
        api.peer_merge(0)
        first, count = D.shard_range(n2, rank, world)
        api.person_run_device(12345, first, count, api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS, 0.03)
        ptr, nd = api.dense_device_ptr()
        t = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda").clone()
        dist.all_reduce(t)
        want.append(t)
    api.peer_merge(2)
    for k, n2 in enumerate(sizes):
        first, count = D.shard_range(n2, rank, world)
        api.person_enqueue(12345, first, count, api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS, 0.03)
        api.sync()
        if k > 0:  # the previous run's sum is in place now
            ptr, nd = api.peer_result_ptr()
            got = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda")
            assert torch.equal(got[-9:-1], want[k - 1][-9:-1]), (k, got[-9:], want[k - 1][-9:])
            assert torch.allclose(got, want[k - 1], rtol=1e-12, atol=0)
    api.peer_flush()
    api.sync()
    ptr, nd = api.peer_result_ptr()
    got = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda").clone()
    assert torch.equal(got[-9:-1], want[-1][-9:-1]) and torch.allclose(got, want[-1], rtol=1e-12, atol=0)
    every = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(every, got)
    assert all(torch.equal(e.view(torch.int64), got.view(torch.int64)) for e in every)
    for _ in range(300):  # back to back: ranks drift apart by up to a launch, slots must not be overwritten early
        api.person_enqueue(12345, rank * 1000, 1000 + 50 * rank, api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS, 0.03)
    api.peer_flush()
    api.sync()
    ptr, nd = api.peer_result_ptr()
    v = torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda").clone()
    assert float(v[-9] + v[-8]) == 1000.0 * world + 50.0 * sum(range(world)), float(v[-9] + v[-8])
    D.disable_peer_merge()
    if rank == 0:
        print("in-launch merge ok (both modes): world %d" % world)
elif rank == 0:
    print("in-launch merge NOT available on this box (peer memory could not be mapped)")
pars = np.tile(np.array([4, 0.5, 0.05, 10, 3, 0.5]), (8, 1)) * (1 + 0.2 * (np.random.RandomState(1).rand(8, 6) - 0.5))
cal = calibration_sweep_sharded(12345, pars, 200_001)
if rank == 0:
    one, c1 = m.person_report(12345, 0, n, 0.03)
    assert np.array_equal(counts[:8], c1[:8]), (counts, c1)
    for tb in ("events", "prev"):
        for k in rep[tb]:
            assert np.array_equal(rep[tb][k], one[tb][k]), (tb, k)
    for tb, k in (("pt", "pt"), ("ut", "utility")):
        np.testing.assert_allclose(rep[tb][k], one[tb][k], rtol=1e-11)
    cal1 = m.calibration_sweep(12345, pars, 200_001)
    for a, b in zip(cal, cal1):
        for k in b:
            if k == "TimeAtRisk":
                np.testing.assert_allclose(a[k], b[k], rtol=1e-11)
            else:
                assert np.array_equal(a[k], b[k]), k
    print("multi-GPU check ok: world %d, n %d, events %d" % (world, n, int(counts[:8].sum())))
dist.destroy_process_group()
