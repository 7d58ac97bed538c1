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
