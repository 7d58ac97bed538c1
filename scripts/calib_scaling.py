"""BASELINE configs[4]: the calibration sweep (256 parameter sets x 1e6 persons) sharded by person range over the
ranks, one all-reduce to merge (strong scaling: the total work is fixed).  Prints one JSON line on rank 0.
   python scripts/calib_scaling.py                       # one GPU
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/calib_scaling.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import microsimulation_b200 as m
import microsimulation_b200.api as api
from microsimulation_b200.distributed import calibration_sweep_sharded

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m.init(local)
api.set_stream(torch.cuda.current_stream().cuda_stream)
sets, n, steps = 256, 10**6, 5
base = np.array([4, 0.5, 0.05, 10, 3, 0.5])
lattice = np.random.RandomState(20251020).randint(-2, 3, size=(sets, 6)) * 0.1  # each component -20%..+20% on a fixed lattice
pars = base * (1 + lattice)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    res = calibration_sweep_sharded(12345, pars, n)
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    res = calibration_sweep_sharded(12345, pars, n)
barrier()
dt = torch.tensor([(time.perf_counter() - t0) / steps], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    occ = sum(float(r[k].sum()) for r in res for k in r if k != "TimeAtRisk")
    print(json.dumps({"workload": "calibration sweep, BASELINE configs[4]: 256 parameter sets x 1e6 persons, person ranges sharded over the ranks, merged by all-reduce",
                      "n_gpus": world, "ms_per_sweep": dt.item() * 1e3, "person_simulations_per_s": sets * n / dt.item(),
                      "kernel_ms_rank0": api.last_kernel_ms(), "scaling": "strong", "count_events_recorded": occ}))
if world > 1:
    dist.destroy_process_group()
