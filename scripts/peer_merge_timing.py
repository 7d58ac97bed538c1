"""torchrun script: what the merge of a sharded person-r step costs — inside the launch (NVLink peer memory) against a
separate NCCL all-reduce — measured on steps so small (1000 persons per rank) that nothing else is in them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import microsimulation_b200 as m
import microsimulation_b200.api as api
from microsimulation_b200 import distributed as D

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m.init(local)
api.set_stream(torch.cuda.current_stream().cuda_stream)
outputs = api.MSIM_OUT_REPORT | api.MSIM_OUT_COUNTS


def timed(n, steps, nccl):
    def step():
        api.person_enqueue(12345, rank * n, n, outputs, 0.03)
        if nccl:
            ptr, nd = api.dense_device_ptr()
            dist.all_reduce(torch.as_tensor(D._CudaBuffer(ptr, nd), device="cuda"))
    for _ in range(20):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3


for n in (1000, 12_500_000 // 1):
    steps = 300 if n < 10 ** 6 else 30
    alone = timed(n, steps, False)
    nccl = timed(n, steps, True)
    ok = D.enable_peer_merge()
    fused = timed(n, steps, False) if ok else float("nan")
    if ok:
        api.peer_merge(2)
    later = timed(n, steps, False) if ok else float("nan")
    D.disable_peer_merge()
    if rank == 0:
        print("world %d  %9d persons per rank: step without merge %.1f us, + NCCL all-reduce %.1f us, merged in the launch %.1f us, "
              "merged in the next launch %.1f us" % (world, n, alone, nccl, fused, later), flush=True)
dist.destroy_process_group()
