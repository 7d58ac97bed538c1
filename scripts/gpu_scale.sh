#!/bin/bash
# bench.py on N GPUs of one box, as the driver launches it: scripts/gpu_scale.sh N   (N = 2, 4, 8; run under gpurun --gpus N)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("weak ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"])
print("strong", json.dumps(d.get("strong")))
print("merge_check", json.dumps(d.get("merge_check")))
c = d.get("calib") or {}
print("calib", {k: c.get(k) for k in ("ms_per_step", "kernel_ms_rank0", "event_loop_kernel_ms_rank0", "split", "value")})
PY
