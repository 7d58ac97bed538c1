#!/bin/bash
# Round-2 GPU session B: parity tests, the bench line (N = 1), staged timings.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/staged_timing.py 1e8 > gpurun_out/staged_timing.log 2>&1; echo "timing rc=$?"; cat gpurun_out/staged_timing.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
