#!/bin/bash
# Round-2 GPU session A: parity tests, staged-kernel timings, one ncu capture of the headline kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/staged_timing.py > gpurun_out/staged_timing.log 2>&1; echo "timing rc=$?"; cat gpurun_out/staged_timing.log
timeout 600 bash scripts/gpu_prof.sh 6 staged_report 1e7 1
ls -la gpurun_out | tail -5
