#!/bin/bash
# One GPU session: parity tests, smoke, bench, then ncu (launch list + full captures of the top kernels).
# Every ncu command runs only after the same command line exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
PROF="python bench.py --steps 3 --warmup 3 --persons 10000000 --no-cpu-baseline --e2e-steps 1"
$PROF > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
bash scripts/gpu_prof.sh 6 staged_report 1e7 1
bash scripts/gpu_prof.sh 1 staged_rowlog 1e7 5
ls -la gpurun_out
