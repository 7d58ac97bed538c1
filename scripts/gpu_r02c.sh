#!/bin/bash
# Round-2 GPU session C: parity tests, timings, bench (N = 1), ncu: launch list, full captures of the person-r kernels at
# the bench's size (1e8 persons), the instruction-floor microbenchmarks.  Every ncu command runs only after the same
# command line exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
python -c "from microsimulation_b200 import build; print(build.source_hash())" > gpurun_out/build_hash.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/staged_timing.py 1e8 > gpurun_out/staged_timing.log 2>&1; echo "timing rc=$?"; cat gpurun_out/staged_timing.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; echo
PROF="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-calib"
$PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 900 bash scripts/gpu_prof.sh 6 staged_report 1e8 1
timeout 900 bash scripts/gpu_prof.sh 1 staged_rowlog 1e8 5
python scripts/prof_calib.py 256 1e6 > gpurun_out/plain_calib.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:calib_shared -s 1 -c 1 -f -o gpurun_out/prof_calib_shared python scripts/prof_calib.py 256 1e6 > gpurun_out/ncu_calib.log 2>&1
echo "ncu calib rc=$?"
python scripts/calib_timing.py > gpurun_out/calib_timing.log 2>&1; cat gpurun_out/calib_timing.log
python scripts/inst_floor.py run > gpurun_out/inst_floor_plain.log 2>&1 && \
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:microbench --csv \
    --log-file gpurun_out/inst_floor.csv python scripts/inst_floor.py run > gpurun_out/inst_floor_ncu.log 2>&1
echo "inst floor rc=$?"; cat gpurun_out/inst_floor_plain.log
ls -la gpurun_out | tail -12
