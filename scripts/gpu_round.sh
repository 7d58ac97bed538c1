#!/bin/bash
# One GPU session: parity tests, smoke, bench (both workloads), then ncu (launch list + full capture of the top kernel).
# Every ncu command runs only after the same command line exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
python bench.py --steps 10 --warmup 3 --workload rowlog > gpurun_out/bench_rowlog.json 2>> gpurun_out/bench.err; echo "bench rowlog rc=$?"; cat gpurun_out/bench_rowlog.json
PROF="python bench.py --steps 3 --warmup 3 --persons 10000000 --no-cpu-baseline --e2e-steps 1"
$PROF > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
$PROF > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'cohort_kernel|person_staged' -s 4 -c 1 -o gpurun_out/prof_report $PROF > gpurun_out/ncu_full.log 2>&1
echo "ncu full (report) rc=$?"
PROF2="$PROF --workload rowlog"
$PROF2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'cohort_kernel|person_staged|rowlog' -s 16 -c 4 -o gpurun_out/prof_rowlog $PROF2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full (rowlog) rc=$?"
ls -la gpurun_out
