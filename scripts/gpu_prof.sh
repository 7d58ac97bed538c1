#!/bin/bash
# ncu --set full of the staged person-r kernel in one output mode: scripts/gpu_prof.sh OUTPUTS NAME [PERSONS]
set -u
O=$1; NAME=$2; N=${3:-1e7}
CMD="python scripts/prof_person.py $O $N"
$CMD > gpurun_out/plain_$NAME.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'person_staged|cohort_kernel|rowpass' -s 3 -c ${4:-1} -f -o gpurun_out/prof_$NAME $CMD > gpurun_out/ncu_$NAME.log 2>&1
echo "ncu $NAME rc=$?"
