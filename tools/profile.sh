#!/bin/bash
# ncu evidence for one round. Usage (under gpurun): bash tools/profile.sh r1 [kernel regex] [extra bench args]
# 1) per-launch device times of a short bench run (shares, not absolutes)
# 2) one --set full capture of the heavy kernels
set -u
TAG=${1:-r1}
KRX=${2:-k_copy|k_ray_update}
EXTRA=${3:-}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-strict --no-full-copy --no-eager $EXTRA"
$CMD > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
$CMD > $OUT/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KRX" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-6} -o $OUT/prof_${TAG} -f $CMD > $OUT/ncu_full_${TAG}.log 2>&1
ls -la $OUT | tail -8
