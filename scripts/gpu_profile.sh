#!/bin/bash
# ncu launch list (per-launch device time) + one full capture of the named kernels, for profiles/.
# usage: scripts/gpu_profile.sh <tag> <kernel-regex> [more regexes...]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-pipeline --no-config4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_${K} $BENCH > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out | tail -12
