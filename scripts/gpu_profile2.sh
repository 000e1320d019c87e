#!/bin/bash
# full ncu capture of named kernels (one launch each, after warm-up) on the default bench workload
# usage: scripts/gpu_profile2.sh <tag> <skip> <kernel-regex> [more regexes...]
TAG=${1:-r1}; SKIP=${2:-3}; shift; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-pipeline"
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG}_${K} $BENCH > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out
