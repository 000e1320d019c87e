#!/bin/bash
# quick loop: stage/fast-path parity tests, bench c2, optional ncu of named kernels.  usage: gpu_quick.sh <tag> [kernel regex...]
TAG=${1:-q}; shift
mkdir -p gpurun_out
echo "=== pytest stages"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or fast_paths or golden_ztf" 2>&1 | tail -15
echo "=== bench c2";      timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_${TAG}.json
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_${K} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
