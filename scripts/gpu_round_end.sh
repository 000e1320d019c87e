#!/bin/bash
# What the driver runs at round end (GPU tests, smoke, both bench arms) plus the launch list and one full ncu capture.
TAG=${1:-r1end}
mkdir -p gpurun_out
exec > >(tee gpurun_out/round_end_${TAG}.log) 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== pytest -m gpu"; (time timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5)
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | tee gpurun_out/bench_${TAG}_reference.json | cut -c1-400
echo "=== bench"; timeout 900 python bench.py | tee gpurun_out/bench_${TAG}.json | cut -c1-3000
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-pipeline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
C5="python bench.py --workload c5half_8192_w12_dk3_db2_fp64 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-pipeline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_seg3_kernel -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_c5half_fit_seg3_pass $C5 > gpurun_out/prof_${TAG}_c5half_fit_seg3_pass.log 2>&1
ls -la gpurun_out | tail -8
