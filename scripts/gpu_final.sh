#!/bin/bash
# final round-1 evidence: launch list of the default bench, full ncu captures of the shared-template kernels,
# C5 stage breakdowns.  usage: scripts/gpu_final.sh <tag>
TAG=${1:-r1v}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
C4="python bench.py --workload c4_template_2048_w8_dk2_db2_fp32 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c4.csv $C4 > gpurun_out/launches_${TAG}_c4.log 2>&1
for K in chol_subst_kernel fit_seg3_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c4_${K} $C4 > gpurun_out/prof_${TAG}_c4_${K}.log 2>&1
done
timeout 600 python bench.py --workload c5half_8192_w12_dk3_db2_fp64 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 | tee gpurun_out/bench_${TAG}_c5half.json
timeout 900 python bench.py --workload c5_16384_w12_dk3_db2_fp64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 | tee gpurun_out/bench_${TAG}_c5.json
ls -la gpurun_out | tail -12
