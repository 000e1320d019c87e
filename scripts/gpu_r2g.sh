#!/bin/bash
# round-2 (second half) evidence: config 3 after the local-support skip / three-slot ring / compacted FIR / 512 R row passes.
# usage: scripts/gpu_r2g.sh <tag>
TAG=${1:-r2g}
mkdir -p gpurun_out
C3="python bench.py --workload c3_6144_bspline_fp32 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c3.csv $C3 > gpurun_out/launches_${TAG}_c3.log 2>&1
for K in fit_gen4_kernel gen_fir_kernel row_fwd_g16_kernel row_inv_g16_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o gpurun_out/prof_${TAG}_${K} $C3 > gpurun_out/prof_${TAG}_${K}.log 2>&1
done
timeout 600 python bench.py --workload c3_6144_bspline_fp32 --steps 5 --warmup 3 | tail -1 > gpurun_out/bench_${TAG}_c3.json
timeout 900 python bench.py --workload c5_16384_w12_dk3_db2_fp64 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 | tail -1 > gpurun_out/bench_${TAG}_c5.json
timeout 200 python scripts/dbg/sizes.py 4096x4096 3080x3072 3072x3072 6144x6144 2560x2560 1536x1536 5120x5120 > gpurun_out/sizes_${TAG}.txt 2>&1
ls -la gpurun_out | tail -14
