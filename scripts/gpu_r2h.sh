#!/bin/bash
# end-of-round evidence (round 2, second half): full GPU suite, default bench line, reference arm, launch list, stage times by size.
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_${TAG}.txt
cat gpurun_out/gputests_${TAG}.txt
timeout 400 python bench.py | tail -1 > gpurun_out/bench_${TAG}.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 > gpurun_out/bench_${TAG}_reference.json
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 python bench.py --workload c4_template_2048_w8_dk2_db2_fp32 --steps 20 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_${TAG}_c4.json
C4="python bench.py --workload c4_template_2048_w8_dk2_db2_fp32 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-pipeline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_jonly_cached -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_fit_jonly_cached_kernel $C4 > gpurun_out/prof_${TAG}_jc.log 2>&1
timeout 200 python scripts/dbg/sizes.py 4096x4096 4088x4088 2046x4094 4000x4072 3080x3072 3072x3072 6144x6144 5120x5120 2560x2560 > gpurun_out/sizes_${TAG}.txt 2>&1
python -c "import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config4']['value'], d['roofline']['frac'])"
ls -la gpurun_out | tail -8
