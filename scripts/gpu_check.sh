#!/bin/bash
# One gpurun call: smoke, GPU parity tests (grouped so that a faulting kernel does not hide later groups), bench.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi -L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== pytest fft";    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fft_engine" 2>&1 | tail -30
echo "=== pytest stages"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stages or fast_paths" 2>&1 | tail -60
echo "=== pytest rest";   timeout 1200 python -m pytest tests -m gpu -q -k "not fft_engine and not stages and not fast_paths" 2>&1 | tail -80
if [ "$1" != "nobench" ]; then
echo "=== bench dev";     timeout 600 python bench.py --workload dev_1024_w4_dk2_db2_fp32 --steps 10 --warmup 3 --no-cpu-baseline
echo "=== bench c2";      timeout 900 python bench.py --steps 20 --warmup 3 --cpu-sample 512
fi
