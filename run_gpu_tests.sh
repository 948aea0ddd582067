#!/bin/bash
# helper for gpurun: GPU parity tests + short bench (default pipelined e2e, then the zero-copy, staged and progressive e2e modes), logs to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tee gpurun_out/bench.log
for mode in 1 2 3; do
CMPC_E2E_MODE=$mode timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e mode $mode (1 = zero-copy, 2 = staged, 3 = progressive)', d['e2e'])" | tee -a gpurun_out/bench_e2e_modes.log
done
