#!/bin/bash
# helper for gpurun: GPU parity tests + short bench (zero-copy and staged e2e), logs to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tee gpurun_out/bench.log
CMPC_NO_ZEROCOPY=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('staged e2e', d['e2e'])" | tee gpurun_out/bench_staged.log
