#!/bin/bash
mkdir -p gpurun_out
for w in 1 2; do
  echo "== CMPC_W0=$w" | tee -a gpurun_out/ab.log
  CMPC_W0=$w timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','mean_ipm_iters','max_kkt')}, d['e2e']['value'], d['roofline']['frac'])" | tee -a gpurun_out/ab.log
done
CMPC_W0=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/ab.log
