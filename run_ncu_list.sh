#!/bin/bash
# ncu launch list of the bench command (per-launch device time; cold-cache, serialised) + one full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -n 2 gpurun_out/plain_bench.log | cut -c1-300
./run_ncu.sh
