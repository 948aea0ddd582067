#!/bin/bash
# scaling bench on one box: ./run_multi.sh 1 2 4 8  -> gpurun_out/bench_n<N>.log (one JSON line each), driver-style launch
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1.log | cut -c1-600
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n$n.log | cut -c1-600
  fi
done
