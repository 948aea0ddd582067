#!/bin/bash
# smoke + default bench + 2-GPU torchrun bench (+ reference arm)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench_ref.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.log | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n2.log | cut -c1-1200
