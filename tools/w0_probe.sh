mkdir -p gpurun_out; : > gpurun_out/w0probe.txt
for B in 256 512 1024 2048; do for W in 1 2 4 8; do
CMPC_W0=$W timeout 300 python bench.py --batch $B --steps 100 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$B W0=$W ms', round(d['ms_per_step'],4), 'p50', round(d['p50_batch_latency_ms'],4), 'Msolves', round(d['value']/1e6,2), 'e2e_ms', round(d['e2e']['ms_per_step'],4), d['e2e'].get('route'))" >> gpurun_out/w0probe.txt
done; done
cat gpurun_out/w0probe.txt
