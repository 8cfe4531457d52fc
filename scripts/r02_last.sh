#!/bin/bash
# Round-2 closing call (1 GPU): what the driver runs at round end -- the GPU tests, smoke(), both bench arms -- on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02u
timeout 600 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.txt 2>&1; tail -3 ${O}_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cut -c1-160 ${O}_bench_reference.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; tail -2 ${O}_bench_1gpu.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02u_bench_1gpu.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','parity_checked','mismatches','winner_key')}, d['e2e'])
print('issue', d['roofline'].get('issue')); print('hbm', {k:d['roofline'].get(k) for k in ('achieved','peak','frac','traffic','kernel_ms')})
print('cpu', d['cpu_baseline']['value'], d['clocks'], d.get('collective'))
P
