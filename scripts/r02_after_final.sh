#!/bin/bash
# after the last kernel change: the archived bench lines (ours with CPU arm + parity + issue roofline, reference arm)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err; cut -c1-200 gpurun_out/r02n_bench_1gpu.json; tail -2 gpurun_out/r02n_bench_1gpu.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err; cut -c1-200 gpurun_out/r02n_bench_reference.json
python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench_1gpu.json').read().strip().splitlines()[-1]); print(d['roofline']['issue']); print(d['value'], d['e2e']['value'], d['parity_checked'], d['mismatches'], d['cpu_baseline']['value'])"
