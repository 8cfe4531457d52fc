#!/bin/bash
# after the final call: the bench line with the issue roofline (profiles/verify_ncu.json now matches the sources) and the
# complete launch list of a bench process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err; cut -c1-200 gpurun_out/r02n_bench_1gpu.json; tail -2 gpurun_out/r02n_bench_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench_1gpu.json').read().strip().splitlines()[-1]); print(d['roofline']['issue']); print(d['roofline']['traffic'], d['roofline']['traffic_source']); print(d['value'], d['parity_checked'], d['mismatches'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_launches.log 2>&1 || true
tail -2 gpurun_out/r02z_launches.log | cut -c1-200; wc -l gpurun_out/r02z_launches.csv
