#!/bin/bash
# Round-2 final GPU call: everything the profiles/ summaries and DESIGN.md section 9 quote, from ONE tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02z
timeout 900 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.txt 2>&1; tail -4 ${O}_gpu_tests.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > ${O}_clocks.csv &
SMI=$!
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; cut -c1-200 ${O}_bench_1gpu.json; tail -2 ${O}_bench_1gpu.err
kill $SMI
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cut -c1-200 ${O}_bench_reference.json
python - <<'P'
import json
for f in ("gpurun_out/r02z_bench_1gpu.json","gpurun_out/r02z_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d.get(k) for k in ('value','ms_per_step','parity_checked','mismatches')}); print('  cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'}); print('  issue', (d.get('roofline') or {}).get('issue'))
    except Exception as e: print(f, 'no line', e)
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_launches.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o ${O}_prof_verify -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_ncu_verify.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -c 3 -o ${O}_prof_pairs -f python scripts/stage_bench.py cfg1 > ${O}_ncu_pairs.log 2>&1 || true
timeout 300 ncu --set full --clock-control none -k regex:k_quad_query -c 12 -o ${O}_prof_quads -f python scripts/stage_bench.py cfg1 > ${O}_ncu_quads.log 2>&1 || true
timeout 300 ncu --set full --clock-control none -k regex:k_rigid -c 9 -o ${O}_prof_rigid -f python scripts/stage_bench.py cfg1 > ${O}_ncu_rigid.log 2>&1 || true
timeout 400 python scripts/stage_bench.py cfg1 cfg3 cfg4 > ${O}_stage.jsonl 2>&1; cut -c1-220 ${O}_stage.jsonl
timeout 200 bash scripts/demo_timing.sh > ${O}_demo_timing.txt 2>&1; head -4 ${O}_demo_timing.txt
for b in 1 32; do S4PCS_BATCH=$b LANES="1" DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh; done > ${O}_batch_bench.jsonl 2>&1; cat ${O}_batch_bench.jsonl
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_verify_gpu.py tests/test_pairs_gpu.py tests/test_batch_gpu.py -x -q -m gpu -k "not full_size and not hippo and not traces" > ${O}_sanitizer_memcheck.txt 2>&1; tail -3 ${O}_sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_verify_gpu.py tests/test_pairs_gpu.py -x -q -m gpu -k "counts_match or edge or ragged or match_oracle" > ${O}_sanitizer_racecheck.txt 2>&1; tail -3 ${O}_sanitizer_racecheck.txt
ls -la gpurun_out | grep r02z_
