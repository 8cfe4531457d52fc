#!/bin/bash
# Round-2 GPU call B: k_verify v11 (delta-field) -- parity first, then bench, then one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_verify_gpu.py tests/test_errors_gpu.py -x -q -m gpu > gpurun_out/r02b_verify_tests.txt 2>&1; tail -15 gpurun_out/r02b_verify_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; cut -c1-300 gpurun_out/r02b_bench_1gpu.json; tail -3 gpurun_out/r02b_bench_1gpu.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02b_bench_1gpu.json').read().strip().splitlines()[-1]); print({k:d['roofline'][k] for k in d['roofline'] if k!='peak_source'})
except Exception as e: print('no bench line', e)
P
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gpu_tests.txt 2>&1; tail -5 gpurun_out/r02b_gpu_tests.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02b_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu_verify.log 2>&1 || true
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "not full and not 1m and not million" > gpurun_out/r02b_sanitizer.txt 2>&1; tail -5 gpurun_out/r02b_sanitizer.txt
ls -la gpurun_out | grep r02b_
