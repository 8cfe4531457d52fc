#!/bin/bash
# k_verify with the merged second-level look-ups; k_pairs with the cp.async.bulk stage as the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02p_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02p_gpu_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null > gpurun_out/r02p_bench.json; python -c "import json; d=json.loads(open('gpurun_out/r02p_bench.json').read()); print({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms']})"
timeout 200 python scripts/stage_bench.py cfg1 cfg3 2>&1 | grep -E "ExtractPairs|cfg3" | cut -c1-260
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02p_prof_verify -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_ncu_verify.log 2>&1 || true
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_pairs_gpu.py tests/test_batch_gpu.py tests/test_verify_gpu.py -x -q -m gpu -k "not full_size and not hippo and not traces" > gpurun_out/r02p_sanitizer.txt 2>&1; tail -3 gpurun_out/r02p_sanitizer.txt
