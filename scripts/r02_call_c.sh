#!/bin/bash
# Round-2 GPU call C: k_pairs v2 (symmetric, queue-compacted) + k_verify v12 (trimmed phase 1, point-parallel phase 2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_gpu_tests.txt 2>&1; tail -15 gpurun_out/r02c_gpu_tests.txt
timeout 200 python scripts/stage_bench.py cfg1 cfg3 > gpurun_out/r02c_stage.jsonl 2>&1; cut -c1-330 gpurun_out/r02c_stage.jsonl
VARIANTS='|-DS4G_FLAT_PHASE2=0|-DS4G_VERIFY_MIN_BLOCKS=10|-DS4G_ITEM_CAP=512' OUT=gpurun_out/r02c_verify_ab.jsonl TESTS="tests/test_verify_gpu.py" bash scripts/verify_ab.sh 2>&1 | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02c_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_verify.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -c 3 -o gpurun_out/r02c_prof_k_pairs -f \
    python scripts/stage_bench.py cfg1 > gpurun_out/r02c_ncu_k_pairs.log 2>&1 || true
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_pairs_gpu.py tests/test_verify_gpu.py -x -q -m gpu -k "not full_size" > gpurun_out/r02c_sanitizer.txt 2>&1; tail -5 gpurun_out/r02c_sanitizer.txt
ls -la gpurun_out | grep r02c_
