#!/bin/bash
# Round-2 GPU call G: per-warp queues (no CTA barriers in the tile loop); small-queue build as a correctness variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_verify_gpu.py tests/test_errors_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/r02g_verify_tests.txt 2>&1; tail -5 gpurun_out/r02g_verify_tests.txt
VARIANTS='|-DS4G_QUEUE_CAP=128 -DS4G_FLUSH_MIN=64|-DS4G_VERIFY_MIN_BLOCKS=10|-DS4G_FLUSH_MIN=512' OUT=gpurun_out/r02g_verify_ab.jsonl TESTS="tests/test_verify_gpu.py" bash scripts/verify_ab.sh 2>&1 | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02g_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_ncu_verify.log 2>&1 || true
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "not full_size" > gpurun_out/r02g_sanitizer.txt 2>&1; tail -4 gpurun_out/r02g_sanitizer.txt
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "counts_match or edge or ragged" > gpurun_out/r02g_racecheck.txt 2>&1; tail -4 gpurun_out/r02g_racecheck.txt
ls -la gpurun_out | grep r02g_
