#!/bin/bash
# Round-2 GPU call H: several bases per launch chain (s4g_try_bases / S4PCS_BATCH): parity, then timing against lanes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_gpu.py -x -q -m gpu > gpurun_out/r02h_batch_tests.txt 2>&1; tail -15 gpurun_out/r02h_batch_tests.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_batch_gpu.py > gpurun_out/r02h_gpu_tests.txt 2>&1; tail -5 gpurun_out/r02h_gpu_tests.txt
for b in 1 4 8 16 32; do S4PCS_BATCH=$b LANES="1" DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh; done > gpurun_out/r02h_batch_bench.jsonl 2>&1; cat gpurun_out/r02h_batch_bench.jsonl
S4PCS_BATCH=1 LANES="4 8" DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh > gpurun_out/r02h_lanes_bench.jsonl 2>&1; cat gpurun_out/r02h_lanes_bench.jsonl
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_batch_gpu.py -x -q -m gpu -k "equals_the_per_base" > gpurun_out/r02h_sanitizer.txt 2>&1; tail -4 gpurun_out/r02h_sanitizer.txt
ls -la gpurun_out | grep r02h_
