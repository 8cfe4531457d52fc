#!/bin/bash
# Round-2 GPU call I: sparse occupancy nibbles (vocc) -- parity, bench, whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_verify_gpu.py tests/test_errors_gpu.py tests/test_batch_gpu.py -x -q -m gpu > gpurun_out/r02i_verify_tests.txt 2>&1; tail -5 gpurun_out/r02i_verify_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_1gpu.json 2> gpurun_out/r02i_bench_1gpu.err; cut -c1-230 gpurun_out/r02i_bench_1gpu.json; tail -2 gpurun_out/r02i_bench_1gpu.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02i_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_ncu_verify.log 2>&1 || true
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_verify_gpu.py --deselect tests/test_batch_gpu.py > gpurun_out/r02i_gpu_tests.txt 2>&1; tail -4 gpurun_out/r02i_gpu_tests.txt
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "not full_size" > gpurun_out/r02i_sanitizer.txt 2>&1; tail -3 gpurun_out/r02i_sanitizer.txt
S4PCS_BATCH=32 LANES="1" DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh > gpurun_out/r02i_batch_bench.jsonl 2>&1; cat gpurun_out/r02i_batch_bench.jsonl
ls -la gpurun_out | grep r02i_
