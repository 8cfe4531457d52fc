#!/bin/bash
# Round-2 GPU call E: two-level tile cull + sub-voxel refinement of the delta-field; A/B of the refinement; lanes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_verify_gpu.py tests/test_errors_gpu.py tests/test_zz_reference_tests_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/r02e_verify_tests.txt 2>&1; tail -8 gpurun_out/r02e_verify_tests.txt
VARIANTS='|-DS4G_SUBVOXEL=0' OUT=gpurun_out/r02e_verify_ab.jsonl TESTS="tests/test_verify_gpu.py" bash scripts/verify_ab.sh 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02e_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_ncu_verify.log 2>&1 || true
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "not full_size" > gpurun_out/r02e_sanitizer.txt 2>&1; tail -4 gpurun_out/r02e_sanitizer.txt
DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh > gpurun_out/r02e_lanes_bench.jsonl 2>&1; cat gpurun_out/r02e_lanes_bench.jsonl
timeout 200 bash scripts/demo_timing.sh > gpurun_out/r02e_demo_timing.txt 2>&1; head -4 gpurun_out/r02e_demo_timing.txt; grep -A9 "n=3000 lanes=1" gpurun_out/r02e_demo_timing.txt
timeout 120 python scripts/stage_bench.py cfg1 > gpurun_out/r02e_stage.jsonl 2>&1; grep '"Verify"' gpurun_out/r02e_stage.jsonl | cut -c1-300
ls -la gpurun_out | grep r02e_
