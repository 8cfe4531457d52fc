#!/bin/bash
# Round-2 GPU call A: baseline of the round-1 tree + the ncu captures VERDICT asked for (pairs / quads / rigid).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02a_gpu_tests.txt
DEVICE_SPECS="1" timeout 200 scripts/lanes_bench.sh > gpurun_out/r02a_lanes_bench.jsonl 2>&1; cat gpurun_out/r02a_lanes_bench.jsonl
timeout 200 bash scripts/demo_timing.sh > gpurun_out/r02a_demo_timing.txt 2>&1; tail -40 gpurun_out/r02a_demo_timing.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench_1gpu.err; cut -c1-300 gpurun_out/r02a_bench_1gpu.json
timeout 120 python scripts/stage_bench.py cfg1 cfg3 > gpurun_out/r02a_stage.jsonl 2>&1; cut -c1-250 gpurun_out/r02a_stage.jsonl
for k in k_pairs k_quad_query k_rigid; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o gpurun_out/r02a_prof_$k -f \
    python scripts/stage_bench.py cfg1 > gpurun_out/r02a_ncu_$k.log 2>&1 || true
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -c 2 -o gpurun_out/r02a_prof_k_pairs_cfg3 -f \
    python scripts/stage_bench.py cfg3 > gpurun_out/r02a_ncu_k_pairs_cfg3.log 2>&1 || true
timeout 150 python tests/fuzz_gpu_vs_port.py 1 90 > gpurun_out/r02a_fuzz_gpu_vs_port.txt 2>&1; tail -3 gpurun_out/r02a_fuzz_gpu_vs_port.txt
ls -la gpurun_out | grep r02a_
