#!/bin/bash
# The first GPU call of round 2, prepared at the end of round 1 (when the GPU budget was spent): everything that could
# not be measured any more, bounded by timeouts, outputs under gpurun_out/r02_*.
#   gpurun --timeout 2400 -- 'bash scripts/round2_first_gpu_call.sh'      (every step has its own timeout; typical total ~12 min)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. the GPU tests added after the last round-1 GPU session (plus everything else, they are fast)
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests.txt
# 2. row f1 + row e in the C++ layer: lane counts x device-context counts (S4PCS_DEVICES=0,0 = two contexts of one GPU), in-process timing (median of 5, warm-up excluded), hippo at three sample sizes
DEVICE_SPECS="1 0,0" timeout 300 scripts/lanes_bench.sh > gpurun_out/r02_lanes_bench.jsonl 2>&1; cat gpurun_out/r02_lanes_bench.jsonl
# 2b. S4PCS_DEVICES on the cfg2-sized pair: TryCongruentSet through the C++ layer, 4096 quads (repeat with gpurun --gpus 8 and --devices "1 2 4 8")
timeout 240 python scripts/devices_bench.py --points 1000000 --devices "1 0,0" > gpurun_out/r02_devices_bench.jsonl 2>&1; cat gpurun_out/r02_devices_bench.jsonl
# 2c. per-stage device time vs host wall clock of the demo (S4PCS_TIMINGS), n = 200 / 1000 / 3000, 1 and 4 lanes
timeout 240 bash scripts/demo_timing.sh > gpurun_out/r02_demo_timing.txt 2>&1; tail -60 gpurun_out/r02_demo_timing.txt
# 3. headline line (re-check against round 1: 483 K candidates/s, 8.48 ms/step)
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; cut -c1-300 gpurun_out/r02_bench_1gpu.json
# 4. ncu --set full of the two stage kernels that were only timed so far (pairs, quad query), small launch counts
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -c 2 -o gpurun_out/r02_prof_pairs -f \
    python scripts/stage_bench.py cfg1 > gpurun_out/r02_ncu_pairs.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_quad_query -c 2 -o gpurun_out/r02_prof_quads -f \
    python scripts/stage_bench.py cfg1 > gpurun_out/r02_ncu_quads.log 2>&1 || true
# 5. stage-level fuzz of the device against the oracle port (2 minutes)
timeout 200 python tests/fuzz_gpu_vs_port.py 1 120 > gpurun_out/r02_fuzz_gpu_vs_port.txt 2>&1; tail -5 gpurun_out/r02_fuzz_gpu_vs_port.txt
ls -la gpurun_out | grep r02_
# next call: gpurun --timeout 1500 -- 'bash scripts/verify_ab.sh'   (compile-time variants of k_verify: parity + bench each)
