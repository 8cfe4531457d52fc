#!/bin/bash
# Round-2 GPU call D: warp-granular cull; cfg3/cfg4-size parity tests; full bench line incl. CPU arm + parity check; ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_gpu_tests.txt 2>&1; tail -15 gpurun_out/r02d_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_1gpu.json 2> gpurun_out/r02d_bench_1gpu.err; cut -c1-250 gpurun_out/r02d_bench_1gpu.json; tail -3 gpurun_out/r02d_bench_1gpu.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench_1gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','mismatches','winner_key')})
    print(d['cpu_baseline'])
    print({k:d['roofline'][k] for k in ('frac','kernel_ms','tile_candidate_pairs_culled_frac','field_words_per_query','points_tested_per_query')})
except Exception as e: print('no bench line', e)
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -o gpurun_out/r02d_prof_verify -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_ncu_verify.log 2>&1 || true
timeout 300 python scripts/stage_bench.py cfg1 cfg3 cfg4 > gpurun_out/r02d_stage.jsonl 2>&1; cut -c1-300 gpurun_out/r02d_stage.jsonl
ls -la gpurun_out | grep r02d_
# A/B: cp.async.bulk (1-D TMA) + mbarrier double-buffered partner tiles in k_pairs
S4G_NVCC_DEFINES="-DS4G_PAIRS_TMA=1" timeout 300 python -c "from super4pcs_b200 import build; build.build_lib()" > gpurun_out/r02d_tma_build.log 2>&1
S4G_NVCC_DEFINES="-DS4G_PAIRS_TMA=1" timeout 300 python -m pytest tests/test_pairs_gpu.py tests/test_quads_gpu.py -x -q -m gpu > gpurun_out/r02d_tma_tests.txt 2>&1; tail -3 gpurun_out/r02d_tma_tests.txt
S4G_NVCC_DEFINES="-DS4G_PAIRS_TMA=1" timeout 200 python scripts/stage_bench.py cfg1 cfg3 > gpurun_out/r02d_stage_tma.jsonl 2>&1; grep ExtractPairs gpurun_out/r02d_stage_tma.jsonl | cut -c1-300
cuobjdump -sass super4pcs_b200/lib/libs4g.so | grep -c "UBLKCP" > gpurun_out/r02d_tma_sass_count.txt; cat gpurun_out/r02d_tma_sass_count.txt
timeout 300 python -c "from super4pcs_b200 import build; build.build_lib()" > /dev/null 2>&1
ls -la gpurun_out | grep r02d_
