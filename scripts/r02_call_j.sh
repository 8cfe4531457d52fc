#!/bin/bash
# A/B of the coarse-occupancy resolution of the tile cull (S4G_CSHIFT_MIN: 1 = 2x2x2 cells (default), 2 = 4x4x4 cells)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cs in 1 2 3; do
  echo "== S4G_CSHIFT_MIN=$cs"
  S4G_CSHIFT_MIN=$cs timeout 300 python -m pytest tests/test_verify_gpu.py -x -q -m gpu -k "counts_match or dense_cloud" 2>&1 | tail -1
  S4G_CSHIFT_MIN=$cs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'culled': d['roofline']['tile_candidate_pairs_culled_frac']})"
done
