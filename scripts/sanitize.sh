#!/bin/bash
# compute-sanitizer passes over small instances of every kernel (memcheck + racecheck); summary lines only
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  echo "== $tool: verify (20000 x 20000, 32 candidates)"
  compute-sanitizer --tool $tool --print-limit 3 python scripts/dbg_verify.py 20000 0.01 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" | head -5
  echo "== $tool: pairs -> quads -> TryCongruentSet chain (golden synthetic stages)"
  compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_quads_gpu.py::test_chain_matches_golden_synthetic -q -m gpu 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|passed|failed" | head -6
done
