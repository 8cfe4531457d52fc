#!/bin/bash
# row f1: end-to-end wall time of the reference's demo main on our headers (hippo pair) for 1 / 2 / 4 / 8 lanes
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import numpy as np
h = np.load("tests/golden/hippo.npz")
for nme, arr in (("/tmp/hippo_a.obj", h["P"]), ("/tmp/hippo_b.obj", h["Q"])):
    with open(nme, "w") as f:
        for p in arr: f.write("v %.9g %.9g %.9g\n" % tuple(p))
PY
super4pcs_b200/lib/Super4PCS -i /tmp/hippo_a.obj /tmp/hippo_b.obj -o 0.7 -d 0.01 -t 1000 -n 200 -m /tmp/mat_w.txt > /dev/null 2>&1 || true   # warm the driver
for n in 200 1000 3000; do
  for lanes in 1 2 4 8; do
    t0=$(date +%s.%N)
    sc=$(S4PCS_LANES=$lanes super4pcs_b200/lib/Super4PCS -i /tmp/hippo_a.obj /tmp/hippo_b.obj -o 0.7 -d 0.01 -t 1000 -n $n -m /tmp/mat_${n}_$lanes.txt 2>&1 | tr "\r" "\n" | grep -E "^Score" | tail -1)
    t1=$(date +%s.%N)
    echo "n=$n lanes=$lanes $sc wall=$(python -c "print(round($t1-$t0,3))")s md5=$(md5sum < /tmp/mat_${n}_$lanes.txt | cut -c1-8)"
  done
done
