#!/bin/bash
# end-to-end wall time of the reference's demo main compiled against our headers (hippo pair written from
# tests/golden/hippo.npz), for sample sizes 200 / 1000 / 3000  (reference 0.22-0.31 s / 42.9 s / 207 s)
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import numpy as np
h = np.load("tests/golden/hippo.npz")
for nme, arr in (("/tmp/hippo_a.obj", h["P"]), ("/tmp/hippo_b.obj", h["Q"])):
    with open(nme, "w") as f:
        for p in arr: f.write("v %.9g %.9g %.9g\n" % tuple(p))
PY
for n in 200 1000 3000; do
  t0=$(date +%s.%N)
  sc=$(super4pcs_b200/lib/Super4PCS -i /tmp/hippo_a.obj /tmp/hippo_b.obj -o 0.7 -d 0.01 -t 1000 -n $n -m /tmp/mat_$n.txt 2>&1 | tr "\r" "\n" | grep -E "^Score" | tail -1)
  t1=$(date +%s.%N)
  echo "n=$n $sc wall=$(python -c "print(round($t1-$t0,3))")s"
done
# where the time of a run goes: the S4PCS_TIMINGS report (device ms per stage from the libs4g events vs the host wall clock
# of ComputeTransformation), sequential loop and 4 lanes
for n in 200 1000 3000; do for lanes in 1 4; do
  echo "--- n=$n lanes=$lanes"
  S4PCS_TIMINGS=1 S4PCS_LANES=$lanes super4pcs_b200/lib/Super4PCS -i /tmp/hippo_a.obj /tmp/hippo_b.obj -o 0.7 -d 0.01 -t 1000 -n $n -m /tmp/mat_t.txt 2>&1 | tr "\r" "\n" | grep -A9 "Timings (msec)"
done; done
