#!/bin/bash
# builds scripts/lanes_bench.cc against the product's headers/libraries and runs it on the hippo pair for three sample sizes
# (needs Eigen: S4_EIGEN_ROOT, the reference's vendored copy, or /usr/include/eigen3).  LD_PRELOAD a stand-in to dry-run on CPU.
set -e
R="$(cd "$(dirname "$0")/.." && pwd)"
EIG="${S4_EIGEN_ROOT:-/root/reference/3rdparty/Eigen}"; [ -d "$EIG/Eigen" ] || EIG=/usr/include/eigen3
W="$(mktemp -d)"
# Eigen is only in the build container: the binary is built there into super4pcs_b200/lib/ (git-ignored, travels to the GPU box)
BIN="$R/super4pcs_b200/lib/lanes_bench"
if [ -d "$EIG/Eigen" ]; then
  g++ -std=c++14 -O2 -w -I "$R/include" -I "$EIG" "$R/scripts/lanes_bench.cc" -o "$BIN" -L "$R/super4pcs_b200/lib" \
      -lsuper4pcs_b200 -ls4g -Wl,-rpath,'$ORIGIN'
fi
[ -x "$BIN" ] || { echo "lanes_bench: no Eigen here and no prebuilt $BIN" >&2; exit 3; }
[ -n "$BUILD_ONLY" ] && exit 0
python - "$R" "$W" <<'PY'
import sys
import numpy as np
h = np.load(sys.argv[1] + "/tests/golden/hippo.npz")
for nme, arr in (("a.obj", h["P"]), ("b.obj", h["Q"])):
    open(sys.argv[2] + "/" + nme, "w").write("\n".join("v %.9g %.9g %.9g" % tuple(p) for p in arr))
PY
# DEVICE_SPECS="1 0,0 2": one pass per S4PCS_DEVICES value (row e inside the C++ layer)
for dv in ${DEVICE_SPECS:-1}; do for n in ${SIZES:-200 1000 3000}; do S4PCS_DEVICES="$dv" "$BIN" "$W/a.obj" "$W/b.obj" 0.7 0.01 "$n" "${REPS:-5}" "${LANES:-1 2 4 8}"; done; done
