"""TEST INFRASTRUCTURE, not collected by pytest: time-boxed fuzz of the WHOLE pipeline -- the product's C++ layer (host logic)
on the CPU stand-in for libs4g against the compiled reference -- over random clouds, options and lane counts.
  LD_PRELOAD=tests/_build/libs4g_oracle_shim.so python tests/fuzz_pipeline_vs_reference.py [seed] [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _build, ref as oref  # noqa: E402
from super4pcs_b200 import synth  # noqa: E402

assert "libs4g_oracle_shim" in os.environ.get("LD_PRELOAD", ""), "run with LD_PRELOAD=tests/_build/libs4g_oracle_shim.so"
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120
t0, n_cfg, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    n = int(rng.randint(120, 600))
    ov = float(rng.choice([0.3, 0.5, 0.7, 0.9]))
    normals = bool(rng.randint(0, 2))
    d = synth.make_pair(n, ov, seed=int(rng.randint(1, 10 ** 6)), with_normals=normals, noise_sigma=float(rng.choice([0.0, 0.002])),
                        outlier_frac=float(rng.choice([0.0, 0.1])))
    kw = dict(delta=float(rng.choice([0.01, 0.02, 0.04])), overlap=ov, sample_size=int(rng.choice([40, 80, 150])),
              max_time_seconds=10000, random_seed=int(rng.randint(0, 2 ** 31 - 1)),
              terminate_threshold=float(rng.choice([1.0, 1.0, max(ov, 0.8)])))
    if normals and rng.randint(0, 2):
        kw["max_normal_difference"] = float(rng.choice([20.0, 45.0]))
    if rng.randint(0, 3) == 0:
        kw["max_translation_distance"] = 3.0
    opt = oref.make_options(**kw)
    lanes = int(rng.choice([1, 2, 3, 5]))
    # S4G_SHIM_REFERENCE_ORDER=1 (pairs in the reference's emission order): one lane, because every lane context has its own
    # id-array history; then NO difference is expected at all, without it only equal-count ties may differ
    os.environ["S4PCS_LANES"] = "1" if os.environ.get("S4G_SHIM_REFERENCE_ORDER") else str(lanes)
    os.environ["S4PCS_FUSED"] = str(int(rng.choice([1, 1, 0])))
    a = oref.compute_transformation(d["P"], d["Q"], opt, Pn=d["Pn"], Qn=d["Qn"])
    b = oref.compute_transformation(d["P"], d["Q"], opt, Pn=d["Pn"], Qn=d["Qn"], libpath=_build.DROPIN_SO)
    same = np.float32(a[0]) == np.float32(b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and \
        np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    if not same:
        bad += 1
        print("DIFF", n, kw, os.environ["S4PCS_LANES"], os.environ["S4PCS_FUSED"], a[0], b[0], flush=True)
        np.savez("/tmp/fuzz_fail_%d.npz" % bad, P=d["P"], Q=d["Q"], Pn=d["Pn"] if d["Pn"] is not None else np.zeros((0, 3), np.float32),
                 Qn=d["Qn"] if d["Qn"] is not None else np.zeros((0, 3), np.float32), kw=repr(kw),
                 lanes=os.environ["S4PCS_LANES"], fused=os.environ["S4PCS_FUSED"])
    n_cfg += 1
print("configs", n_cfg, "bad", bad, "secs", round(time.time() - t0, 1))
