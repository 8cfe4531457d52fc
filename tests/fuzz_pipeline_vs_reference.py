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
    shape = os.environ.get("FUZZ_SHAPES") and rng.choice(["normal", "planar", "line", "duplicates", "tiny"])
    if shape == "planar":                       # degenerate geometry: rarely taken branches of the base selection
        d["P"][:, 2] = 0
        d["Q"][:, 2] = np.float32(0.1)
    elif shape == "line":
        t = np.linspace(-1, 1, len(d["P"]), dtype=np.float32)
        d["P"] = np.stack([t, 0.3 * t, -0.2 * t], 1).astype(np.float32) + np.float32(1e-3) * rng.standard_normal((len(t), 3)).astype(np.float32)
        d["Pn"] = d["Pn"][:len(t)] if d["Pn"] is not None else None
    elif shape == "duplicates":
        d["P"][::3] = d["P"][0]
        d["Q"][::4] = d["Q"][1]
    elif shape == "tiny":
        k = int(rng.randint(4, 24))
        d["P"], d["Q"] = d["P"][:k].copy(), d["Q"][:k].copy()
        if d["Pn"] is not None:
            d["Pn"], d["Qn"] = d["Pn"][:k].copy(), d["Qn"][:k].copy()
    kw = dict(delta=float(rng.choice([0.01, 0.02, 0.04])), overlap=ov, sample_size=int(rng.choice([40, 80, 150])),
              max_time_seconds=10000, random_seed=int(rng.randint(0, 2 ** 31 - 1)),
              terminate_threshold=float(rng.choice([1.0, 1.0, max(ov, 0.8)])))
    if normals and rng.randint(0, 2):
        kw["max_normal_difference"] = float(rng.choice([20.0, 45.0]))
    if rng.randint(0, 3) == 0:
        kw["max_translation_distance"] = 3.0
    rgbP = rgbQ = None
    if rng.randint(0, 4) == 0:                       # colours + colour filter
        rgbP = rng.uniform(0, 255, size=d["P"].shape).astype(np.float32)
        rgbQ = rng.uniform(0, 255, size=d["Q"].shape).astype(np.float32)
        kw["max_color_distance"] = float(rng.choice([150.0, 250.0]))
    if rng.randint(0, 4) == 0:
        kw["max_angle"] = float(rng.choice([60.0, 120.0]))
    opt = oref.make_options(**kw)
    lanes = int(rng.choice([1, 2, 3, 5]))
    # S4G_SHIM_REFERENCE_ORDER=1 (pairs in the reference's emission order): one lane, because every lane context has its own
    # id-array history; then NO difference is expected at all, without it only equal-count ties may differ
    os.environ["S4PCS_LANES"] = "1" if os.environ.get("S4G_SHIM_REFERENCE_ORDER") else str(lanes)
    os.environ["S4PCS_FUSED"] = str(int(rng.choice([1, 1, 0])))
    if os.environ.get("FUZZ_DEVICES"):           # candidate sharding over 1..4 device contexts per lane (S4PCS_DEVICES)
        os.environ["S4PCS_DEVICES"] = str(int(rng.choice([1, 2, 3, 4])))
        if os.environ.get("FUZZ_NCCL"):          # ... reduced inside the library (the stand-in's communicator) on half of them
            os.environ["S4PCS_NCCL"] = str(int(rng.randint(0, 2)))
    if os.environ.get("FUZZ_TRACE"):
        # per-iteration visitor reports (fraction, best LCP, global transform) of the whole RANSAC loop, not only its result
        a = oref.compute_transformation_traced(d["P"], d["Q"], opt, max_trace=20000)
        b = oref.compute_transformation_traced(d["P"], d["Q"], opt, libpath=_build.DROPIN_SO, max_trace=20000)
    else:
        a = oref.compute_transformation(d["P"], d["Q"], opt, Pn=d["Pn"], Qn=d["Qn"], Prgb=rgbP, Qrgb=rgbQ)
        b = oref.compute_transformation(d["P"], d["Q"], opt, Pn=d["Pn"], Qn=d["Qn"], Prgb=rgbP, Qrgb=rgbQ, libpath=_build.DROPIN_SO)
    if os.environ.get("FUZZ_TRACE"):
        # reports issued before the first adopted candidate carry a global transform built from members the reference never
        # initialises (qcentroid1_/2_, match4pcsBase.h:137): only their (fraction, best LCP) columns are comparable
        ta, tb = a[2], b[2]
        eye = np.eye(4, dtype=np.float32)[:3, :3].T.reshape(-1)            # rotation block still the identity = nothing adopted yet
        adopted = np.array([not np.array_equal(r[2:].reshape(4, 4)[:3, :3].reshape(-1), eye) for r in ta], bool) if len(ta) else np.zeros(0, bool)
        same = np.float32(a[0]) == np.float32(b[0]) and ta.shape == tb.shape and \
            np.array_equal(ta[:, :2].view(np.uint32), tb[:, :2].view(np.uint32)) and \
            np.array_equal(ta[adopted].view(np.uint32), tb[adopted].view(np.uint32)) and \
            (not adopted.any() or np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)))
    else:
        same = np.float32(a[0]) == np.float32(b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and \
            np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    if not same:
        bad += 1
        print("DIFF", n, kw, os.environ["S4PCS_LANES"], os.environ["S4PCS_FUSED"], os.environ.get("S4PCS_DEVICES"), a[0], b[0], flush=True)
        np.savez("/tmp/fuzz_fail_%d.npz" % bad, P=d["P"], Q=d["Q"], Pn=d["Pn"] if d["Pn"] is not None else np.zeros((0, 3), np.float32),
                 Qn=d["Qn"] if d["Qn"] is not None else np.zeros((0, 3), np.float32), kw=repr(kw),
                 lanes=os.environ["S4PCS_LANES"], fused=os.environ["S4PCS_FUSED"])
    n_cfg += 1
print("configs", n_cfg, "bad", bad, "secs", round(time.time() - t0, 1))
