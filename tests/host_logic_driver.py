"""Subprocess body of tests/test_host_logic_cpu.py and tests/test_zz_lanes_gpu.py: drives the header-compatible C++
layer through the drop-in harness (oracle/_dropin/libb200_harness.so) and prints one JSON line.  Run with
LD_PRELOAD=tests/_build/libs4g_oracle_shim.so it exercises the HOST logic on the CPU oracle; run without, the real
CUDA library.  S4PCS_LANES / S4PCS_FUSED are read by the C++ layer from the environment."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _build, ref as oref  # noqa: E402


def _h(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run(which, libpath):
    gold = os.path.join(ROOT, "tests", "golden")
    out = {}
    if which == "hippo":
        h = np.load(os.path.join(gold, "hippo.npz"))
        opt = oref.make_options(delta=0.01, overlap=0.7, sample_size=200, max_time_seconds=1000)
        score, T, Qt = oref.compute_transformation(h["P"], h["Q"], opt, libpath=libpath)
        out = dict(score=float(np.float32(score)), T=[int(x) for x in T.view(np.uint32)], Q=_h(Qt))
    elif which == "trace":
        h = np.load(os.path.join(gold, "hippo.npz"))
        opt = oref.make_options(delta=0.01, overlap=0.7, sample_size=200, max_time_seconds=1000, random_seed=77)
        score, T, tr = oref.compute_transformation_traced(h["P"], h["Q"], opt, libpath=libpath)
        out = dict(score=float(np.float32(score)), T=[int(x) for x in T.view(np.uint32)], n=len(tr), trace=_h(tr))
    elif which == "steps":
        # stepwise use (Meshlab pattern) + early termination in the middle of a speculative batch, then the RNG /
        # base state is probed by selecting the next base by hand
        h = np.load(os.path.join(gold, "hippo.npz"))
        opt = oref.make_options(delta=0.01, sample_size=200, overlap=0.5, terminate_threshold=0.55, random_seed=3,
                                max_time_seconds=1000)
        m = oref.RefMatcher(h["P"], h["Q"], opt, identity_sampler=False, libpath=libpath)
        log = []
        for n in (1, 3, 2, 7, 40, 100):
            r = m.perform_n_steps(n)
            ok, i1, i2, ids = m.select_quadrilateral()
            bx, _, _ = m.base3d()
            log.append([r["ret"], float(np.float32(r["best_lcp"])), r["n_progress"], _h(r["T"]), bool(ok),
                        float(np.float32(i1)), float(np.float32(i2)), [int(x) for x in ids], _h(bx)])
        out = dict(log=log)
        m.close()
    else:
        raise SystemExit("unknown scenario " + which)
    return out


if __name__ == "__main__":
    which, target = sys.argv[1], sys.argv[2]
    libpath = _build.DROPIN_SO if target == "dropin" else None     # None = the compiled reference
    print("RESULT " + json.dumps(run(which, libpath)))
