"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference
(oracle/_ref = /root/reference compiled by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Outputs (committed):
  hippo.npz         vertex positions of the reference's bundled assets hippo1.obj / hippo2.obj
                    (float32; data, not source) -- BASELINE.json configs[0]
  hippo_result.npz  reference results on hippo with scripts/run-example.sh:68 parameters
                    (-o 0.7 -d 0.01 -n 200): score, 4x4, sampled sizes, first-base stage outputs
  stages.npz        per-stage reference outputs on a small synthetic pair: base, sorted pairs,
                    quads, rigid fits, Verify LCPs, TryCongruentSet winner
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from super4pcs_b200 import synth  # noqa: E402


def read_obj_vertices(path):
    v = []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                _, x, y, z = line.split()[:4]
                v.append((np.float32(x), np.float32(y), np.float32(z)))
    return np.array(v, np.float32)


def stage_dump(m, delta, n_bases, max_rigid=4000):
    """walks `n_bases` bases of the reference matcher and records every stage"""
    out = {}
    P, _, _ = m.sampled_p()
    Q, _, _ = m.sampled_q()
    out["P"], out["Q"] = P, Q
    out["init_best_lcp"] = np.float32(m.best_lcp())
    for b in range(n_bases):
        ok, inv1, inv2, ids = m.select_quadrilateral()
        bx, _, _ = m.base3d()
        d1 = np.linalg.norm((bx[0] - bx[1]).astype(np.float32))
        d2 = np.linalg.norm((bx[2] - bx[3]).astype(np.float32))
        # the reference computes these with Eigen's norm(): x^2 + (y^2 + z^2)
        def en(v):
            v = v.astype(np.float32)
            return np.sqrt(np.float32(v[0] * v[0]) + (np.float32(v[1] * v[1]) + np.float32(v[2] * v[2])))
        d1, d2 = en(bx[0] - bx[1]), en(bx[2] - bx[3])
        p1 = m.extract_pairs(d1, 0.0, 2 * delta, 0, 1)
        p2 = m.extract_pairs(d2, 0.0, 2 * delta, 2, 3)
        quads = m.find_quads(inv1, inv2, 2 * delta, 2 * delta, p1, p2)
        pre = "b%d_" % b
        out[pre + "ok"] = np.int32(ok)
        out[pre + "inv"] = np.array([inv1, inv2], np.float32)
        out[pre + "ids"] = ids
        out[pre + "base_xyz"] = bx
        out[pre + "d"] = np.array([d1, d2], np.float32)
        out[pre + "pairs1"], out[pre + "pairs2"], out[pre + "quads"] = p1, p2, quads
        qs = quads[:max_rigid]
        T, rms, okr = m.rigid_batch(ids, qs) if len(qs) else (np.zeros((0, 16), np.float32), np.zeros(0, np.float32), np.zeros(0, bool))
        out[pre + "rigid_T"], out[pre + "rigid_rms"], out[pre + "rigid_ok"] = T, rms, okr
        gate = okr & (rms >= 0) & (rms < 2 * delta)
        Tv = T[gate][:64]
        lcp, _ = m.verify_batch(Tv, 0.0) if len(Tv) else (np.zeros(0, np.float32), 0)
        out[pre + "verify_T"], out[pre + "verify_lcp"] = Tv, lcp
        before = m.best_lcp()
        r = m.try_congruent_set(ids, quads)
        out[pre + "tcs_best_before"] = np.float32(before)
        out[pre + "tcs_best_lcp"] = np.float32(r["best_lcp"])
        out[pre + "tcs_n_gate"] = np.int64(r["n_gate"])
        out[pre + "tcs_T"] = r["T"]
        out[pre + "tcs_congruent"] = r["congruent"]
    return out


def main():
    assert ref.available(), "needs /root/reference"
    assets = os.path.join(os.environ.get("S4_REFERENCE_ROOT", "/root/reference"), "assets")
    P = read_obj_vertices(os.path.join(assets, "hippo1.obj"))
    Q = read_obj_vertices(os.path.join(assets, "hippo2.obj"))
    np.savez_compressed(os.path.join(HERE, "hippo.npz"), P=P, Q=Q)

    # cfg0: scripts/run-example.sh:68  -o 0.7 -d 0.01 -t 1000 -n 200
    opt = ref.make_options(delta=0.01, overlap=0.7, sample_size=200, max_time_seconds=1000)
    score, T, Qt = ref.compute_transformation(P, Q, opt)
    m = ref.RefMatcher(P, Q, opt, identity_sampler=False)
    st = m.init_state()
    res = dict(score=np.float32(score), T_colmajor=T, nP=np.int64(m.nP), nQ=np.int64(m.nQ),
               init_best_lcp=np.float32(st["best_lcp"]), number_of_trials=np.int64(st["number_of_trials"]),
               diameter=np.float32(st["diameter"]), centroid_p=st["centroid_p"], centroid_q=st["centroid_q"],
               Q_transformed_head=Qt[:64])
    dump = stage_dump(m, 0.01, n_bases=3)
    for k, v in dump.items():
        res["stage_" + k] = v
    np.savez_compressed(os.path.join(HERE, "hippo_result.npz"), **res)
    print("hippo: score %.6f nP %d nQ %d trials %d" % (score, m.nP, m.nQ, st["number_of_trials"]))
    print(T.reshape(4, 4).T)

    # small synthetic pair, whole clouds (identity sampler), 3 bases
    n, delta = 1000, 0.03
    d = synth.make_pair(n, 0.5, seed=11)
    opt = ref.make_options(delta=delta, overlap=0.5, sample_size=10 ** 8, random_seed=11)
    m = ref.RefMatcher(d["P"], d["Q"], opt)
    dump = stage_dump(m, delta, n_bases=3)
    dump["delta"] = np.float32(delta)
    dump["raw_P"], dump["raw_Q"] = d["P"], d["Q"]
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **dump)
    for b in range(3):
        print("synthetic base", b, "pairs", len(dump["b%d_pairs1" % b]), len(dump["b%d_pairs2" % b]),
              "quads", len(dump["b%d_quads" % b]), "gate", int(dump["b%d_tcs_n_gate" % b]),
              "best", float(dump["b%d_tcs_best_lcp" % b]))


if __name__ == "__main__":
    main()
