"""Shared scenario builders for the parity tests."""
import functools
import numpy as np

from super4pcs_b200 import synth


@functools.lru_cache(maxsize=8)
def scenario(n, overlap, delta, seed=1, noise=0.0, outliers=0.0, normals=False):
    """centred clouds + candidate transforms (column-major 16) in the centred frames"""
    d = synth.make_pair(n, overlap, seed=seed, noise_sigma=noise, outlier_frac=outliers, with_normals=normals)
    P, cp = synth.center(d["P"])
    Q, cq = synth.center(d["Q"])
    return dict(P=P, Q=Q, Pn=d["Pn"], Qn=d["Qn"], cp=cp, cq=cq, delta=delta, raw=d)


def candidates_colmajor(sc, K, seed=7, n_near=None):
    n_near = max(1, K // 8) if n_near is None else n_near
    M = synth.candidate_transforms(K, sc["delta"], seed=seed, n_near=n_near, centroid_p=sc["cp"], centroid_q=sc["cq"])
    return np.ascontiguousarray(M.transpose(0, 2, 1)).reshape(K, 16)


def random_quads(nq, K, seed=0):
    return np.random.RandomState(seed).randint(0, nq, size=(K, 4)).astype(np.int32)


def congruent_like_quads(sc, base_ids, K, seed=0):
    """quads of Q whose first three points are near-congruent to the base triangle (built from the
    ground truth + random ones), so that a useful fraction passes the rms gate"""
    rng = np.random.RandomState(seed)
    P, Q = sc["P"], sc["Q"]
    gt = sc["raw"]["gt"]
    # map base points into Q's centred frame: q = R^T (p + cp - t) - cq
    R, t = gt[:3, :3], gt[:3, 3]
    out = np.empty((K, 4), np.int32)
    tgt = ((P[base_ids].astype(np.float64) + sc["cp"]) - t) @ R - sc["cq"]
    from scipy.spatial import cKDTree
    tree = cKDTree(Q)
    for k in range(K):
        if k % 2 == 0:
            jitter = tgt + rng.standard_normal(tgt.shape) * sc["delta"] * (0.2 + 2.0 * rng.random_sample())
            _, idx = tree.query(jitter)
            out[k] = idx
        else:
            out[k] = rng.randint(0, len(Q), 4)
    return out


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def read_ply_xyz(path):
    """vertex positions of a PLY written by IOManager::WritePly (binary little endian: float x y z [nx ny nz]
    [uchar r g b]) or of an ascii PLY"""
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode().splitlines()
    n = next(int(ln.split()[2]) for ln in header if ln.startswith("element vertex"))
    props = [ln.split()[1:] for ln in header if ln.startswith("property")]
    if any(ln.startswith("format ascii") for ln in header):
        rows = raw[end:].decode().splitlines()[:n]
        return np.array([r.split()[:3] for r in rows], np.float32)
    stride = sum(4 if t == "float" else 1 for t, _ in props)
    body = np.frombuffer(raw[end:end + n * stride], np.uint8).reshape(n, stride)
    return np.ascontiguousarray(body[:, :12]).view("<f4").reshape(n, 3)
