"""Seeded synthetic clouds for the parity tests and bench.py.

A bumpy sphere  x = (1 + 0.15 sin(5 dx) cos(4 dy) + 0.1 sin(7 dz)) d  sampled at random unit
directions d; P keeps the slab dz < z_c, Q keeps dz > -z_c (overlap fraction of each cloud is
2 z_c / (1 + z_c)); Q is then mapped by the INVERSE of a fixed ground-truth rigid motion, so
that  P ~ GT(Q)  on the overlap.  Optional isotropic Gaussian noise on Q, uniform bbox outliers
on both clouds, analytic normals.
"""
import numpy as np

GT_AXIS = np.array([0.3, 1.0, 0.2], dtype=np.float64)
GT_ANGLE = 0.7
GT_T = np.array([0.4, -0.2, 0.3], dtype=np.float64)


def gt_transform():
    """4x4 float64 ground-truth motion taking Q onto P."""
    a = GT_AXIS / np.linalg.norm(GT_AXIS)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(GT_ANGLE) * K + (1 - np.cos(GT_ANGLE)) * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = GT_T
    return T


def zc_for_overlap(overlap):
    """slab half-width giving the requested overlap fraction  2 z / (1 + z)."""
    return overlap / (2.0 - overlap)


def _radius(d):
    return 1.0 + 0.15 * np.sin(5 * d[:, 0]) * np.cos(4 * d[:, 1]) + 0.1 * np.sin(7 * d[:, 2])


def _surface(rng, n, keep):
    out = []
    got = 0
    while got < n:
        d = rng.standard_normal((max(1024, 2 * (n - got)), 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = d[keep(d[:, 2])]
        out.append(d)
        got += len(d)
    d = np.concatenate(out)[:n]
    return d, d * _radius(d)[:, None]


def _normals(d, x):
    """gradient of the implicit F(x) = |x| - r(x/|x|), evaluated numerically in float64."""
    eps = 1e-6
    n = np.empty_like(x)
    for k in range(3):
        dx = np.zeros(3)
        dx[k] = eps

        def F(y):
            ny = np.linalg.norm(y, axis=1)
            return ny - _radius(y / ny[:, None])
        n[:, k] = (F(x + dx) - F(x - dx)) / (2 * eps)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return n


def make_pair(n, overlap, seed=42, noise_sigma=0.0, outlier_frac=0.0, with_normals=False):
    """Returns dict(P, Q, Pn, Qn, gt) -- float32 (n,3) clouds (+ normals or None), float64 gt 4x4."""
    rng = np.random.RandomState(seed)
    zc = zc_for_overlap(overlap)
    n_out = int(round(outlier_frac * n))
    n_in = n - n_out
    dP, P = _surface(rng, n_in, lambda z: z < zc)
    dQ, Qs = _surface(rng, n_in, lambda z: z > -zc)
    Pn = _normals(dP, P) if with_normals else None
    Qn = _normals(dQ, Qs) if with_normals else None
    gt = gt_transform()
    Rinv = gt[:3, :3].T
    Q = (Qs - gt[:3, 3]) @ Rinv.T          # Q = GT^-1 (surface)
    if with_normals:
        Qn = Qn @ Rinv.T
    if noise_sigma > 0:
        Q = Q + rng.standard_normal(Q.shape) * noise_sigma
    if n_out > 0:
        def outl(X):
            lo, hi = X.min(0), X.max(0)
            return lo + rng.random_sample((n_out, 3)) * (hi - lo)
        P = np.concatenate([P, outl(P)])
        Q = np.concatenate([Q, outl(Q)])
        if with_normals:
            def rn():
                v = rng.standard_normal((n_out, 3))
                return v / np.linalg.norm(v, axis=1, keepdims=True)
            Pn = np.concatenate([Pn, rn()])
            Qn = np.concatenate([Qn, rn()])
        permP, permQ = rng.permutation(n), rng.permutation(n)
        P, Q = P[permP], Q[permQ]
        if with_normals:
            Pn, Qn = Pn[permP], Qn[permQ]
    f = np.float32
    return dict(P=P.astype(f), Q=Q.astype(f),
                Pn=None if Pn is None else Pn.astype(f), Qn=None if Qn is None else Qn.astype(f), gt=gt)


def center(X):
    """float32 centring like Match4PCSBase::init (reference match4pcsBase.hpp:142-149):
    sequential float accumulation, divide, subtract."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    c = np.zeros(3, np.float32)
    # sequential float32 accumulation (np.cumsum keeps float32 and is sequential)
    c = np.cumsum(X, axis=0, dtype=np.float32)[-1] if len(X) else c
    c = (c / np.float32(len(X))).astype(np.float32)
    return (X - c).astype(np.float32), c


def candidate_transforms(K, delta, seed=7, n_near=64, centroid_p=None, centroid_q=None):
    """K candidate 4x4 (float32, row-indexed [k,r,c]) in the CENTRED frames: `n_near` small
    perturbations of the ground truth (<= 2 delta translation, <= 0.5 deg rotation) followed by
    random rigid motions."""
    rng = np.random.RandomState(seed)
    gt = gt_transform()
    cp = np.zeros(3) if centroid_p is None else np.asarray(centroid_p, np.float64)
    cq = np.zeros(3) if centroid_q is None else np.asarray(centroid_q, np.float64)

    def rot(axis, ang):
        a = axis / np.linalg.norm(axis)
        Km = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        return np.eye(3) + np.sin(ang) * Km + (1 - np.cos(ang)) * (Km @ Km)

    out = np.empty((K, 4, 4), np.float64)
    for k in range(K):
        if k < n_near:
            dR = rot(rng.standard_normal(3), np.deg2rad(0.5) * rng.random_sample())
            dt = rng.standard_normal(3)
            dt *= 2 * delta * rng.random_sample() / np.linalg.norm(dt)
            R = dR @ gt[:3, :3]
            t = dR @ gt[:3, 3] + dt
        else:
            R = rot(rng.standard_normal(3), np.pi * rng.random_sample())
            t = rng.standard_normal(3) * 0.3
        # centred frames: p - cp = R (q - cq) + (R cq + t - cp)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = R @ cq + t - cp
        out[k] = T
    return out.astype(np.float32)
