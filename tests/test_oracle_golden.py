"""CPU tests (no GPU): the oracle port (oracle/port.cc) against the golden vectors generated
from the unmodified reference (tests/golden/make_golden.py), and -- when oracle/_ref is
present -- directly against the reference on fresh seeded inputs."""
import os

import numpy as np
import pytest

from oracle import port as oport
from oracle import ref as oref
from tests import common

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def stages():
    return dict(np.load(os.path.join(GOLD, "stages.npz")))


@pytest.fixture(scope="module")
def hippo():
    return dict(np.load(os.path.join(GOLD, "hippo_result.npz")))


def _check_stage_dump(g, prefix, delta, n_bases=3):
    P, Q = g[prefix + "P"], g[prefix + "Q"]
    pt = oport.Port(P, Q, delta)
    for b in range(n_bases):
        k = lambda s: g["%sb%d_%s" % (prefix, b, s)]  # noqa: E731
        bx, inv, d = k("base_xyz"), k("inv"), k("d")
        b9 = lambda i: np.concatenate([bx[i], [0, 0, 0], [-1, -1, -1]]).astype(np.float32)  # noqa: E731
        p1 = pt.extract_pairs(d[0], 0.0, 2 * delta, b9(0), b9(1))
        p2 = pt.extract_pairs(d[1], 0.0, 2 * delta, b9(2), b9(3))
        assert np.array_equal(p1, k("pairs1")) and np.array_equal(p2, k("pairs2"))
        quads = pt.find_quads(inv[0], inv[1], 2 * delta, bx, p1, p2)
        assert np.array_equal(quads, k("quads"))
        nr = len(k("rigid_rms"))
        T, rms, ok = pt.rigid_batch(k("ids"), quads[:nr])
        assert np.array_equal(ok, k("rigid_ok"))
        assert np.array_equal(_bits(rms), _bits(k("rigid_rms")))
        assert np.array_equal(_bits(T), _bits(k("rigid_T")))
        if len(k("verify_lcp")):
            lcp, good, _ = pt.verify_batch(k("verify_T"), 0.0, nthreads=oport.num_threads())
            assert np.array_equal(lcp, k("verify_lcp"))
            assert np.array_equal(good, pt.verify_bruteforce(k("verify_T")))
        r = pt.try_congruent_set(k("ids"), quads, best_lcp_in=float(k("tcs_best_before")))
        assert r["n_gate"] == int(k("tcs_n_gate"))
        assert np.float32(r["best_lcp"]) == np.float32(k("tcs_best_lcp"))
        if r["best_index"] >= 0:
            assert np.array_equal(_bits(r["T"]), _bits(k("tcs_T")))
            assert np.array_equal(quads[r["best_index"]], k("tcs_congruent"))


def test_port_matches_golden_synthetic_stages(stages):
    _check_stage_dump(stages, "", float(stages["delta"]))


def test_port_matches_golden_hippo_stages(hippo):
    _check_stage_dump(hippo, "stage_", 0.01)


def test_port_centering_matches_golden(stages):
    from super4pcs_b200 import synth
    Pc, _ = synth.center(stages["raw_P"])
    Qc, _ = synth.center(stages["raw_Q"])
    assert np.array_equal(_bits(Pc), _bits(stages["P"])) and np.array_equal(_bits(Qc), _bits(stages["Q"]))


needs_ref = pytest.mark.skipif(not oref.available(), reason="oracle/_ref (compiled reference) not present")


@needs_ref
def test_reference_reproduces_golden_hippo(hippo):
    """the committed golden equals what the compiled reference produces now (cfg0, run-example.sh:68)"""
    h = np.load(os.path.join(GOLD, "hippo.npz"))
    opt = oref.make_options(delta=0.01, overlap=0.7, sample_size=200, max_time_seconds=1000)
    score, T, _ = oref.compute_transformation(h["P"], h["Q"], opt)
    assert np.float32(score) == hippo["score"] == np.float32(0.64)
    assert np.array_equal(_bits(T), _bits(hippo["T_colmajor"]))
    # the matrix SURVEY.md 8(c) recorded from the reference's own demo binary
    want = np.array([[0.739900, 0.062655, -0.669793, -0.097583], [-0.104949, 0.994213, -0.022932, -0.005567],
                     [0.664480, 0.087262, 0.742194, -0.032100], [0, 0, 0, 1]], np.float32)
    assert np.abs(T.reshape(4, 4).T - want).max() < 1e-6


@needs_ref
@pytest.mark.parametrize("n,delta,seed,filt", [
    (1200, 0.03, 21, {}),
    (900, 0.05, 22, dict(max_normal_difference=25.0, max_angle=70.0, max_translation_distance=2.0)),
])
def test_port_vs_reference_fresh_inputs(n, delta, seed, filt):
    from super4pcs_b200 import synth
    normals = bool(filt)
    s = synth.make_pair(n, 0.5, seed=seed, with_normals=normals)
    opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.5, random_seed=seed, **filt)
    m = oref.RefMatcher(s["P"], s["Q"], opt, Pn=s["Pn"], Qn=s["Qn"])
    P, Pn, _ = m.sampled_p()
    Q, Qn, Qrgb = m.sampled_q()
    pt = oport.Port(P, Q, delta, Qn=Qn if normals else None)
    f4 = (filt.get("max_normal_difference", -1), filt.get("max_translation_distance", -1),
          filt.get("max_angle", -1), filt.get("max_color_distance", -1))
    for _ in range(3):
        ok, inv1, inv2, ids = m.select_quadrilateral()
        assert ok
        bx, bn, brgb = m.base3d()
        b9 = lambda i: np.concatenate([bx[i], bn[i], brgb[i]]).astype(np.float32)  # noqa: E731
        en = lambda v: np.sqrt(np.float32(v[0] * v[0]) + (np.float32(v[1] * v[1]) + np.float32(v[2] * v[2])))  # noqa: E731
        d1, d2 = en(bx[0] - bx[1]), en(bx[2] - bx[3])
        a1, a2 = en(bn[0] - bn[1]), en(bn[2] - bn[3])
        p1r, p2r = m.extract_pairs(d1, a1, 2 * delta, 0, 1), m.extract_pairs(d2, a2, 2 * delta, 2, 3)
        p1 = pt.extract_pairs(d1, a1, 2 * delta, b9(0), b9(1), f4)
        p2 = pt.extract_pairs(d2, a2, 2 * delta, b9(2), b9(3), f4)
        assert np.array_equal(p1, p1r) and np.array_equal(p2, p2r)
        qr = m.find_quads(inv1, inv2, 2 * delta, 2 * delta, p1r, p2r)
        assert np.array_equal(pt.find_quads(inv1, inv2, 2 * delta, bx, p1, p2), qr)
        if len(qr):
            Tr, rr, okr = m.rigid_batch(ids, qr[:3000])
            Tp, rp, okp = pt.rigid_batch(ids, qr[:3000], max_angle_deg=filt.get("max_angle", -1.0))
            assert np.array_equal(okr, okp) and np.array_equal(_bits(rr), _bits(rp)) and np.array_equal(_bits(Tr), _bits(Tp))


@needs_ref
def test_port_verify_vs_reference_with_early_exit():
    sc = common.scenario(4000, 0.4, 0.02, seed=5)
    opt = oref.make_options(delta=0.02, sample_size=10 ** 8, overlap=0.4)
    m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], opt)
    T = common.candidates_colmajor(sc, 48)
    pt = oport.Port(sc["P"], sc["Q"], 0.02)
    for best in (0.0, 0.05, 0.3):
        lr, _ = m.verify_batch(T, best)
        lp, _, _ = pt.verify_batch(T, best)
        assert np.array_equal(lr, lp)


@needs_ref
def test_port_quads_nearly_opposite_quaternion_branch_vs_reference():
    """hundreds of query directions within 0.0045 rad of -z take Eigen's nearly-opposite branch of
    setFromTwoVectors (JacobiSVD of a 2x3): the port's Householder restatement must give the same quads"""
    rng = np.random.RandomState(5)
    n, delta = 1500, 0.03
    base = rng.uniform(-1, 1, size=(n // 2, 3)).astype(np.float32)
    up = base + np.array([0, 0, 0.9], np.float32) + rng.normal(0, 2e-4, size=base.shape).astype(np.float32)
    Q = np.concatenate([base, up]).astype(np.float32)
    opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.5)
    m = oref.RefMatcher(Q, Q, opt)
    Qs, _, _ = m.sampled_q()
    pt = oport.Port(Qs, Qs, delta)
    bx = np.array([[0, 0, 0], [0.02, 0.01, -0.9], [0.3, 0, 0.1], [0.29, 0.01, -0.8]], np.float32)
    m.set_base3d(bx)
    p1 = m.extract_pairs(0.9, 0.0, 2 * delta, 0, 1)
    p2 = m.extract_pairs(0.9, 0.0, 2 * delta, 2, 3)
    assert np.array_equal(p1, pt.extract_pairs(0.9, 0.0, 2 * delta))
    g, ratio = pt.normalization()
    U = (Qs - g) / ratio + 0.5
    d = U[p2[:, 1]] - U[p2[:, 0]]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    assert (d[:, 2] < -1 + 1e-5).sum() > 100               # the branch really is exercised
    for inv1, inv2 in ((0.5, 0.5), (0.3, 0.7)):
        qr = m.find_quads(inv1, inv2, 2 * delta, 2 * delta, p1, p2)
        assert len(qr) > 1000
        assert np.array_equal(qr, pt.find_quads(inv1, inv2, 2 * delta, bx, p1, p2))


def _sphere_cloud(n, seed):
    """unit vectors like the reference's Testing::generateSphereCloud (tests/testing.h:158-168)"""
    v = np.random.RandomState(seed).uniform(-1, 1, size=(n, 3)).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _bruteforce_pairs(Q, d, eps):
    """the reference test's ground truth (tests/testing.h:172-194): ordered pairs with |‖qi - qj‖ - d| <= eps"""
    out = []
    for j in range(len(Q)):
        dist = np.linalg.norm(Q[j + 1:] - Q[j], axis=1)
        for i in np.nonzero(np.abs(dist - np.float32(d)) <= np.float32(eps))[0] + j + 1:
            out += [(j, int(i)), (int(i), j)]
    return np.array(sorted(out), np.int32).reshape(-1, 2)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_known_answer_of_the_reference_pair_extraction_test(seed):
    """restatement of the reference's own ExtractPairs test (tests/pair_extraction.cc:239-314): sphere clouds of
    200 / 150 points, delta = 0.1, distances 0.3 / 0.5, epsilon = 2 delta: sorted output == brute force.  Run for
    the port and, when present, for the compiled reference (whose test this is)."""
    P, Q = _sphere_cloud(200, seed), _sphere_cloud(150, 100 + seed)
    delta = 0.1
    opt = oref.make_options(delta=delta, overlap=0.5, sample_size=10 ** 8)
    m = oref.RefMatcher(P, Q, opt) if oref.available() else None
    Qc = m.sampled_q()[0] if m else Q - Q.mean(axis=0, dtype=np.float32)
    Pc = m.sampled_p()[0] if m else P - P.mean(axis=0, dtype=np.float32)
    pt = oport.Port(Pc, Qc, delta)
    for d, ang in ((0.3, 0.6), (0.5, 0.4)):
        want = _bruteforce_pairs(Q, d, 2 * delta)
        assert len(want) > 100
        got = pt.extract_pairs(d, ang, 2 * delta)
        assert np.array_equal(got, want)
        if m:
            assert np.array_equal(m.extract_pairs(d, ang, 2 * delta, 0, 1), want)


def _same_floats(a, b):
    """bit-equal, with NaNs equal to NaNs (the sign / payload of a NaN is not part of the contract)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(_bits(a[~na]), _bits(b[~nb]))


@needs_ref
def test_port_rigid_degenerate_quads_vs_reference():
    """coincident, collinear and tiny frames in the base and in the candidates (match4pcsBase.cc:365-500: the
    'true with rms = kLargeNumber' and the scale-check exits)"""
    rng = np.random.RandomState(9)
    Q = rng.uniform(-1, 1, size=(60, 3)).astype(np.float32)
    Q[10] = Q[11]                                              # coincident points
    Q[20:24] = Q[20] + np.outer(np.arange(4), np.array([0.1, 0.2, -0.05], np.float32))   # collinear run
    Q[30] = Q[31] + np.float32(1e-7)                           # nearly coincident
    P = Q.copy()
    P[5] = P[6]
    P[40:43] = P[40] + np.outer(np.arange(3), np.array([0.3, 0.0, 0.1], np.float32))
    opt = oref.make_options(delta=0.05, sample_size=10 ** 8, overlap=0.5)
    m = oref.RefMatcher(P, Q, opt)
    Ps, Qs = m.sampled_p()[0], m.sampled_q()[0]
    pt = oport.Port(Ps, Qs, 0.05)
    quads = np.array([[10, 11, 12, 13], [10, 10, 10, 10], [20, 21, 22, 23], [30, 31, 32, 33], [1, 2, 3, 4],
                      [3, 2, 1, 0], [20, 22, 21, 5], [11, 10, 12, 13]] + rng.randint(0, 60, size=(200, 4)).tolist(), np.int32)
    produced = 0
    for base in ([0, 1, 2, 3], [5, 6, 7, 8], [40, 41, 42, 9], [7, 7, 8, 9], [1, 2, 3, 4]):
        Tr, rr, okr = m.rigid_batch(base, quads)
        Tp, rp, okp = pt.rigid_batch(base, quads)
        assert np.array_equal(okr, okp), base
        assert _same_floats(rr, rp), base
        sel = okr & (rr < 1e8)                                 # T is only defined where a transform was produced
        assert _same_floats(Tr[sel], Tp[sel]), base
        produced += int(sel.sum())
        if base in ([5, 6, 7, 8], [7, 7, 8, 9]):
            assert not sel.any()                               # first two base points coincide: rms = kLargeNumber for every quad
    assert produced > 100


@needs_ref
def test_port_verify_pathological_transforms_vs_reference():
    sc = common.scenario(3000, 0.5, 0.02, seed=8)
    m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], oref.make_options(delta=0.02, sample_size=10 ** 8, overlap=0.5))
    pt = oport.Port(sc["P"], sc["Q"], 0.02)
    T = common.candidates_colmajor(sc, 12).copy()
    T[1, 12:15] += 1e3                                         # far away: no inliers
    T[2, 12:15] += 1e30                                        # overflowing translation
    T[3, :] = 0                                                # collapses Q onto the origin
    T[4, 0] = np.nan                                           # NaN rotation entry
    T[5, 13] = np.inf
    T[6, :12] *= 1e-3                                          # shrinks Q into a ball
    lr, _ = m.verify_batch(T, 0.0)
    lp, good, _ = pt.verify_batch(T, 0.0)
    assert np.array_equal(lr, lp)
    assert good[1] == good[2] == good[4] == good[5] == 0
    assert np.array_equal(good, pt.verify_bruteforce(T))


@needs_ref
def test_port_pairs_colour_filter_and_quads_empty_lists_vs_reference():
    rng = np.random.RandomState(4)
    s = dict(P=rng.uniform(-1, 1, (500, 3)).astype(np.float32), Q=rng.uniform(-1, 1, (500, 3)).astype(np.float32))
    rgb = rng.uniform(0, 255, (500, 3)).astype(np.float32)
    opt = oref.make_options(delta=0.05, sample_size=10 ** 8, overlap=0.5, max_color_distance=120.0)
    m = oref.RefMatcher(s["P"], s["Q"], opt, Qrgb=rgb, Prgb=rgb)
    Ps, _, _ = m.sampled_p()
    Qs, _, Qrgb = m.sampled_q()
    pt = oport.Port(Ps, Qs, 0.05, Qrgb=Qrgb)
    bx = np.array([[0, 0, 0], [0.5, 0.1, 0], [0.2, 0.4, 0.1], [0.1, -0.3, 0.2]], np.float32)
    brgb = np.array([[200, 30, 30], [30, 200, 30], [10, 10, 250], [128, 128, 128]], np.float32)
    m.set_base3d(bx, rgb=brgb)
    b9 = lambda i: np.concatenate([bx[i], [0, 0, 0], brgb[i]]).astype(np.float32)  # noqa: E731
    f4 = (-1, -1, -1, 120.0)
    pr = m.extract_pairs(0.6, 0.0, 0.1, 0, 1)
    pp = pt.extract_pairs(0.6, 0.0, 0.1, b9(0), b9(1), f4)
    nofilter = pt.extract_pairs(0.6, 0.0, 0.1, b9(0), b9(1))
    assert np.array_equal(pr, pp) and 0 < len(pp) < len(nofilter)          # the colour filter really rejects pairs
    empty = np.zeros((0, 2), np.int32)
    for a, b in ((empty, pp), (pp, empty), (empty, empty)):
        assert len(m.find_quads(0.5, 0.5, 0.1, 0.1, a, b)) == 0
        assert len(pt.find_quads(0.5, 0.5, 0.1, bx, a, b)) == 0


@needs_ref
@pytest.mark.parametrize("seed", [1, 2])
def test_port_pair_emission_order_vs_reference(seed):
    """The reference does NOT sort its pairs (super4pcs.cc:183-224): FindCongruentQuadrilaterals, and with it the
    candidate order and the winner among candidates with equal inlier counts, sees them in the emission order of the
    octree traversal, which depends on the persistent, in-place partitioned id array (intersectionFunctor.h:104-236,
    intersectionNode.h:165-249).  port_extract_pairs_ordered restates that traversal: same SEQUENCE, call after call."""
    from super4pcs_b200 import synth
    rng = np.random.RandomState(seed)
    calls = 0
    for _ in range(8):
        normals = bool(rng.randint(0, 2))
        d = synth.make_pair(int(rng.randint(60, 900)), 0.6, seed=int(rng.randint(1, 10 ** 6)), with_normals=normals)
        delta = float(rng.choice([0.01, 0.02, 0.04, 0.08]))
        filt = {}
        if normals and rng.randint(0, 2):
            filt["max_normal_difference"] = 30.0
        if rng.randint(0, 3) == 0:
            filt["max_angle"] = 60.0
        opt = oref.make_options(delta=delta, sample_size=int(rng.choice([40, 150, 400, 10 ** 6])), overlap=0.6,
                                random_seed=int(rng.randint(1, 10 ** 6)), **filt)
        m = oref.RefMatcher(d["P"], d["Q"], opt, Pn=d["Pn"], Qn=d["Qn"], identity_sampler=False)
        P, _, _ = m.sampled_p()
        Q, Qn, _ = m.sampled_q()
        pt = oport.Port(P, Q, delta, Qn=Qn if normals else None)
        f4 = (filt.get("max_normal_difference", -1), -1, filt.get("max_angle", -1), -1)
        en = lambda v: np.sqrt(np.float32(v[0] * v[0]) + (np.float32(v[1] * v[1]) + np.float32(v[2] * v[2])))  # noqa: E731
        for _b in range(5):
            ok, _, _, _ = m.select_quadrilateral()
            if not ok:
                continue
            bx, bn, brgb = m.base3d()
            b9 = lambda i: np.concatenate([bx[i], bn[i], brgb[i]]).astype(np.float32)  # noqa: E731
            for i0, i1 in ((0, 1), (2, 3)):
                dd, aa = en(bx[i0] - bx[i1]), en(bn[i0] - bn[i1])
                want = m.extract_pairs(dd, aa, 2 * delta, i0, i1, sort=False)
                got = pt.extract_pairs_ordered(dd, aa, 2 * delta, b9(i0), b9(i1), f4)
                assert np.array_equal(got, want)
                calls += 1
                if len(want) > 4:
                    sorted_want = want[np.lexsort((want[:, 1], want[:, 0]))]
                    assert np.array_equal(pt.extract_pairs(dd, aa, 2 * delta, b9(i0), b9(i1), f4), sorted_want)
    assert calls > 40
