"""GPU parity of a4 + a5 (IndexedNormalSet / FindCongruentQuadrilaterals) and of the whole
per-base chain ExtractPairs x2 -> FindCongruentQuadrilaterals -> TryCongruentSet through the
C ABI, against the golden vectors of the unmodified reference and against the oracle port."""
import os

import numpy as np
import pytest

from oracle import port as oport
from oracle import ref as oref
from tests import common

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx(s4g_lib):
    from super4pcs_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _b9(bx, i):
    return np.concatenate([bx[i], [0, 0, 0], [-1, -1, -1]]).astype(np.float32)


def _chain_against_golden(ctx, g, prefix, delta, n_bases=3):
    P, Q = g[prefix + "P"], g[prefix + "Q"]
    ctx.set_cloud_p(P, delta)
    ctx.set_cloud_q(Q)
    for b in range(n_bases):
        k = lambda s: g["%sb%d_%s" % (prefix, b, s)]  # noqa: E731
        bx, inv, d = k("base_xyz"), k("inv"), k("d")
        p1 = ctx.extract_pairs(d[0], 0.0, 2 * delta, _b9(bx, 0), _b9(bx, 1), slot=0)
        p2 = ctx.extract_pairs(d[1], 0.0, 2 * delta, _b9(bx, 2), _b9(bx, 3), slot=1)
        assert np.array_equal(p1, k("pairs1")) and np.array_equal(p2, k("pairs2"))       # bit-exact sets
        quads = ctx.find_quads(inv[0], inv[1], 2 * delta, bx)
        assert np.array_equal(quads, k("quads"))                                         # bit-exact, same order
        # a7 on the resident quads: winner, LCP, transform
        r = ctx.try_congruent_set_resident(P[k("ids")], 2 * delta)
        assert r["n_gate_pass"] == int(k("tcs_n_gate"))
        best_before = np.float32(k("tcs_best_before"))
        lcp = np.float32(r["best_count"]) / np.float32(r["n_q"])
        if lcp > best_before:
            assert lcp == np.float32(k("tcs_best_lcp"))
            assert np.array_equal(quads[r["best_index"]], k("tcs_congruent"))
            T = r["T"].reshape(4, 4).T
            Tg = k("tcs_T").reshape(4, 4).T
            assert np.linalg.norm(T - Tg) <= 1e-5                                        # north_star tolerance
            assert np.array_equal(common.bits(r["T"]), common.bits(k("tcs_T")))          # in fact bit-exact
        else:
            assert np.float32(k("tcs_best_lcp")) == best_before
        # verify LCPs of the golden's gate-passing transforms
        if len(k("verify_lcp")):
            got = ctx.verify(k("verify_T")).astype(np.float32) / np.float32(len(Q))
            assert np.array_equal(got, k("verify_lcp"))


def test_chain_matches_golden_synthetic(ctx):
    g = dict(np.load(os.path.join(GOLD, "stages.npz")))
    _chain_against_golden(ctx, g, "", float(g["delta"]))


def test_chain_matches_golden_hippo(ctx):
    g = dict(np.load(os.path.join(GOLD, "hippo_result.npz")))
    _chain_against_golden(ctx, g, "stage_", 0.01)


@pytest.mark.parametrize("n,delta,seed", [(3000, 0.01, 2), (2000, 0.02, 1), (6000, 0.006, 9)])
def test_quads_match_port_on_fresh_bases(ctx, n, delta, seed):
    sc = common.scenario(n, 0.5, delta, seed=seed)
    ctx.set_cloud_p(sc["P"], delta)
    ctx.set_cloud_q(sc["Q"])
    pt = oport.Port(sc["P"], sc["Q"], delta)
    rng = np.random.RandomState(seed)
    P = sc["P"]
    tried = 0
    while tried < 3:
        ids = rng.choice(len(P), 4, replace=False)
        bx = P[ids]
        d1 = float(np.linalg.norm(bx[0] - bx[1]))
        d2 = float(np.linalg.norm(bx[2] - bx[3]))
        if min(d1, d2) < 0.3:
            continue
        tried += 1
        inv1, inv2 = float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.2, 0.8))
        p1 = ctx.extract_pairs(d1, 0.0, 2 * delta, _b9(bx, 0), _b9(bx, 1), slot=0)
        p2 = ctx.extract_pairs(d2, 0.0, 2 * delta, _b9(bx, 2), _b9(bx, 3), slot=1)
        want = pt.find_quads(inv1, inv2, 2 * delta, bx, p1, p2)
        got = ctx.find_quads(inv1, inv2, 2 * delta, bx)
        assert np.array_equal(got, want)


def test_quads_degenerate_directions(ctx):
    """query directions (anti)parallel to z exercise Eigen's nearly-opposite quaternion branch"""
    rng = np.random.RandomState(5)
    n, delta = 1500, 0.03
    base = rng.uniform(-1, 1, size=(n // 2, 3)).astype(np.float32)
    up = base + np.array([0, 0, 0.9], np.float32) + rng.normal(0, 2e-4, size=base.shape).astype(np.float32)
    Q = np.concatenate([base, up]).astype(np.float32)
    Q -= Q.mean(0)
    ctx.set_cloud_p(Q, delta)
    ctx.set_cloud_q(Q)
    pt = oport.Port(Q, Q, delta)
    bx = np.array([[0, 0, 0], [0.02, 0.01, -0.9], [0.3, 0, 0.1], [0.29, 0.01, -0.8]], np.float32)
    p1 = ctx.extract_pairs(0.9, 0.0, 2 * delta, _b9(bx, 0), _b9(bx, 1), slot=0)
    p2 = ctx.extract_pairs(0.9, 0.0, 2 * delta, _b9(bx, 2), _b9(bx, 3), slot=1)
    assert len(p1) > 1000
    want = pt.find_quads(0.5, 0.5, 2 * delta, bx, p1, p2)
    got = ctx.find_quads(0.5, 0.5, 2 * delta, bx)
    assert len(want) > 0 and np.array_equal(got, want)
    if oref.available():
        opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.5)
        m = oref.RefMatcher(Q, Q, opt)       # centres again: Q is already centred -> identical up to rounding
        Qs, _, _ = m.sampled_q()
        if np.array_equal(Qs, Q):
            m.set_base3d(bx)
            assert np.array_equal(m.find_quads(0.5, 0.5, 2 * delta, 2 * delta, p1, p2), got)


def test_quads_edge_cases(ctx):
    sc = common.scenario(2000, 0.5, 0.02, seed=1)
    ctx.set_cloud_p(sc["P"], 0.02)
    ctx.set_cloud_q(sc["Q"])
    bx = sc["P"][[1, 50, 100, 200]]
    # empty pair lists -> no quads
    ctx.set_pairs(0, np.zeros((0, 2), np.int32))
    ctx.set_pairs(1, np.zeros((0, 2), np.int32))
    assert len(ctx.find_quads(0.5, 0.5, 0.04, bx)) == 0
    # parallel base segments: alpha = 0 -> zero cone samples -> no quads (normalset.hpp:176-178)
    par = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    ctx.extract_pairs(1.0, 0, 0.04, slot=0, fetch=False)
    ctx.extract_pairs(1.0, 0, 0.04, slot=1, fetch=False)
    assert len(ctx.find_quads(0.5, 0.5, 0.04, par)) == 0
    # uploaded (unsorted) lists: ids index into the caller's order, like the reference's std::set<(id,i)>
    p1 = ctx.extract_pairs(0.9, 0, 0.04, slot=0)
    p2 = ctx.extract_pairs(0.7, 0, 0.04, slot=1)
    rng = np.random.RandomState(0)
    s1, s2 = rng.permutation(len(p1)), rng.permutation(len(p2))
    ctx.set_pairs(0, p1[s1])
    ctx.set_pairs(1, p2[s2])
    got = ctx.find_quads(0.4, 0.6, 0.04, bx)
    pt = oport.Port(sc["P"], sc["Q"], 0.02)
    assert np.array_equal(got, pt.find_quads(0.4, 0.6, 0.04, bx, p1[s1], p2[s2]))
