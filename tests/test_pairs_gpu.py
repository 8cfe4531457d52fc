"""GPU parity of a2 + a3 (ExtractPairs / PairCreationFunctor::process) through the C ABI:
ordered pair SETS bit-exact against the unmodified reference (when it travelled) and the port."""
import numpy as np
import pytest

from oracle import port as oport
from oracle import ref as oref
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(s4g_lib):
    from super4pcs_b200 import Context
    c = Context(0)
    yield c
    c.close()


CASES = [
    # n, delta, d, seed, normals, filters(max_normal_difference, max_translation_distance, max_angle, max_color)
    (3000, 0.02, 1.0, 1, False, (-1, -1, -1, -1)),
    (5000, 0.01, 0.7, 2, True, (30.0, -1, -1, -1)),
    (2000, 0.03, 1.3, 3, True, (20.0, 1.5, 60.0, -1)),
    (4097, 0.005, 0.31, 4, False, (-1, -1, -1, -1)),
    (777, 0.05, 2.0, 5, False, (-1, -1, 45.0, -1)),
]


@pytest.mark.parametrize("n,delta,d,seed,normals,filt", CASES)
def test_pairs_match_oracle(ctx, n, delta, d, seed, normals, filt):
    from super4pcs_b200 import PairFilters
    sc = common.scenario(n, 0.4, delta, seed=seed, normals=normals)
    Qn = None
    if normals:
        Qn = sc["Qn"] / np.linalg.norm(sc["Qn"], axis=1, keepdims=True)
        Qn = Qn.astype(np.float32)
    ctx.set_cloud_p(sc["P"], delta)
    ctx.set_cloud_q(sc["Q"], normals=Qn)
    rng = np.random.RandomState(seed)
    ids = rng.randint(0, n, 2)
    b1 = np.concatenate([sc["P"][ids[0]], [0.3, 0.5, 0.81], [-1, -1, -1]]).astype(np.float32)
    b2 = np.concatenate([sc["P"][ids[1]], [0.1, -0.7, 0.7], [-1, -1, -1]]).astype(np.float32)
    got = ctx.extract_pairs(d, 0.5, 2 * delta, b1, b2, PairFilters(*filt), slot=0)
    pt = oport.Port(sc["P"], sc["Q"], delta, Qn=Qn)
    want = pt.extract_pairs(d, 0.5, 2 * delta, b1, b2, filt)
    assert len(got) == len(want) and len(want) > 0
    assert np.array_equal(got, want)
    assert ctx.count_pairs(d, 2 * delta) >= len(want) if filt != (-1, -1, -1, -1) else True


def test_pairs_against_reference(ctx):
    if not oref.available():
        pytest.skip("oracle/_ref not present on this box")
    n, delta, d = 3000, 0.02, 0.9
    sc = common.scenario(n, 0.4, delta, seed=7)
    opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.4)
    m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], opt)
    want = m.extract_pairs(d, 0.0, 2 * delta, 0, 1)
    ctx.set_cloud_p(sc["P"], delta)
    ctx.set_cloud_q(sc["Q"])
    got = ctx.extract_pairs(d, 0.0, 2 * delta)
    assert np.array_equal(got, want)
    # brute force, the reference's own test criterion (tests/pair_extraction.cc:172-194,307-311)
    Q = sc["Q"].astype(np.float32)
    D = np.sqrt(((Q[:, None, :] - Q[None, :, :]) ** 2).sum(-1, dtype=np.float32))
    ii, jj = np.nonzero(np.abs(D.astype(np.float64) - np.float32(d)) <= np.float32(2 * delta))
    keep = ii != jj
    assert len(got) == keep.sum()


def test_pairs_edge_cases(ctx):
    sc = common.scenario(3000, 0.4, 0.02)
    ctx.set_cloud_p(sc["P"], 0.02)
    # fewer points than one group, a single point, no pair at that distance
    ctx.set_cloud_q(sc["Q"][:10])
    pt = oport.Port(sc["P"], sc["Q"][:10], 0.02)
    assert np.array_equal(ctx.extract_pairs(0.8, 0, 0.3), pt.extract_pairs(0.8, 0, 0.3))
    ctx.set_cloud_q(sc["Q"][:1])
    assert len(ctx.extract_pairs(0.8, 0, 0.3)) == 0
    ctx.set_cloud_q(sc["Q"])
    assert len(ctx.extract_pairs(50.0, 0, 0.04)) == 0
    assert ctx.count_pairs(50.0, 0.04) == 0
    # slots are independent and round-trip through set_pairs / get_pairs
    a = ctx.extract_pairs(1.0, 0, 0.04, slot=0)
    b = ctx.extract_pairs(0.5, 0, 0.04, slot=1)
    assert np.array_equal(ctx.get_pairs(0, len(a)), a) and np.array_equal(ctx.get_pairs(1, len(b)), b)
    ctx.set_pairs(0, b[::-1])                                # uploaded lists keep the caller's order
    assert np.array_equal(ctx.get_pairs(0, len(b)), b[::-1])


def test_pairs_full_size_properties(ctx):
    """50K-point stage size (BASELINE cfg1) and a 1M-point counting query: symmetry + counts."""
    n, delta = 50_000, 0.01
    sc = common.scenario(n, 0.4, delta, seed=42)
    ctx.set_cloud_p(sc["P"], delta)
    ctx.set_cloud_q(sc["Q"])
    k = ctx.extract_pairs(1.0, 0, 2 * delta, fetch=False)
    assert k == ctx.count_pairs(1.0, 2 * delta)
    p = ctx.get_pairs(0, k)
    assert (p[:, 0] != p[:, 1]).all()
    key = p[:, 0].astype(np.int64) * n + p[:, 1]
    assert (np.diff(key) > 0).all()                          # sorted, no duplicates
    rkey = np.sort(p[:, 1].astype(np.int64) * n + p[:, 0])
    assert np.array_equal(key, rkey)                         # (a,b) present <=> (b,a) present
    d = np.linalg.norm(sc["Q"][p[:, 0]].astype(np.float64) - sc["Q"][p[:, 1]], axis=1)
    assert (np.abs(d - 1.0) <= 2 * delta + 1e-6).all()
    # a random sample of rows against brute force
    Q = sc["Q"]
    for a in np.random.RandomState(0).choice(n, 20, replace=False):
        dd = np.sqrt(((Q - Q[a]) ** 2).astype(np.float32).sum(1, dtype=np.float32))
        # Eigen order x^2 + (y^2 + z^2)
        df = (Q - Q[a]).astype(np.float32)
        dd = np.sqrt(df[:, 0] * df[:, 0] + (df[:, 1] * df[:, 1] + df[:, 2] * df[:, 2]))
        want = np.nonzero(np.abs(dd.astype(np.float64) - np.float32(1.0)) <= np.float32(2 * delta))[0]
        want = want[want != a]
        assert np.array_equal(p[p[:, 0] == a][:, 1], want)
