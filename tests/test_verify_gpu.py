"""GPU parity of a6 / a7 / a8 (rigid fit, TryCongruentSet, Verify) through the C ABI against the
CPU oracle: oracle/_ref (the unmodified reference) when it travelled, and the port always."""
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


def _setup(ctx, sc):
    ctx.set_cloud_p(sc["P"], sc["delta"])
    ctx.set_cloud_q(sc["Q"])


@pytest.mark.parametrize("n,delta", [(3000, 0.02), (20000, 0.01)])
def test_verify_counts_match_reference_and_port(ctx, n, delta):
    sc = common.scenario(n, 0.4, delta)
    _setup(ctx, sc)
    T = common.candidates_colmajor(sc, 96)
    got = ctx.verify(T)
    pt = oport.Port(sc["P"], sc["Q"], delta)
    lcp_p, good_p, _ = pt.verify_batch(T, 0.0, nthreads=oport.num_threads())
    assert np.array_equal(got, good_p)                       # integer counts: bit-exact
    if oref.available():
        opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.4)
        m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], opt)
        lcp_r, _ = m.verify_batch(T, 0.0, nthreads=oref.num_threads())
        assert np.array_equal((got.astype(np.float32) / np.float32(n)), lcp_r)


def test_verify_edge_cases(ctx):
    sc = common.scenario(3000, 0.4, 0.02)
    _setup(ctx, sc)
    assert len(ctx.verify(np.zeros((0, 16), np.float32))) == 0
    # far-away transform: nothing within delta
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = 1000.0
    assert ctx.verify(T.T.reshape(1, 16))[0] == 0
    # NaN transform must not crash and counts nothing
    Tn = np.full((1, 16), np.nan, np.float32)
    assert ctx.verify(Tn)[0] == 0
    # Q == P with identity: every point is its own neighbour
    ctx.set_cloud_q(sc["P"])
    assert ctx.verify(np.eye(4, dtype=np.float32).reshape(1, 16))[0] == len(sc["P"])


def test_verify_ragged_sizes(ctx):
    # sizes that are not multiples of the tile / chunk, single point clouds
    sc = common.scenario(3000, 0.4, 0.02)
    P, Q = sc["P"][:1001], sc["Q"][:777]
    ctx.set_cloud_p(P, 0.05)
    ctx.set_cloud_q(Q)
    T = common.candidates_colmajor(sc, 37)
    pt = oport.Port(P, Q, 0.05)
    assert np.array_equal(ctx.verify(T), pt.verify_bruteforce(T))
    ctx.set_cloud_p(P[:1], 0.05)
    ctx.set_cloud_q(Q[:1])
    pt = oport.Port(P[:1], Q[:1], 0.05)
    assert np.array_equal(ctx.verify(T), pt.verify_bruteforce(T))


def test_rigid_bit_exact(ctx):
    sc = common.scenario(3000, 0.4, 0.02)
    _setup(ctx, sc)
    rng = np.random.RandomState(3)
    base = rng.randint(0, len(sc["P"]), 4).astype(np.int32)
    quads = np.concatenate([common.random_quads(len(sc["Q"]), 20000, 1),
                            common.congruent_like_quads(sc, base, 2000, 2),
                            np.array([[5, 5, 9, 1], [5, 9, 5, 1], [7, 8, 8, 2]], np.int32)])  # degenerate frames
    T, rms, ok = ctx.rigid_batch(sc["P"][base], quads)
    pt = oport.Port(sc["P"], sc["Q"], sc["delta"])
    Tp, rp, okp = pt.rigid_batch(base, quads)
    assert np.array_equal(ok, okp)
    assert np.array_equal(common.bits(rms), common.bits(rp))
    assert np.array_equal(common.bits(T), common.bits(Tp))
    if oref.available():
        opt = oref.make_options(delta=sc["delta"], sample_size=10 ** 8, overlap=0.4)
        m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], opt)
        Tr, rr, okr = m.rigid_batch(base, quads)
        assert np.array_equal(ok, okr)
        assert np.array_equal(common.bits(rms), common.bits(rr))
        assert np.array_equal(common.bits(T), common.bits(Tr))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_try_congruent_set_matches_oracle(ctx, world):
    delta = 0.05
    sc = common.scenario(3000, 0.4, delta)
    _setup(ctx, sc)
    rng = np.random.RandomState(11)
    # a base inside the overlap so that ground-truth-like quads exist
    gt = sc["raw"]["gt"]
    zP = (sc["P"] + sc["cp"])[:, 2]
    base = rng.choice(np.nonzero(np.abs(zP) < 0.15)[0], 4, replace=False).astype(np.int32)
    quads = common.congruent_like_quads(sc, base, 3000, 5)
    pt = oport.Port(sc["P"], sc["Q"], delta)
    want = pt.try_congruent_set(base, quads, best_lcp_in=0.0)
    assert want["n_gate"] > 20 and want["best_index"] >= 0
    shards = [ctx.try_congruent_set(sc["P"][base], quads, 2 * delta, shard_rank=r, shard_world=world)
              for r in range(world)]
    assert sum(s["n_gate_pass"] for s in shards) == want["n_gate"]
    key = max(s["key"] for s in shards)                      # what the allreduce(MAX) computes
    win = [s for s in shards if s["key"] == key][0]
    assert win["best_index"] == want["best_index"]
    lcp = np.float32(win["best_count"]) / np.float32(win["n_q"])
    assert lcp == np.float32(want["best_lcp"])
    assert np.array_equal(common.bits(win["T"]), common.bits(want["T"]))
    if oref.available() and world == 1:
        opt = oref.make_options(delta=delta, sample_size=10 ** 8, overlap=0.4)
        m = oref.RefMatcher(sc["raw"]["P"], sc["raw"]["Q"], opt)
        m.set_best_lcp(0.0)
        r = m.try_congruent_set(base, quads)
        assert r["n_gate"] == win["n_gate_pass"]
        assert np.float32(r["best_lcp"]) == lcp
        assert np.array_equal(common.bits(r["T"]), common.bits(win["T"]))
        assert np.array_equal(r["congruent"], quads[win["best_index"]])


def test_try_congruent_set_empty_and_none_pass(ctx):
    sc = common.scenario(3000, 0.4, 0.02)
    _setup(ctx, sc)
    base = np.array([0, 1, 2, 3], np.int32)
    r = ctx.try_congruent_set(sc["P"][base], np.zeros((0, 4), np.int32), 0.04)
    assert r["best_index"] == -1 and r["key"] == 0 and r["n_gate_pass"] == 0
    # degenerate quads never pass the gate (rms = 1e9)
    r = ctx.try_congruent_set(sc["P"][base], np.array([[4, 4, 4, 4]] * 10, np.int32), 0.04)
    assert r["best_index"] == -1 and r["n_gate_pass"] == 0


def test_verify_full_size_properties(ctx):
    """1M x 1M (BASELINE cfg2 size): size-independent properties instead of the CPU oracle"""
    n, delta = 1_000_000, 0.003
    sc = common.scenario(n, 0.3, delta, seed=42)
    _setup(ctx, sc)
    T = common.candidates_colmajor(sc, 24, n_near=8)
    c1 = ctx.verify(T)
    assert np.array_equal(c1, ctx.verify(T))                 # deterministic
    assert np.array_equal(c1[::-1], ctx.verify(T[::-1]))      # independent of batch order
    assert np.array_equal(c1[:5], ctx.verify(T[:5]))          # independent of batch size
    assert (c1[:8] > 0.05 * n).all() and (c1[:8] <= n).all()  # near-GT candidates see the overlap
    assert (c1[8:] < c1[:8].min()).all()                      # random motions do not
    # additivity over a split of Q: counts(Q) = counts(Q[:m]) + counts(Q[m:])
    m = 400_003
    ctx.set_cloud_q(sc["Q"][:m])
    a = ctx.verify(T)
    ctx.set_cloud_q(sc["Q"][m:])
    b = ctx.verify(T)
    assert np.array_equal(a + b, c1)
    # sampled subset against the port (exact): 20k queries of the 1M cloud
    sub = np.random.RandomState(0).choice(n, 20000, replace=False)
    ctx.set_cloud_q(sc["Q"][sub])
    pt = oport.Port(sc["P"], sc["Q"][sub], delta)
    _, good, _ = pt.verify_batch(T, 0.0, nthreads=oport.num_threads())
    assert np.array_equal(ctx.verify(T), good)


def test_verify_dense_cloud_culling_paths_match_port(ctx):
    """200K x 200K, small delta: compact Morton tiles -> the tile-level cull, the occupancy bitmap and
    the compaction queue are all exercised; counts must still equal the oracle's exactly."""
    n, delta = 200_000, 0.004
    sc = common.scenario(n, 0.3, delta, seed=17)
    _setup(ctx, sc)
    T = np.concatenate([common.candidates_colmajor(sc, 40, seed=3, n_near=10),
                        # pure small translations of the identity-on-GT: partially overlapping tiles
                        common.candidates_colmajor(sc, 8, seed=4, n_near=8)])
    pt = oport.Port(sc["P"], sc["Q"], delta)
    _, good, _ = pt.verify_batch(T, 0.0, nthreads=oport.num_threads())
    got = ctx.verify(T)
    assert np.array_equal(got, good)
    st = ctx.verify_probe_stats(T)
    assert st["tile_pairs_culled"] > 0                     # the cull really fires on this workload
    assert np.array_equal(ctx.verify(T[:1]), good[:1])     # chunk with a single candidate
    assert np.array_equal(ctx.verify(T[:17]), good[:17])   # chunk boundary (16 + 1)


def test_verify_queue_overflow_rounds(ctx):
    """Q == P.  (a) 16 (near-)identity candidates: every pair is decided by the delta-field's CERTAIN bit (a query that
    coincides with a P point) or by the exact test; all n points must be found.  (b) 16 shifts of 0.9 delta in random
    directions: most (query, candidate) pairs are MAYBE-but-not-CERTAIN, i.e. > 1024 queued pairs per 128-query tile,
    which overflows the CTA's 2048-entry queue and forces the leftover rounds of the kernel.  (c) half-delta steps
    along one axis.  Counts must equal the oracle's exactly in every case."""
    n, delta = 100_000, 0.004
    sc = common.scenario(n, 0.3, delta, seed=23)
    P = sc["P"]
    ctx.set_cloud_p(P, delta)
    ctx.set_cloud_q(P)
    rng = np.random.RandomState(1)
    T = np.tile(np.eye(4, dtype=np.float32), (16, 1, 1))
    for k in range(1, 16):
        v = rng.standard_normal(3)
        T[k, :3, 3] = (v / np.linalg.norm(v) * delta * 0.45 * rng.random_sample()).astype(np.float32)
    Tc = np.ascontiguousarray(T.transpose(0, 2, 1)).reshape(16, 16)
    got = ctx.verify(Tc)
    assert (got == n).all()                                  # every point finds (at least) its own source
    pt = oport.Port(P, P, delta)
    _, good, _ = pt.verify_batch(Tc[:3], 0.0, nthreads=oport.num_threads())
    assert np.array_equal(got[:3], good)
    # (b) 0.9 delta shifts: the queue overflows
    T1 = np.tile(np.eye(4, dtype=np.float32), (16, 1, 1))
    for k in range(16):
        v = rng.standard_normal(3)
        T1[k, :3, 3] = (v / np.linalg.norm(v) * delta * 0.9).astype(np.float32)
    T1c = np.ascontiguousarray(T1.transpose(0, 2, 1)).reshape(16, 16)
    _, good1, _ = pt.verify_batch(T1c, 0.0, nthreads=oport.num_threads())
    got1 = ctx.verify(T1c)
    assert np.array_equal(got1, good1)
    st = ctx.verify_probe_stats(T1c)
    assert st["ranges_read"] > 800 * (n // 128)              # hundreds of exact tests per tile: the flush / leftover rounds run
    # (c) half-delta shifts along one axis: not everything matches any more; still exact
    T2 = np.tile(np.eye(4, dtype=np.float32), (16, 1, 1))
    T2[:, 0, 3] = np.linspace(0.5, 3.0, 16, dtype=np.float32) * delta
    T2c = np.ascontiguousarray(T2.transpose(0, 2, 1)).reshape(16, 16)
    _, good2, _ = pt.verify_batch(T2c, 0.0, nthreads=oport.num_threads())
    assert np.array_equal(ctx.verify(T2c), good2)


def test_verify_widened_cells(ctx):
    """delta tiny against the extent: more than 2000 cells per axis, so the cell edge is widened beyond 2.02 delta (voxels of
    the delta-field are then larger than delta/2: CERTAIN bits become rare, MAYBE bits common, counts must not change)."""
    n, delta = 20_000, 0.0004
    sc = common.scenario(n, 0.4, delta, seed=31)
    _setup(ctx, sc)
    gs = ctx.grid_stats()
    assert gs["cell_edge"] > 2.02 * delta * 1.05            # widened beyond 2.02 delta
    T = common.candidates_colmajor(sc, 24, seed=2, n_near=12)
    pt = oport.Port(sc["P"], sc["Q"], delta)
    _, good, _ = pt.verify_batch(T, 0.0, nthreads=oport.num_threads())
    assert np.array_equal(ctx.verify(T), good)
    # identity on P == Q still finds every point
    ctx.set_cloud_q(sc["P"])
    assert ctx.verify(np.eye(4, dtype=np.float32).reshape(1, 16))[0] == n


def test_verify_non_rigid_and_huge_transforms(ctx):
    """s4g_verify is documented for ANY 4x4 (round-1 ADVICE: the tile cull assumed an isometry): scaled, sheared and
    far-translated transforms take the robust path of the delta-field / the scaled cull radius and must still give the
    oracle's counts"""
    n, delta = 20_000, 0.01
    sc = common.scenario(n, 0.4, delta, seed=5)
    _setup(ctx, sc)
    base = common.candidates_colmajor(sc, 12, seed=9, n_near=12).reshape(-1, 4, 4).transpose(0, 2, 1).copy()   # row-major 4x4, near GT
    rng = np.random.RandomState(3)
    Ts = []
    for k, M in enumerate(base):
        A = np.eye(4, dtype=np.float64)
        if k % 4 == 0:
            A[:3, :3] *= 1.0 + 0.2 * rng.random_sample()                   # uniform scale > 1
        elif k % 4 == 1:
            A[0, 1] = 0.3 * rng.random_sample()                            # shear
        elif k % 4 == 2:
            A[:3, :3] = np.diag([1.3, 0.8, 1.1])                           # anisotropic scale
        else:
            A[:3, :3] *= 1e4                                               # huge coefficients: nothing can match, must not crash
        Ts.append((M.astype(np.float64) @ A).astype(np.float32))
    T = np.ascontiguousarray(np.stack(Ts).transpose(0, 2, 1)).reshape(-1, 16)
    pt = oport.Port(sc["P"], sc["Q"], delta)
    _, good, _ = pt.verify_batch(T, 0.0, nthreads=oport.num_threads())
    assert np.array_equal(ctx.verify(T), good)
    assert good.max() > 0                                                 # some scaled / sheared near-GT candidate still hits something
