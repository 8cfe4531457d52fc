"""GPU restatement of the reference's own test of this path (tests/pair_extraction.cc:239-314, the only runnable
reference test that pins it -- SURVEY.md 8(c)): sphere clouds of 200 / 150 points, delta = 0.1, pair distances
0.3 / 0.5, epsilon = 2 delta; ExtractPairs, sorted, must equal the brute-force set.  Checked through the C ABI and
through MatchSuper4PCS::ExtractPairs of the header-compatible layer (the reference test's TestMatcher pattern)."""
import numpy as np
import pytest

from oracle import _build
from oracle import ref as oref
from tests.test_oracle_golden import _bruteforce_pairs, _sphere_cloud

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reference_pair_extraction_test_on_the_gpu(s4g_lib, seed):
    from super4pcs_b200 import Context
    P, Q = _sphere_cloud(200, seed), _sphere_cloud(150, 100 + seed)
    delta = 0.1
    Pc = P - P.mean(axis=0, dtype=np.float32)
    Qc = Q - Q.mean(axis=0, dtype=np.float32)
    with Context(0) as ctx:
        ctx.set_cloud_p(Pc, delta)
        ctx.set_cloud_q(Qc)
        for d, ang in ((0.3, 0.6), (0.5, 0.4)):
            want = _bruteforce_pairs(Q, d, 2 * delta)
            assert len(want) > 100
            assert np.array_equal(ctx.extract_pairs(d, ang, 2 * delta), want)
    harness = _build.build_dropin_harness()
    if harness is None:
        pytest.skip("C++ layer not available")
    m = oref.RefMatcher(P, Q, oref.make_options(delta=delta, overlap=0.5, sample_size=10 ** 8), libpath=harness)
    for d, ang in ((0.3, 0.6), (0.5, 0.4)):
        assert np.array_equal(m.extract_pairs(d, ang, 2 * delta, 0, 1), _bruteforce_pairs(Q, d, 2 * delta))
    m.close()
