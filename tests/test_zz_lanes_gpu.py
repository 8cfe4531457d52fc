"""GPU runs of the host-logic scenarios of tests/test_host_logic_cpu.py against the real CUDA library: speculative
multi-base execution (SURVEY.md 8 row f1, S4PCS_LANES > 1) must be indistinguishable from the sequential RANSAC loop --
same result, same per-iteration visitor reports, same RNG state after early termination -- and from the reference."""
import os

import numpy as np
import pytest

from oracle import _build
from tests.test_host_logic_cpu import ROOT, run_driver

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(_build.build_ref() is None, reason="oracle/_ref (compiled reference) not present")


@pytest.fixture(scope="module")
def built(s4g_lib):
    from super4pcs_b200 import build_cpp
    if build_cpp.build_all()["lib"] is None or _build.build_dropin_harness() is None:
        pytest.skip("C++ layer not available")


@pytest.mark.parametrize("lanes,fused", [(4, 1), (2, 0)])
def test_hippo_with_lanes_matches_golden(built, lanes, fused):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    r = run_driver("hippo", "dropin", lanes=lanes, fused=fused)
    assert np.float32(r["score"]) == g["score"] == np.float32(0.64)
    assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32))


@needs_ref
def test_ransac_trace_with_lanes_matches_reference(built):
    want = run_driver("trace", "reference")
    for lanes in (1, 4, 6):
        assert run_driver("trace", "dropin", lanes=lanes) == want, lanes


@needs_ref
def test_stepwise_termination_and_rng_state_with_lanes_match_reference(built):
    want = run_driver("steps", "reference")
    assert any(row[0] for row in want["log"])
    for lanes in (4, 7):
        assert run_driver("steps", "dropin", lanes=lanes) == want, lanes

