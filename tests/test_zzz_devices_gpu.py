"""GPU runs of S4PCS_DEVICES (SURVEY.md 8 row e inside the C++ layer): the candidates of every base sharded over several
device contexts of ONE process -- two contexts of GPU 0 on a one-GPU box ("0,0"), two GPUs when the box has them -- must be
indistinguishable from the single-context run and from the reference (tests/test_host_logic_cpu.py has the CPU twin)."""
import os

import numpy as np
import pytest

from oracle import _build
from tests.test_host_logic_cpu import ROOT, run_driver

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(_build.build_ref() is None, reason="oracle/_ref (compiled reference) not present")


def _device_specs():
    import torch
    return ["0,0"] + (["2", "all"] if torch.cuda.device_count() >= 2 else [])


def test_device_count_is_the_driver_s(s4g_lib):
    import torch
    from super4pcs_b200 import s4g
    assert s4g.device_count() == torch.cuda.device_count() >= 1


@pytest.fixture(scope="module")
def built(s4g_lib):
    from super4pcs_b200 import build_cpp
    if build_cpp.build_all()["lib"] is None or _build.build_dropin_harness() is None:
        pytest.skip("C++ layer not available")


@pytest.mark.parametrize("lanes,fused", [(1, 1), (2, 1), (1, 0)])
def test_hippo_sharded_over_device_contexts_matches_golden(built, lanes, fused):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    for spec in _device_specs():
        r = run_driver("hippo", "dropin", lanes=lanes, fused=fused, extra_env={"S4PCS_DEVICES": spec}, timeout=300)
        assert np.float32(r["score"]) == g["score"] == np.float32(0.64), spec
        assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32)), spec


@needs_ref
def test_sharded_trace_and_sweep_match_reference(built):
    for which in ("trace", "sweep2"):
        want = run_driver(which, "reference")
        for spec in _device_specs():
            assert run_driver(which, "dropin", extra_env={"S4PCS_DEVICES": spec}, timeout=300) == want, (which, spec)
    want = run_driver("steps", "reference")
    assert run_driver("steps", "dropin", lanes=3, extra_env={"S4PCS_DEVICES": "0,0,0"}, timeout=300) == want


def _multi_gpu():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.skipif("not _multi_gpu()", reason="needs two GPUs (NCCL wants one rank per device)")
def test_library_side_reduction_over_nccl_matches_golden_and_reference(built):
    """S4PCS_NCCL=1: the contexts share a communicator (ncclCommInitAll) and libs4g reduces key + record on the devices"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    for spec, fused in (("2", 1), ("all", 1), ("2", 0)):
        r = run_driver("hippo", "dropin", fused=fused, extra_env={"S4PCS_DEVICES": spec, "S4PCS_NCCL": "1"}, timeout=300)
        assert np.float32(r["score"]) == g["score"] == np.float32(0.64), spec
        assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32)), spec
    if _build.build_ref() is not None:
        for which in ("trace", "sweep2"):
            want = run_driver(which, "reference")
            got = run_driver(which, "dropin", extra_env={"S4PCS_DEVICES": "all", "S4PCS_NCCL": "1"}, timeout=300)
            assert got == want, which


def test_library_side_reduction_refuses_two_contexts_on_one_gpu(built):
    """NCCL wants one rank per device: "0,0" with S4PCS_NCCL=1 is an error, not a silent host merge"""
    with pytest.raises(AssertionError):
        run_driver("hippo", "dropin", extra_env={"S4PCS_DEVICES": "0,0", "S4PCS_NCCL": "1"}, timeout=300)


def test_a_missing_device_is_an_error_not_a_fallback(built):
    """ordinal 63 does not exist on any box: the run must fail loudly (std::runtime_error -> non-zero exit)"""
    with pytest.raises(AssertionError):
        run_driver("hippo", "dropin", extra_env={"S4PCS_DEVICES": "0,63"}, timeout=300)
