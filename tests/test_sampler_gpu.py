"""GPU parity of f2 (Sampling::UniformDistSampler, reference sampling.h:59-121): first point of
every voxel, in input order -- against a plain host restatement and the compiled reference."""
import numpy as np
import pytest

from oracle import ref as oref
from tests import common

pytestmark = pytest.mark.gpu


def host_voxel_sample(X, voxel):
    scale = np.float32(1.0) / np.float32(voxel)
    c = np.floor(X.astype(np.float32) * scale).astype(np.int64)
    c -= c.min(0)
    key = c[:, 0] << 42 | c[:, 1] << 21 | c[:, 2]
    _, first = np.unique(key, return_index=True)
    return np.sort(first).astype(np.int32)


@pytest.mark.parametrize("n,voxel", [(5000, 0.05), (200_000, 0.01), (1_000_003, 0.004)])
def test_voxel_sampler_matches_host(s4g_lib, n, voxel):
    from super4pcs_b200 import Context
    sc = common.scenario(min(n, 200_000), 0.4, 0.01, seed=8)
    X = sc["raw"]["P"]
    if n > len(X):
        X = np.concatenate([X + np.float32(k * 0.001) for k in range(n // len(X) + 1)])[:n]
    with Context(0) as ctx:
        got = ctx.voxel_sample(X, voxel)
    assert np.array_equal(got, host_voxel_sample(X, voxel))


def test_voxel_sampler_georeferenced_coordinates(s4g_lib):
    """coordinates around 1e5 with a 0.05 voxel: absolute voxel coordinates ~2e6 exceed 2^20, the span does not (ADVICE
    round 1: the key is relative to the voxel of the bounding-box minimum); non-finite input is an argument error"""
    from super4pcs_b200 import Context
    sc = common.scenario(20000, 0.4, 0.01, seed=12)
    X = (sc["raw"]["P"] * np.float32(20.0) + np.array([1.2e5, -3.4e5, 7.7e4], np.float32)).astype(np.float32)
    with Context(0) as ctx:
        got = ctx.voxel_sample(X, 0.05)
        assert np.array_equal(got, host_voxel_sample(X, 0.05))
        Y = X.copy()
        Y[0, 1] = np.inf
        with pytest.raises(Exception):
            ctx.voxel_sample(Y, 0.05)


def test_voxel_sampler_matches_reference_sampler(s4g_lib):
    if not oref.available():
        pytest.skip("oracle/_ref not present")
    from super4pcs_b200 import Context
    sc = common.scenario(20000, 0.4, 0.02, seed=9)
    raw = sc["raw"]["P"]
    # the reference samples P when |P| > sample_size (match4pcsBase.hpp:112-119); sampled_P is then centred
    opt = oref.make_options(delta=0.02, overlap=0.4, sample_size=100)
    m = oref.RefMatcher(raw, sc["raw"]["Q"], opt, identity_sampler=False)
    P, _, _ = m.sampled_p()
    with Context(0) as ctx:
        keep = ctx.voxel_sample(raw, 0.02)
    assert len(keep) == m.nP
    Pc = raw[keep]
    c = np.cumsum(Pc, axis=0, dtype=np.float32)[-1] / np.float32(len(Pc))
    assert np.array_equal((Pc - c.astype(np.float32)).astype(np.float32).view(np.uint32), P.view(np.uint32))
