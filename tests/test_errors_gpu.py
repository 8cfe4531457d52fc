"""Error behaviour of the C ABI on a GPU box: status codes + messages instead of crashes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_call_order_and_argument_errors(s4g_lib):
    from super4pcs_b200 import Context, S4GError
    with Context(0) as ctx:
        T = np.eye(4, dtype=np.float32).reshape(1, 16)
        with pytest.raises(S4GError, match="set_cloud"):
            ctx.verify(T)                                   # S4G_ERR_STATE: no clouds yet
        with pytest.raises(S4GError):
            ctx.extract_pairs(1.0, 0.0, 0.02)               # no Q cloud
        P = np.random.RandomState(0).rand(100, 3).astype(np.float32)
        with pytest.raises(S4GError, match="delta"):
            ctx.set_cloud_p(P, 0.0)                         # S4G_ERR_ARG
        bad = P.copy()
        bad[3, 1] = np.nan
        with pytest.raises(S4GError, match="NaN"):
            ctx.set_cloud_p(bad, 0.05)
        ctx.set_cloud_p(P, 0.05)
        ctx.set_cloud_q(P)
        with pytest.raises(S4GError, match="slot"):
            ctx.extract_pairs(0.5, 0.0, 0.1, slot=2)
        with pytest.raises(S4GError, match="range"):
            ctx.set_pairs(0, np.array([[0, 100]], np.int32))   # index >= n
        with pytest.raises(S4GError):
            ctx.extract_pairs(0.5, 0.0, -1.0)               # epsilon <= 0
        # quads referencing points outside sampled_Q are ignored by the rigid fit, never dereferenced
        r = ctx.try_congruent_set(P[:4], np.array([[0, 1, 2, 1000]], np.int32), 0.1)
        assert r["n_gate_pass"] == 0
        # the context is still usable after the errors
        assert ctx.verify(T)[0] == 100


def test_create_on_missing_device_fails(s4g_lib):
    from super4pcs_b200 import S4GError, s4g
    with pytest.raises(S4GError):
        s4g.Context(99)
