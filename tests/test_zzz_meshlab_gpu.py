"""GPU run of the reference's MeshLab plugin on the product (see tests/test_meshlab_plugin.py)."""
import os

import numpy as np
import pytest

from tests import build_meshlab_stub
from tests.test_meshlab_plugin import GOLD, run_plugin

pytestmark = pytest.mark.gpu


def test_meshlab_plugin_registers_hippo_like_the_reference(s4g_lib, tmp_path):
    exe = build_meshlab_stub.build()
    if exe is None:
        pytest.skip("meshlab plugin test binary not available")
    final, rows = run_plugin(exe, tmp_path)
    assert final.endswith("Final LCP = 0.640000")
    g = np.load(os.path.join(GOLD, "hippo_result.npz"))
    M = np.array([r.split()[1:] for r in rows], np.float32)
    assert np.abs(M - g["T_colmajor"].reshape(4, 4).T).max() <= 1e-6
