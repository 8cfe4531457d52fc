"""The reference's MeshLab plugin (demos/MeshlabPlugin/filter_globalregistration, both files UNCHANGED) compiled against the
product's headers with a small MeshLab/Qt stub (tests/stubs/meshlab; neither MeshLab nor Qt is in this image), -std=c++11
like the reference's own build.  On the CPU it runs on the oracle-backed stand-in for libs4g and must end exactly like the
same plugin built against the reference library; tests/test_zzz_meshlab_gpu.py runs it on the real CUDA library."""
import os
import subprocess

import numpy as np
import pytest

from tests import build_meshlab_stub, build_shim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_plugin(exe, tmp_path, args=("70", "0.01", "200"), env=None):
    h = np.load(os.path.join(GOLD, "hippo.npz"))
    for nme, arr in (("a.xyz", h["P"]), ("b.xyz", h["Q"])):
        if not os.path.exists(tmp_path / nme):
            np.savetxt(tmp_path / nme, arr, fmt="%.9g")
    r = subprocess.run([exe, str(tmp_path / "a.xyz"), str(tmp_path / "b.xyz")] + list(args), capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    final = [ln for ln in lines if "Final LCP" in ln]
    rows = [ln for ln in lines if ln.startswith("Tr-row:")]
    assert lines[0].startswith("filter: Global registration") and len(final) == 1 and len(rows) == 4
    return final[0], rows


def test_meshlab_plugin_compiles_and_matches_the_reference_build(s4g_lib, tmp_path):
    from super4pcs_b200 import build_cpp
    build_cpp.build_all()
    exe = build_meshlab_stub.build()
    if exe is None:
        pytest.skip("needs the reference tree + Eigen at build time (or the prebuilt binary)")
    ref = build_meshlab_stub.build_reference()
    shim = build_shim.build()
    ours = run_plugin(exe, tmp_path, env=dict(os.environ, LD_PRELOAD=shim, S4PCS_LANES="2"))
    assert ours[0].endswith("Final LCP = 0.640000")
    g = np.load(os.path.join(GOLD, "hippo_result.npz"))
    M = np.array([r.split()[1:] for r in ours[1]], np.float32)
    assert np.abs(M - g["T_colmajor"].reshape(4, 4).T).max() <= 1e-6
    if ref is not None:
        assert run_plugin(ref, tmp_path) == ours                      # same final log line, same matrix text
        other = ("55", "0.02", "150")
        assert run_plugin(ref, tmp_path, other) == run_plugin(exe, tmp_path, other, env=dict(os.environ, LD_PRELOAD=shim))
