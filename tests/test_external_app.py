"""CPU test: the reference's packaging test (tests/externalAppTest/main.cpp), compiled UNCHANGED against
the product's headers and libraries, links and runs.  It registers two EMPTY clouds, which returns the
kLargeNumber sentinel before any device work (reference match4pcsBase.hpp:69-70), so no GPU is needed."""
import os
import subprocess

import pytest


def test_reference_external_app_links_and_runs(tmp_path, s4g_lib):
    from super4pcs_b200 import build_cpp
    out = build_cpp.build_all()
    exe = out.get("external_app_test")
    if not exe:
        pytest.skip("externalAppTest not built (needs the reference source + Eigen at build time)")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Score: 1e+09" in r.stdout                      # kLargeNumber
    assert os.path.exists(tmp_path / "output.map")         # IOManager::WriteMatrix ran
    assert open(tmp_path / "output.map").read().startswith("VERSION\t=\t1\nMATRIX\t=\n")
