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


@pytest.mark.parametrize("std", ["c++11", "c++14", "c++17", "c++20"])
def test_headers_compile_under_every_standard_callers_use(std):
    """the reference builds with C++11; newer callers exist: the reference's demo main must compile against include/ under each"""
    ref = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")
    main = os.path.join(ref, "demos", "Super4PCS", "super4pcs_test.cc")
    if not os.path.exists(main):
        pytest.skip("needs the reference tree")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(["g++", "-std=" + std, "-fsyntax-only", "-w", "-I", os.path.join(root, "include"),
                        "-I", os.path.join(ref, "3rdparty", "Eigen"), "-I", os.path.join(ref, "demos"), main],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
