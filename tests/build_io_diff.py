"""TEST INFRASTRUCTURE: builds tests/cpp/io_diff_main.cc twice -- with the reference's IOManager (sources where they
lie under /root/reference; outputs into tests/_build/, git-ignored) and with the product's -- for tests/test_io_cpu.py."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")
REF_EXE = os.path.join(HERE, "_build", "io_diff_ref")
OUR_EXE = os.path.join(HERE, "_build", "io_diff_ours")


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps if os.path.exists(d))


def build(force=False):
    """returns (ref_exe or None, our_exe or None)"""
    src = os.path.join(HERE, "cpp", "io_diff_main.cc")
    eig = os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen")
    ref_io = os.path.join(REFERENCE_ROOT, "src", "super4pcs", "io", "io.cc")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    os.makedirs(os.path.dirname(REF_EXE), exist_ok=True)
    if os.path.exists(ref_io) and (force or _newer(REF_EXE, [src])):
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-w", "-I", os.path.join(REFERENCE_ROOT, "src"), "-I", eig, src, ref_io,
                               "-o", REF_EXE], env=env)
    our_io = os.path.join(ROOT, "cpp", "io.cc")
    hdrs = [os.path.join(ROOT, "include", "super4pcs", p) for p in ("io/io.h", "shared4pcs.h", "utils/geometry.h")]
    if os.path.exists(os.path.join(eig, "Eigen", "Core")) and (force or _newer(OUR_EXE, [src, our_io] + hdrs)):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-w", "-I", os.path.join(ROOT, "include"), "-I", eig, src, our_io,
                               "-o", OUR_EXE], env=env)
    return (REF_EXE if os.path.exists(REF_EXE) else None, OUR_EXE if os.path.exists(OUR_EXE) else None)
