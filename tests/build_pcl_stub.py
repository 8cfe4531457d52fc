"""Test infrastructure: compiles tests/cpp/pcl_wrapper_main.cc -- our driver around the REFERENCE's
pcl::Super4PCS wrapper (demos/PCLWrapper, included unchanged) -- against the product's headers and the
PCL stub in tests/stubs/.  Needs the reference tree and Eigen at build time; the binary travels."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")
EXE = os.path.join(HERE, "_build", "pcl_wrapper_test")


def build(force=False):
    wrapper = os.path.join(REFERENCE_ROOT, "demos", "PCLWrapper")
    eig = os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen")
    libdir = os.path.join(ROOT, "super4pcs_b200", "lib")
    src = os.path.join(HERE, "cpp", "pcl_wrapper_main.cc")
    have = os.path.isdir(wrapper) and os.path.exists(os.path.join(eig, "Eigen", "Core")) and \
        os.path.exists(os.path.join(libdir, "libsuper4pcs_b200.so"))
    stale = not os.path.exists(EXE) or os.path.getmtime(src) > os.path.getmtime(EXE) or \
        os.path.getmtime(os.path.join(libdir, "libsuper4pcs_b200.so")) > os.path.getmtime(EXE)
    if have and (force or stale):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-w", "-I", os.path.join(HERE, "stubs"), "-I", wrapper,
                               "-I", os.path.join(ROOT, "include"), "-I", eig, src, "-o", EXE, "-L", libdir,
                               "-lsuper4pcs_b200", "-ls4g", "-Wl,-rpath,$ORIGIN/../../super4pcs_b200/lib"], env=env)
    return EXE if os.path.exists(EXE) else None


if __name__ == "__main__":
    print(build(force=True))
