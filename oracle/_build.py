"""TEST INFRASTRUCTURE ONLY: builds oracle/liboracle_port.so and, when /root/reference is
present, oracle/_ref/liboracle_ref.so (see oracle/Makefile)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "liboracle_ref.so")
REFERENCE_ROOT = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_port(force=False):
    if force or _stale(PORT_SO, [os.path.join(HERE, "port.cc")]):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def build_ref(force=False):
    """Returns the path of the reference oracle, or None when it cannot be built here
    (no /root/reference on the GPU box) and no prebuilt copy travelled with the repo."""
    have_ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "super4pcs"))
    if have_ref and (force or _stale(REF_SO, [os.path.join(HERE, "ref_harness.cc")])):
        subprocess.check_call(["make", "-C", HERE, "ref", "REF=" + REFERENCE_ROOT],
                              stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


DROPIN_SO = os.path.join(HERE, "_dropin", "libb200_harness.so")


def build_dropin_harness(force=False):
    """The SAME harness source (oracle/ref_harness.cc, written against the reference's headers)
    compiled unchanged against the PRODUCT's header-compatible layer (include/super4pcs/ +
    libsuper4pcs_b200.so): the drop-in proof tests/test_dropin_gpu.py drives.  Test infrastructure:
    the dependency points from here to the product, never the other way.  Needs Eigen (host-side API
    dependency); without it the prebuilt copy that travelled with the repo is used."""
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "super4pcs_b200", "lib")
    prod = os.path.join(libdir, "libsuper4pcs_b200.so")
    eig = None
    for c in (os.environ.get("S4_EIGEN_ROOT"), os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen"), "/usr/include/eigen3"):
        if c and os.path.exists(os.path.join(c, "Eigen", "Core")):
            eig = c
            break
    src = os.path.join(HERE, "ref_harness.cc")
    if eig and os.path.exists(prod) and (force or _stale(DROPIN_SO, [src, prod])):
        os.makedirs(os.path.dirname(DROPIN_SO), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++14", "-O3", "-DNDEBUG", "-fPIC", "-w", "-fopenmp", "-DSUPER4PCS_USE_OPENMP",
                               "-shared", "-I", os.path.join(root, "include"), "-I", eig, src, "-o", DROPIN_SO,
                               "-L", libdir, "-lsuper4pcs_b200", "-ls4g", "-Wl,-rpath,$ORIGIN/../../super4pcs_b200/lib"], env=env)
    return DROPIN_SO if os.path.exists(DROPIN_SO) else None
