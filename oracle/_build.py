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
