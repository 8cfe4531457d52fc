"""TEST INFRASTRUCTURE ONLY: builds tests/_build/libs4g_oracle_shim.so (tests/stubs/s4g_oracle_shim.cc), a
stand-in for libs4g.so answered by the CPU oracle, used through LD_PRELOAD by tests/test_host_logic_cpu.py to run the
host logic of the C++ layer without a GPU.  Never part of the product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build", "libs4g_oracle_shim.so")


def build(force=False):
    from oracle import _build
    port = _build.build_port()
    src = os.path.join(HERE, "stubs", "s4g_oracle_shim.cc")
    deps = [src, port, os.path.join(ROOT, "include", "s4g.h")]
    if force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               src, "-o", OUT, "-L", os.path.dirname(port), "-loracle_port",
                               "-Wl,-rpath," + os.path.dirname(port)], env=env)
    return OUT


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ROOT)
    print(build(force=True))
