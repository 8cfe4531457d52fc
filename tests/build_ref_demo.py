"""TEST INFRASTRUCTURE: the reference's own demo executable, compiled from the sources where they lie under /root/reference
(g++ on five .cc files, the reference's Release flags; output in tests/_build/, git-ignored), so that
tests/test_host_logic_cpu.py can compare its console output and written files with the product's demo byte for byte."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")
EXE = os.path.join(HERE, "_build", "Super4PCS_reference")


def build(force=False):
    src = os.path.join(REFERENCE_ROOT, "src", "super4pcs")
    main = os.path.join(REFERENCE_ROOT, "demos", "Super4PCS", "super4pcs_test.cc")
    if os.path.exists(main) and (force or not os.path.exists(EXE)):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++11", "-O3", "-DNDEBUG", "-w", "-fopenmp", "-DSUPER4PCS_USE_OPENMP", "-DEIGEN_DONT_PARALLELIZE",
                               "-I", os.path.join(REFERENCE_ROOT, "src"), "-I", os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen"),
                               "-I", os.path.join(REFERENCE_ROOT, "demos"), main,
                               os.path.join(src, "algorithms", "4pcs.cc"), os.path.join(src, "algorithms", "super4pcs.cc"),
                               os.path.join(src, "algorithms", "match4pcsBase.cc"), os.path.join(src, "io", "io.cc"), "-o", EXE], env=env)
    return EXE if os.path.exists(EXE) else None
