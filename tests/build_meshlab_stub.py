"""Test infrastructure: compiles tests/cpp/meshlab_plugin_main.cc -- our driver around the REFERENCE's MeshLab plugin
(demos/MeshlabPlugin/filter_globalregistration, used unchanged) -- against the product's headers and the MeshLab/Qt stub in
tests/stubs/meshlab, with -std=c++11 like the reference's own build.  Needs the reference tree and Eigen at build time; the
binary travels."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = os.environ.get("S4_REFERENCE_ROOT", "/root/reference")
EXE = os.path.join(HERE, "_build", "meshlab_plugin_test")


def build(force=False):
    plugin = os.path.join(REFERENCE_ROOT, "demos", "MeshlabPlugin", "filter_globalregistration")
    eig = os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen")
    libdir = os.path.join(ROOT, "super4pcs_b200", "lib")
    lib = os.path.join(libdir, "libsuper4pcs_b200.so")
    src = os.path.join(HERE, "cpp", "meshlab_plugin_main.cc")
    stub = os.path.join(HERE, "stubs", "meshlab", "common", "interfaces.h")
    have = os.path.isdir(plugin) and os.path.exists(os.path.join(eig, "Eigen", "Core")) and os.path.exists(lib)
    stale = not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in (src, stub, lib))
    if have and (force or stale):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-w", "-I", os.path.join(HERE, "stubs", "meshlab"), "-I", plugin,
                               "-I", os.path.join(ROOT, "include"), "-I", eig, src, "-o", EXE, "-L", libdir,
                               "-lsuper4pcs_b200", "-ls4g", "-Wl,-rpath,$ORIGIN/../../super4pcs_b200/lib"], env=env)
    return EXE if os.path.exists(EXE) else None


REF_EXE = os.path.join(HERE, "_build", "meshlab_plugin_test_reference")


def build_reference(force=False):
    """the same driver + plugin compiled against the REFERENCE's headers and sources (CPU only): the expected output"""
    plugin = os.path.join(REFERENCE_ROOT, "demos", "MeshlabPlugin", "filter_globalregistration")
    src_root = os.path.join(REFERENCE_ROOT, "src")
    algo = os.path.join(src_root, "super4pcs", "algorithms")
    if os.path.isdir(plugin) and (force or not os.path.exists(REF_EXE)):
        os.makedirs(os.path.dirname(REF_EXE), exist_ok=True)
        env = dict(os.environ)
        env.pop("CXX", None)
        env.pop("CC", None)
        subprocess.check_call(["g++", "-std=c++11", "-O3", "-DNDEBUG", "-w", "-fopenmp", "-DSUPER4PCS_USE_OPENMP", "-DEIGEN_DONT_PARALLELIZE",
                               "-I", os.path.join(HERE, "stubs", "meshlab"), "-I", plugin, "-I", src_root,
                               "-I", os.path.join(REFERENCE_ROOT, "3rdparty", "Eigen"), os.path.join(HERE, "cpp", "meshlab_plugin_main.cc"),
                               os.path.join(algo, "4pcs.cc"), os.path.join(algo, "super4pcs.cc"), os.path.join(algo, "match4pcsBase.cc"),
                               "-o", REF_EXE], env=env)
    return REF_EXE if os.path.exists(REF_EXE) else None


if __name__ == "__main__":
    print(build(force=True))
