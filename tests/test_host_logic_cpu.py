"""CPU tests of the HOST logic of the header-compatible C++ layer (cpp/*.cc, include/super4pcs/**): sampling,
centring, RNG order, base selection, RANSAC loop, visitor protocol, global transform and the speculative multi-base
execution of SURVEY.md 8 row f1.  The device stages are answered by the CPU oracle through an LD_PRELOADed stand-in
for libs4g.so (tests/stubs/s4g_oracle_shim.cc, test infrastructure), so everything ABOVE the C ABI is the product's
own code.  The same scenarios run against the real CUDA library in tests/test_zz_lanes_gpu.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import _build
from tests import build_shim, common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "host_logic_driver.py")


def run_driver(which, target, lanes=1, fused=1, preload=None, timeout=900, stats=None, extra_env=None):
    env = dict(os.environ, S4PCS_LANES=str(lanes), S4PCS_FUSED=str(fused), S4G_SHIM_STATS="1")
    if lanes > 1:
        env["S4PCS_BATCH"] = "1"      # a test that asks for lanes means the lanes: the batched pass (default on) would take over
    env.update(extra_env or {})
    if preload:
        env["LD_PRELOAD"] = preload
    r = subprocess.run([sys.executable, DRIVER, which, target], env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    if stats is not None:
        for ln in r.stderr.splitlines():
            if ln.startswith("SHIM "):
                stats.update({k: int(v) for k, v in (kv.split("=") for kv in ln.split()[1:])})
    return out


@pytest.fixture(scope="module")
def shim(s4g_lib):
    from super4pcs_b200 import build_cpp
    if build_cpp.build_all()["lib"] is None or _build.build_dropin_harness() is None:
        pytest.skip("C++ layer not buildable here (no Eigen)")
    return build_shim.build()


needs_ref = pytest.mark.skipif(_build.build_ref() is None, reason="oracle/_ref (compiled reference) not present")


@pytest.mark.parametrize("lanes,fused", [(1, 1), (4, 1), (3, 0)])
def test_hippo_through_cpp_layer_on_oracle_shim_matches_golden(shim, lanes, fused):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    r = run_driver("hippo", "dropin", lanes=lanes, fused=fused, preload=shim)
    assert np.float32(r["score"]) == g["score"] == np.float32(0.64)
    assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32))


@needs_ref
def test_ransac_trace_identical_for_any_lane_count(shim):
    """per-iteration visitor reports (fraction, best LCP, global transform) of the whole run"""
    want = run_driver("trace", "reference")
    for lanes in (1, 4, 7):
        st = {}
        assert run_driver("trace", "dropin", lanes=lanes, preload=shim, stats=st) == want, lanes
        # one device context per lane, and the lanes' stage calls really overlapped
        assert st["contexts"] == lanes and (lanes == 1 or st["max_inflight"] > 1), st


@needs_ref
def test_stepwise_termination_and_rng_state_identical_for_any_lane_count(shim):
    """Perform_N_steps in pieces, termination inside a speculative batch, then the NEXT base selected by hand:
    the RNG must be where the sequential loop leaves it (bases selected ahead but not tried are rolled back)"""
    want = run_driver("steps", "reference")
    assert any(row[0] for row in want["log"])               # the scenario does terminate
    for lanes, fused in ((1, 1), (4, 1), (7, 1), (3, 0)):   # 80 bases are tried in the last call: mid-batch for 7 and 3
        assert run_driver("steps", "dropin", lanes=lanes, fused=fused, preload=shim) == want, (lanes, fused)


@needs_ref
@pytest.mark.parametrize("which,lanes", [("synth1", 1), ("synth2n", 4), ("whole", 2), ("trials", 1)])
def test_whole_pipeline_scenarios_match_reference(shim, which, lanes):
    """synthetic pairs with / without normals (voxel sampler, shuffle, filters), whole-cloud and demo-default
    corner cases, empty input, and init() over a sweep of overlaps: C++ layer on the stand-in == the reference"""
    want = run_driver(which, "reference")
    assert run_driver(which, "dropin", lanes=lanes, preload=shim) == want
    if which.startswith("synth"):
        assert want["score"] > 0.2                          # and it is a real registration


@needs_ref
@pytest.mark.parametrize("seed,lanes,fused", [(2, 3, 1), (3, 1, 1), (5, 4, 0)])
def test_randomised_pipeline_sweep_matches_reference(shim, seed, lanes, fused):
    """6 random configurations per seed (cloud size, overlap, noise, outliers, delta, sample size, RNG seed, normal /
    translation filters, terminate threshold): score, matrix bits and the transformed cloud equal the reference's"""
    want = run_driver("sweep%d" % seed, "reference")
    assert run_driver("sweep%d" % seed, "dropin", lanes=lanes, fused=fused, preload=shim) == want


def test_equal_count_ties_follow_the_candidate_order(shim):
    """When candidates with different transforms tie for the best inlier count the reference keeps the first in ITS
    pair-emission order.  The product (default since round 2) replays the reference's traversal on
    the host (cpp/pair_order.cc) and resolves the tie like the reference, bit for bit -- fused pass, staged pass and with
    bases tried ahead (lanes).  S4PCS_EXACT_ORDER=0 turns the replay off: first candidate in sorted order -- same score,
    another matrix.  Cross-check: a stand-in that hands the pairs over in the reference's order (oracle/port.cc) needs
    no replay."""
    same = {"rows": [[True, True]] * 4}
    assert run_driver("ties", "reference") == same                                   # the fixture is the reference's output
    assert run_driver("ties", "dropin", preload=shim) == same                        # no environment variable needed
    off = run_driver("ties", "dropin", preload=shim, extra_env={"S4PCS_EXACT_ORDER": "0"})
    assert all(score_equal for score_equal, _ in off["rows"])
    assert not any(matrix_equal for _, matrix_equal in off["rows"])
    for lanes, fused in ((1, 1), (1, 0), (4, 1), (3, 0)):
        assert run_driver("ties", "dropin", lanes=lanes, fused=fused, preload=shim) == same
    assert run_driver("ties", "dropin", preload=shim, extra_env={"S4G_SHIM_REFERENCE_ORDER": "1", "S4PCS_EXACT_ORDER": "0"}) == same


@needs_ref
@pytest.mark.parametrize("which,lanes,fused", [("sweep2", 3, 1), ("trace", 4, 1), ("steps", 7, 1), ("steps", 2, 0), ("synth2n", 2, 1)])
def test_exact_order_mode_matches_reference_on_the_other_scenarios(shim, which, lanes, fused):
    want = run_driver(which, "reference")
    assert run_driver(which, "dropin", lanes=lanes, fused=fused, preload=shim, extra_env={"S4PCS_EXACT_ORDER": "1"}) == want


@needs_ref
@pytest.mark.parametrize("lanes,fused", [(1, 1), (3, 1), (1, 0)])
def test_initial_lcp_above_the_terminate_threshold_keeps_drawing_bases(shim, lanes, fused):
    """with best_LCP_ already above the threshold a base without pairs / congruent quads must return false
    (reference hpp:335-347) so that the loop goes on: return values, progress reports and RNG state equal the reference's"""
    want = run_driver("prealigned", "reference")
    assert want["log"][0] > 0.04 and not any(row[0] for row in want["log"][1:])     # the scenario really is the advisor's case
    assert run_driver("prealigned", "dropin", lanes=lanes, fused=fused, preload=shim) == want


@needs_ref
@pytest.mark.parametrize("which,batch,lanes", [("hippo", 8, 1), ("trace", 5, 1), ("steps", 7, 1), ("steps", 3, 2), ("ties", 8, 1),
                                               ("prealigned", 4, 1), ("synth2n", 16, 1)])
def test_batched_bases_follow_the_sequential_loop(shim, which, batch, lanes):
    """row f1, single-launch form (S4PCS_BATCH -> s4g_try_bases, on the CPU stand-in a loop over the per-base chain): the
    speculation logic -- bases selected ahead, consumed in order, RNG / pair-order replay rolled back on termination, tie
    resolution by re-running the adopted base -- leaves every observable equal to the reference's"""
    want = run_driver(which, "reference")
    if which == "ties":
        want = {"rows": [[True, True]] * 4}
    assert run_driver(which, "dropin", lanes=lanes, preload=shim, extra_env={"S4PCS_BATCH": str(batch)}) == want


def test_reference_pair_extraction_test_through_cpp_layer(shim):
    assert run_driver("pairtest", "dropin", preload=shim) == {"equal": [True, True]}


def _write_obj(path, pts, final_newline=True):
    """final_newline=True makes the reference's OBJ reader (and therefore ours) repeat the last vertex (cpp/io.cc)"""
    with open(path, "w") as f:
        f.write("\n".join("v %.9g %.9g %.9g" % tuple(p) for p in pts) + ("\n" if final_newline else ""))


@pytest.mark.parametrize("lanes", [1, 4])
def test_reference_demo_main_on_oracle_shim(shim, tmp_path, lanes):
    """the reference's demo main (compiled unchanged on our headers): CLI, OBJ reader, sampler, RANSAC driver,
    global transform, matrix + PLY writers -- all host code of the product -- with the stages answered by the oracle"""
    from super4pcs_b200 import build_cpp
    demo = build_cpp.build_all()["demo"]
    if not demo:
        pytest.skip("demo binary not built (needs the reference's demo source at build time)")
    gold = os.path.join(ROOT, "tests", "golden")
    h = np.load(os.path.join(gold, "hippo.npz"))
    g = dict(np.load(os.path.join(gold, "hippo_result.npz")))
    _write_obj(tmp_path / "a.obj", h["P"], final_newline=False)
    _write_obj(tmp_path / "b.obj", h["Q"], final_newline=False)
    mat, out = tmp_path / "mat.txt", tmp_path / "registered.ply"
    env = dict(os.environ, LD_PRELOAD=shim, S4PCS_LANES=str(lanes))
    r = subprocess.run([demo, "-i", str(tmp_path / "a.obj"), str(tmp_path / "b.obj"), "-o", "0.7", "-d", "0.01", "-t", "1000",
                        "-n", "200", "-m", str(mat), "-r", str(out)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Score: 0.64" in r.stdout
    M = np.array([ln.split() for ln in open(mat).read().splitlines()[2:6]], np.float64)
    assert np.abs(M - g["T_colmajor"].reshape(4, 4).T).max() < 1e-5
    xyz = common.read_ply_xyz(out)
    assert len(xyz) == len(h["Q"])
    assert np.abs(xyz[:64] - g["Q_transformed_head"]).max() < 1e-6


@pytest.mark.parametrize("args", [
    ["-o", "0.7", "-d", "0.01", "-t", "1000", "-n", "200"],                    # scripts/run-example.sh:68
    ["-o", "0.5", "-d", "0.02", "-t", "1000", "-n", "120", "-c", "0.6"],        # another sample size + terminate threshold
])
def test_demo_console_output_and_files_identical_to_the_reference_binary(shim, tmp_path, args):
    """the reference's demo executable (built from its own sources) and the same main on our headers + the stand-in:
    same console output line for line (progress lines aside), same matrix file, same registered PLY -- byte for byte.
    The input files end with a newline, so this also covers the reference reader's repeated last vertex."""
    from super4pcs_b200 import build_cpp
    from tests import build_ref_demo
    ours, ref = build_cpp.build_all()["demo"], build_ref_demo.build()
    if not ours or not ref:
        pytest.skip("needs the reference tree at build time")
    h = np.load(os.path.join(ROOT, "tests", "golden", "hippo.npz"))
    outs = {}
    for tag, exe, env in (("ref", ref, dict(os.environ)), ("ours", ours, dict(os.environ, LD_PRELOAD=shim, S4PCS_LANES="3"))):
        d = tmp_path / tag
        d.mkdir()
        _write_obj(d / "a.obj", h["P"])
        _write_obj(d / "b.obj", h["Q"])
        r = subprocess.run([exe, "-i", "a.obj", "b.obj", "-m", "mat.txt", "-r", "registered.ply"] + args, cwd=d,
                           capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        log = [ln for ln in r.stdout.replace("\r", "\n").splitlines() if ln.strip() and not ln.startswith("done:")]
        outs[tag] = (log, open(d / "mat.txt", "rb").read(), open(d / "registered.ply", "rb").read())
    assert outs["ref"][0] == outs["ours"][0]
    assert any(ln.startswith("Score:") for ln in outs["ref"][0])
    assert outs["ref"][1] == outs["ours"][1]
    assert outs["ref"][2] == outs["ours"][2]


def test_reference_pcl_wrapper_on_oracle_shim(shim, tmp_path):
    from tests import build_pcl_stub
    exe = build_pcl_stub.build()
    if exe is None:
        pytest.skip("pcl wrapper test binary not available")
    gold = os.path.join(ROOT, "tests", "golden")
    h = np.load(os.path.join(gold, "hippo.npz"))
    g = dict(np.load(os.path.join(gold, "hippo_result.npz")))
    for nme, arr in (("a.xyz", h["P"]), ("b.xyz", h["Q"])):
        np.savetxt(tmp_path / nme, arr, fmt="%.9g")
    r = subprocess.run([exe, str(tmp_path / "a.xyz"), str(tmp_path / "b.xyz"), "0.7", "0.01", "200"], capture_output=True,
                       text=True, timeout=600, env=dict(os.environ, LD_PRELOAD=shim))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Final score: 0.64" in r.stdout
    M = np.array([ln.split()[1:] for ln in r.stdout.splitlines() if ln.startswith("Score-matrix:")], np.float32)
    assert np.abs(M - g["T_colmajor"].reshape(4, 4).T).max() <= 1e-6


def test_without_a_device_the_real_library_refuses(shim, tmp_path):
    """no LD_PRELOAD: in a container without a GPU the product fails loudly instead of computing on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from super4pcs_b200 import build_cpp
    demo = build_cpp.build_all()["demo"]
    if not demo:
        pytest.skip("demo binary not built")
    h = np.load(os.path.join(ROOT, "tests", "golden", "hippo.npz"))
    _write_obj(tmp_path / "a.obj", h["P"][:500])
    _write_obj(tmp_path / "b.obj", h["Q"][:500])
    r = subprocess.run([demo, "-i", str(tmp_path / "a.obj"), str(tmp_path / "b.obj"), "-n", "100", "-d", "0.01"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr) or "s4g_create" in (r.stdout + r.stderr)


# ---- S4PCS_DEVICES: candidate-set sharding across several device contexts inside the C++ layer (SURVEY.md 8 row e) ----

@pytest.mark.parametrize("devices,lanes,fused,contexts,max_device", [("3", 1, 1, 3, 2), ("2", 3, 1, 6, 1), ("0,4,4,1", 1, 0, 4, 4)])
def test_hippo_sharded_over_device_contexts_matches_golden(shim, devices, lanes, fused, contexts, max_device):
    """a count (S4PCS_DEVICE .. +n-1) or an explicit ordinal list; with lanes every lane gets its own peers"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    st = {}
    r = run_driver("hippo", "dropin", lanes=lanes, fused=fused, preload=shim, stats=st, extra_env={"S4PCS_DEVICES": devices})
    assert np.float32(r["score"]) == g["score"] == np.float32(0.64)
    assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32))
    assert st["contexts"] == contexts and st["max_device"] == max_device and st["sharded_calls"] > 0, st


@needs_ref
@pytest.mark.parametrize("which,devices,lanes,fused", [("trace", "4", 1, 1), ("steps", "2", 4, 1), ("steps", "3", 1, 0),
                                                        ("sweep2", "5", 2, 1), ("trace", "0,0", 2, 0), ("synth2n", "8", 1, 1)])
def test_sharded_runs_are_indistinguishable_from_one_device_and_from_the_reference(shim, which, devices, lanes, fused):
    """visitor trace, stepwise termination + RNG state, random configurations: every observable equals the reference's
    (the shard maximum keeps the first-maximum rule: highest count, ties -> smallest quad index)"""
    want = run_driver(which, "reference")
    assert run_driver(which, "dropin", lanes=lanes, fused=fused, preload=shim, extra_env={"S4PCS_DEVICES": devices}) == want


@needs_ref
def test_exact_order_ties_with_sharded_candidates(shim):
    same = {"rows": [[True, True]] * 4}
    for devices, lanes, fused in (("3", 2, 1), ("2", 1, 0)):
        env = {"S4PCS_EXACT_ORDER": "1", "S4PCS_DEVICES": devices}
        assert run_driver("ties", "dropin", lanes=lanes, fused=fused, preload=shim, extra_env=env) == same


# ---- S4PCS_NCCL=1: the shards' records are reduced inside the library (s4g_comm_init_all; the stand-in's threads meet at a
# condition variable instead of in ncclAllReduce) and the C++ layer only checks that every context returned the same one

def test_hippo_sharded_with_the_library_side_reduction_matches_golden(shim):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    for devices, fused, contexts in (("3", 1, 3), ("2", 0, 2)):
        st = {}
        r = run_driver("hippo", "dropin", fused=fused, preload=shim, stats=st,
                       extra_env={"S4PCS_DEVICES": devices, "S4PCS_NCCL": "1"})
        assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32))
        assert st["contexts"] == contexts and st["sharded_calls"] > 0
        assert st["collectives"] == 2 * st["sharded_calls"] // contexts, st   # key max + record sum, once per base


@needs_ref
@pytest.mark.parametrize("which,devices,lanes,fused", [("trace", "4", 1, 1), ("steps", "3", 1, 0), ("sweep2", "5", 2, 1),
                                                        ("synth2n", "8", 1, 1)])
def test_library_side_reduction_is_indistinguishable_from_the_reference(shim, which, devices, lanes, fused):
    """(lanes > 1 is folded to one lane in this mode: communicators of different lanes must not interleave)"""
    want = run_driver(which, "reference")
    env = {"S4PCS_DEVICES": devices, "S4PCS_NCCL": "1"}
    assert run_driver(which, "dropin", lanes=lanes, fused=fused, preload=shim, extra_env=env) == want


def test_library_side_reduction_fails_loudly_without_nccl(shim):
    """no substitute transport: a missing NCCL is an exception, not a silent host merge"""
    env = dict(os.environ, LD_PRELOAD=shim, S4PCS_DEVICES="2", S4PCS_NCCL="1", S4G_SHIM_NO_NCCL="1")
    r = subprocess.run([sys.executable, DRIVER, "hippo", "dropin"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "S4PCS_NCCL" in r.stderr and "NCCL is not loadable" in r.stderr, r.stderr[-1500:]


def test_a_shard_that_fails_before_the_reduction_does_not_hang_the_others(shim):
    """the stand-in's reduction (like NCCL's) waits for every rank; cpp/shards.h ShardGate keeps the healthy shards out of
    it when a peer left early, and the peer's error is what the caller sees"""
    env = dict(os.environ, LD_PRELOAD=shim, S4PCS_DEVICES="3", S4PCS_NCCL="1", S4G_SHIM_FAIL_QUADS_ON_CONTEXT="1")
    r = subprocess.run([sys.executable, DRIVER, "hippo", "dropin"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "injected failure" in r.stderr, r.stderr[-1500:]


def test_device_list_parsing_is_forgiving(shim):
    """an empty / malformed S4PCS_DEVICES falls back to the single S4PCS_DEVICE context"""
    for spec, contexts in (("", 1), ("1", 1), ("0", 1), (",", 1), ("2,", 1), ("x", 1), ("all", 1)):
        st = {}
        run_driver("hippo", "dropin", preload=shim, stats=st, extra_env={"S4PCS_DEVICES": spec})
        assert st["contexts"] == contexts and st["sharded_calls"] == 0, (spec, st)


def test_all_devices_of_the_box(shim):
    """S4PCS_DEVICES=all asks the library how many devices there are (s4g_device_count; the stand-in pretends 3)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hippo_result.npz"))
    for first, contexts in (("0", 3), ("1", 2), ("5", 1)):
        st = {}
        r = run_driver("hippo", "dropin", preload=shim, stats=st,
                       extra_env={"S4PCS_DEVICES": "all", "S4G_SHIM_DEVICE_COUNT": "3", "S4PCS_DEVICE": first})
        assert np.array_equal(np.array(r["T"], np.uint32), g["T_colmajor"].view(np.uint32))
        assert st["contexts"] == contexts and st["max_device"] == max(2, int(first)), (first, st)


def test_shard_reduction_and_fan_out_unit(tmp_path):
    """cpp/shards.h on its own (tests/cpp/shards_test.cc): first-maximum rule across shards, error propagation"""
    exe = str(tmp_path / "shards_test")
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "cpp"), os.path.join(ROOT, "tests", "cpp", "shards_test.cc"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr


def timings_report(demo, tmp_path, preload=None, **env):
    """runs the demo on the hippo pair; returns the rows of the S4PCS_TIMINGS frame (all but the wall-clock one) or None"""
    h = np.load(os.path.join(ROOT, "tests", "golden", "hippo.npz"))
    if not os.path.exists(tmp_path / "a.obj"):
        _write_obj(tmp_path / "a.obj", h["P"], final_newline=False)
        _write_obj(tmp_path / "b.obj", h["Q"], final_newline=False)
    e = dict(os.environ, **env)
    if preload:
        e["LD_PRELOAD"] = preload
    r = subprocess.run([demo, "-i", str(tmp_path / "a.obj"), str(tmp_path / "b.obj"), "-o", "0.7", "-d", "0.01", "-t", "1000",
                        "-n", "200", "-m", str(tmp_path / "m.txt")], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0 and "Score: 0.64" in r.stdout, r.stderr[-2000:]
    lines = r.stdout.replace("\r", "\n").splitlines()
    at = [k for k, ln in enumerate(lines) if "Timings (msec)" in ln]
    return lines[at[0] + 2:at[0] + 7] if at else None


def test_stage_timings_report(shim, tmp_path):
    """S4PCS_TIMINGS=1: the run-time analogue of the reference's TEST_GLOBAL_TIMINGS report (match4pcsBase.hpp:77-83) -- the
    stage counters are those of the sequential loop for any lane / device-context count (the stand-in reports 1 ms per
    stage call, so the 'ms' columns count stage calls); nothing is printed without the switch"""
    from super4pcs_b200 import build_cpp
    demo = build_cpp.build_all()["demo"]
    if not demo:
        pytest.skip("demo binary not built (needs the reference's demo source at build time)")
    assert timings_report(demo, tmp_path, shim) is None
    want = timings_report(demo, tmp_path, shim, S4PCS_TIMINGS="1", S4PCS_BATCH="1")
    assert want is not None and "Bases tried             : 139" in want[-1] and "ordered pairs" in want[2]
    assert timings_report(demo, tmp_path, shim, S4PCS_TIMINGS="1", S4PCS_LANES="4", S4PCS_BATCH="1") == want
    # default = batched bases (s4g_try_bases): one set of stage calls per BATCH, the same output counts
    batched = timings_report(demo, tmp_path, shim, S4PCS_TIMINGS="1")
    assert [r.split("(device;")[1] for r in batched[:3]] == [r.split("(device;")[1] for r in want[:3]] and batched[-1] == want[-1]
    sharded = timings_report(demo, tmp_path, shim, S4PCS_TIMINGS="1", S4PCS_LANES="2", S4PCS_DEVICES="3")
    assert sharded[1:] == want[1:] and "candidates)" in sharded[0]      # (the Verify ms row is rank 0's share)
    assert sharded[0].split("(device;")[1] == want[0].split("(device;")[1]


def test_reference_compile_time_timings_switch(shim, tmp_path):
    """a caller built with the reference's own -DTEST_GLOBAL_TIMINGS gets the report without any environment switch"""
    ref_demo = os.path.join(_build.REFERENCE_ROOT, "demos", "Super4PCS", "super4pcs_test.cc") if hasattr(_build, "REFERENCE_ROOT") \
        else "/root/reference/demos/Super4PCS/super4pcs_test.cc"
    eig = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(ref_demo))), "3rdparty", "Eigen")
    if not os.path.exists(ref_demo) or not os.path.exists(os.path.join(eig, "Eigen", "Core")):
        pytest.skip("needs the reference's demo source and Eigen")
    exe = str(tmp_path / "demo_timings")
    libdir = os.path.join(ROOT, "super4pcs_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", "-DTEST_GLOBAL_TIMINGS", "-I", os.path.join(ROOT, "include"), "-I", eig,
                           "-I", os.path.dirname(os.path.dirname(ref_demo)), ref_demo, "-o", exe, "-L", libdir, "-lsuper4pcs_b200",
                           "-ls4g", "-Wl,-rpath," + libdir])
    rows = timings_report(exe, tmp_path, shim)
    assert rows is not None and "Bases tried             : 139" in rows[-1]
