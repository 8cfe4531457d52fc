#!/usr/bin/env python
"""Copies / condenses the outputs of scripts/r02_final.sh (gpurun_out/r02z_*) into profiles/ (tracked)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
        print("copied", dst)


def summary(rep, dst, *notes, extra=()):
    if not os.path.exists(os.path.join(G, rep)):
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join("gpurun_out", rep), *notes],
                         capture_output=True, text=True, cwd=ROOT).stdout
    open(os.path.join(P, dst), "w").write(out)
    print("wrote", dst)


def launches():
    f = os.path.join(G, "r02z_launches.csv")
    if not os.path.exists(f):
        return
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, data = rows[hdr], rows[hdr + 1:]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    agg = {}
    for r in data:
        try:
            ms = float(r[mv].replace(",", "")) * scale.get(r[mu], 1e-6)
        except ValueError:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "r02_launches_final.txt"), "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000  python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
        o.write("# every launch of the bench process (cold-cache, serialised: compare SHARES); %d launches, %.2f ms in total\n" % (sum(v[0] for v in agg.values()), tot))
        o.write("# %-70s %8s %10s %7s\n" % ("kernel", "launches", "ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("%-72s %8d %10.3f %6.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
    print("wrote r02_launches_final.txt")


def main():
    cp("r02z_gpu_tests.txt", "r02_gpu_tests.txt")
    cp("r02z_bench_1gpu.json", "r02_bench_1gpu_final.json")
    cp("r02z_bench_reference.json", "r02_bench_reference_arm.json")
    cp("r02z_stage.jsonl", "r02_stage.jsonl")
    cp("r02z_demo_timing.txt", "r02_demo_timing.txt")
    cp("r02z_batch_bench.jsonl", "r02_batch_bench_hippo_final.jsonl")
    launches()
    cmd = "ncu --set full --clock-control none --import-source on -k regex:%s ... "
    summary("r02z_prof_verify.ncu-rep", "r02_verify_v14_ncu_summary.txt",
            "ROUND-2 FINAL k_verify (v14: delta-field + sub-voxel bits, two-level tile cull on a 4 MB SAT, per-warp queues); one launch = 4096 candidates x 1M queries (cfg2)",
            cmd % "k_verify" + "-s 3 -c 1  python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
    summary("r02z_prof_pairs.ncu-rep", "r02_pairs_v2_ncu_summary.txt",
            "ROUND-2 k_pairs v2: cfg1 stage, 50K points, 62.1M ordered pairs (first launch = the counting pass that overflowed the initial buffer)",
            cmd % "k_pairs" + "-c 3  python scripts/stage_bench.py cfg1")
    summary("r02z_prof_quads.ncu-rep", "r02_quads_ncu_summary.txt",
            "k_quad_query (count, fill) of the candidate-list generation and the six whole bases of scripts/stage_bench.py cfg1 (n = 3000 and 10000; the large launches are the n = 10000 bases with 2-7 M quads)",
            cmd % "k_quad_query" + "-c 12  python scripts/stage_bench.py cfg1")
    summary("r02z_prof_rigid.ncu-rep", "r02_rigid_ncu_summary.txt",
            "k_rigid of scripts/stage_bench.py cfg1 (the large launches are the n = 10000 bases with 2-7 M quads)",
            cmd % "k_rigid" + "-c 9  python scripts/stage_bench.py cfg1")
    rep = os.path.join(G, "r02z_prof_verify.ncu-rep")
    if os.path.exists(rep):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        open("/tmp/r02z_verify_src.csv", "w").write(src)
        blocks = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_sass_blocks.py"), "/tmp/r02z_verify_src.csv"], capture_output=True, text=True).stdout
        open(os.path.join(P, "r02_verify_v14_sass_blocks.txt"), "w").write(
            "# basic-block view (share of executed warp instructions, stall samples, active lanes) of the same capture as r02_verify_v14_ncu_summary.txt\n" + blocks)
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_verify_ncu_json.py"), os.path.join("gpurun_out", "r02z_prof_verify.ncu-rep"),
                        "ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline"], cwd=ROOT)
    with open(os.path.join(P, "r02_sanitizers.txt"), "w") as o:
        for f, what in (("r02z_sanitizer_memcheck.txt", "compute-sanitizer --tool memcheck  pytest tests/test_verify_gpu.py tests/test_pairs_gpu.py tests/test_batch_gpu.py (without the full-size / C++-layer cases)"),
                        ("r02z_sanitizer_racecheck.txt", "compute-sanitizer --tool racecheck  pytest tests/test_verify_gpu.py tests/test_pairs_gpu.py (small cases)")):
            p = os.path.join(G, f)
            if os.path.exists(p):
                o.write("# " + what + "\n" + "".join(open(p).readlines()[-6:]) + "\n")
    f = os.path.join(G, "r02z_clocks.csv")
    if os.path.exists(f):
        rows = [r.strip().split(", ") for r in open(f).readlines()[1:] if r.strip()]
        sm = sorted(float(r[1].split()[0]) for r in rows)
        act = [r for r in rows if any(x.strip() == "Active" for x in r[5:9])]
        open(os.path.join(P, "r02_clocks_summary.txt"), "w").write(
            "# nvidia-smi -lms 200 during the final bench.py run: %d samples, SM clock min / median / max = %.0f / %.0f / %.0f MHz (max %s), "
            "samples with hw_slowdown / thermal / power-cap active: %d\n" % (len(rows), sm[0], sm[len(sm) // 2], sm[-1], rows[0][2], len(act)))
    print(open(os.path.join(P, "verify_ncu.json")).read()[:300] if os.path.exists(os.path.join(P, "verify_ncu.json")) else "")


if __name__ == "__main__":
    main()
