"""TEST INFRASTRUCTURE, not collected by pytest: differential fuzz of the OBJ / PLY-ascii / PTX readers -- the product's
IOManager against the reference's -- on random, partly malformed files (grammar restricted to what the reference reads
without touching uninitialised memory or indexing out of range).
  python tests/fuzz_io_vs_reference.py [seed] [n_files]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import build_io_diff  # noqa: E402

ref, ours = build_io_diff.build()
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 300


def num():
    return rng.choice(["%.6g" % rng.uniform(-5, 5), "%d" % rng.randint(-9, 9), "1e-3", "-0", "%.3f" % rng.uniform(0, 1)])


def obj_file():
    lines, nv = ["v %s %s %s" % (num(), num(), num())], 1   # (the reference's x, y, z start uninitialised: first record complete)
    for _ in range(rng.randint(1, 40)):
        kind = rng.choice(["v", "v", "v", "vn", "blank", "comment", "space", "short", "other", "vlead"])
        if kind == "v":
            lines.append("v %s %s %s" % (num(), num(), num()) + ("" if rng.randint(0, 4) else " " + num()))
            nv += 1
        elif kind == "vlead":
            lines.append("  v\t%s  %s %s" % (num(), num(), num()))
            nv += 1
        elif kind == "vn":
            lines.append("vn %s %s %s" % (num(), num(), num()))
        elif kind == "short":
            lines.append(rng.choice(["v", "vn"]) + " " + " ".join(num() for _ in range(rng.randint(0, 3))))
        elif kind == "blank":
            lines.append("")
        elif kind == "space":
            lines.append("   ")
        elif kind == "comment":
            lines.append("# " + num())
        else:
            lines.append(rng.choice(["g grp", "usemtl m", "o obj", "s off", "vx 1 2 3"]))
    text = "\n".join(lines) + ("\n" if rng.randint(0, 2) else "")
    return "f.obj", text.replace("\n", "\r\n") if rng.randint(0, 6) == 0 else text


def ply_file():
    n = rng.randint(1, 12)
    layout = rng.choice(["xyz", "xyzn", "xyzc", "xyznc", "xyznca"])
    props = ["property float x", "property float y", "property float z"]
    if "n" in layout:
        props += ["property float nx", "property float ny", "property float nz"]
    if "c" in layout:
        props += ["property uchar red", "property uchar green", "property uchar blue"]
    if "a" in layout:
        props += ["property uchar alpha"]
    head = ["ply", "format ascii 1.0"] + (["comment fuzz"] if rng.randint(0, 2) else []) + ["element vertex %d" % n] + props + ["end_header"]
    rows = []
    for _ in range(n):
        vals = ["%.6g" % rng.uniform(-3, 3) for _ in range(6 if "n" in layout else 3)]
        cols = ["%d" % rng.randint(0, 256) for _ in range((4 if "a" in layout else 3) if "c" in layout else 0)]
        rows.append(" ".join(vals + cols))
    return "f.ply", "\n".join(head + rows) + "\n"


def ptx_file():
    cols, rows = rng.randint(1, 5), rng.randint(1, 5)
    body = ["%.5f %.5f %.5f %.3f %d %d %d" % (*rng.uniform(-2, 2, 3), rng.uniform(0, 1), *rng.randint(0, 256, 3))
            for _ in range(cols * rows - rng.choice([0, 0, 1]))]
    return "f.ptx", "\n".join(["%d" % cols, "%d" % rows] + ["0 0 0"] * 4 + ["1 0 0 0"] * 4 + body) + "\n"


bad = 0
with tempfile.TemporaryDirectory() as d:
    for k in range(n_files):
        name, text = [obj_file, obj_file, ply_file, ptx_file][rng.randint(0, 4)]()
        path = os.path.join(d, name)
        with open(path, "w", newline="") as f:
            f.write(text)
        outs = []
        for tag, exe in (("r", ref), ("o", ours)):
            r = subprocess.run([exe, path, os.path.join(d, tag + ".dump"), os.path.join(d, tag + "_out.ply"), os.path.join(d, tag + ".mat")],
                               capture_output=True, text=True)
            dump = open(os.path.join(d, tag + ".dump")).read() if r.returncode == 0 else "CRASH %d" % r.returncode
            written = b""
            for ext in ("ply", "obj"):
                p = os.path.join(d, tag + "_out." + ext)
                if os.path.exists(p):
                    written += open(p, "rb").read()
                    os.remove(p)
            outs.append((dump, written))
        if outs[0] != outs[1]:
            bad += 1
            keep = "/tmp/io_fuzz_fail_%d_%s" % (bad, name)
            open(keep, "w", newline="").write(text)
            print("DIFF", keep, outs[0][0].splitlines()[0] if outs[0][0] else "", "|", outs[1][0].splitlines()[0] if outs[1][0] else "")
print("files", n_files, "bad", bad)
