#!/usr/bin/env python
"""profiles/verify_ncu.json from an `ncu --set full` capture of k_verify on the bench workload: executed warp instructions
and DRAM bytes per launch, stamped with the digest of the kernel sources they belong to (bench.py ignores the file when
the sources or the workload have changed).
  python scripts/make_verify_ncu_json.py gpurun_out/x.ncu-rep "<ncu command line>" [candidates] [n_points]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(rep, cmd, candidates=4096, n_points=1000000):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def val(name, scale):
        v, u = float(r[col[name]]), units[col[name]]
        return v * scale.get(u, 1.0)
    byt = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tim = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    out = {"kernel": r[col["Kernel Name"]][:80], "report": os.path.basename(rep), "command": cmd,
           "source_digest": bench.source_digest(), "candidates": int(candidates), "n_points": int(n_points),
           "warp_instructions_per_launch": val("smsp__inst_executed.sum", {}),
           "dram_bytes_per_launch": val("dram__bytes_read.sum", byt) + val("dram__bytes_write.sum", byt),
           "ncu_duration_ms": val("gpu__time_duration.sum", tim),
           "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
           "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]]), "l1_hit_pct": float(r[col["l1tex__t_sector_hit_rate.pct"]])}
    json.dump(out, open(os.path.join(ROOT, "profiles", "verify_ncu.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(*sys.argv[1:])
