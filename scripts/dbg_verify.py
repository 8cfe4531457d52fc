"""Debug helper: one small Verify call through the Python binding (python scripts/dbg_verify.py <n_points> <delta>); needs a GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from super4pcs_b200 import Context, synth
n, delta = int(sys.argv[1]), float(sys.argv[2])
d = synth.make_pair(n, 0.4, seed=3)
P, cp = synth.center(d["P"]); Q, cq = synth.center(d["Q"])
M = synth.candidate_transforms(32, delta, seed=5, n_near=8, centroid_p=cp, centroid_q=cq)
T = np.ascontiguousarray(M.transpose(0, 2, 1)).reshape(-1, 16)
with Context(0) as ctx:
    ctx.set_cloud_p(P, delta); print("P ok", ctx.grid_stats())
    ctx.set_cloud_q(Q); print("Q ok")
    print(ctx.verify(T))
