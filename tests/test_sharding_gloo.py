"""CPU test (gloo, world_size 2) of the multi-GPU host logic: candidate sharding + the single
max-allreduce of the packed key reproduce the unsharded winner (ties -> smallest index)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from super4pcs_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_key_packing_orders_like_the_reference():
    # higher count wins; equal count -> smaller index wins; "nothing verified" loses to everything
    assert sharding.pack_key(10, 5) > sharding.pack_key(9, 0)
    assert sharding.pack_key(10, 3) > sharding.pack_key(10, 4)
    assert sharding.pack_key(0, 2 ** 32 - 2) > sharding.KEY_NONE
    assert sharding.unpack_key(sharding.pack_key(123456, 789)) == (123456, 789)
    assert sharding.unpack_key(sharding.KEY_NONE) == (0, -1)
    idx = [sharding.shard_indices(10, r, 3) for r in range(3)]
    assert sorted(np.concatenate(idx).tolist()) == list(range(10))


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch.distributed as dist
    from super4pcs_b200 import sharding
    from oracle import port as oport
    from tests import common
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    delta = 0.05
    sc = common.scenario(3000, 0.4, delta)
    rng = np.random.RandomState(11)
    zP = (sc["P"] + sc["cp"])[:, 2]
    base = rng.choice(np.nonzero(np.abs(zP) < 0.15)[0], 4, replace=False).astype(np.int32)
    quads = common.congruent_like_quads(sc, base, 1500, 5)
    quads[700] = quads[3]                      # duplicate candidates -> ties across ranks
    quads[701] = quads[2]
    pt = oport.Port(sc["P"], sc["Q"], delta)
    whole = pt.try_congruent_set(base, quads, best_lcp_in=0.0)
    mine = sharding.shard_indices(len(quads), rank, world)
    # this rank's shard through the oracle (stands in for s4g_try_congruent_set(shard_rank, shard_world))
    T, rms, ok = pt.rigid_batch(base, quads[mine])
    gate = ok & (rms >= 0) & (rms < 2 * delta)
    key, Tbest = sharding.KEY_NONE, np.eye(4, dtype=np.float32).reshape(16)
    if gate.any():
        _, good, _ = pt.verify_batch(T[gate], 0.0)
        for g, t, qi in zip(good, T[gate], mine[gate]):
            k = sharding.pack_key(int(g), int(qi))
            if k > key:
                key, Tbest = k, t
    count, index, Tw = sharding.reduce_best(key, Tbest)
    assert index == whole["best_index"], (index, whole["best_index"])
    assert np.float32(count) / np.float32(len(sc["Q"])) == np.float32(whole["best_lcp"])
    assert np.array_equal(Tw.view(np.uint32), whole["T"].view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok", index, count)
""") % ROOT


def test_two_rank_gloo_reduction_matches_unsharded(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29000 + (os.getpid() % 2000))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
        assert "ok" in o
