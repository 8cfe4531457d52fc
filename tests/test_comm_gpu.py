"""Row e inside the library (include/s4g.h, csrc/comm.cu): with an NCCL communicator attached to the contexts the shards'
winners are reduced on the device -- ncclAllReduce(ncclMax) of the packed key, then ncclAllReduce(ncclSum) of the record
the non-owners zeroed -- and every rank returns the global record.  It must be the record the unsharded call returns
(which tests/test_verify_gpu.py pins against the oracle), bit for bit.  One GPU: a communicator of one rank runs the same
chain (NCCL copies in place).  Two or more GPUs: one context per device in this process (s4g_comm_init_all), one host
thread per context; the process-per-GPU form (s4g_comm_init_rank) is what bench.py runs under torchrun."""
import threading

import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _setup(ctx, sc):
    ctx.set_cloud_p(sc["P"], sc["delta"])
    ctx.set_cloud_q(sc["Q"])


def _case(n=3000, delta=0.05, k=3000):
    sc = common.scenario(n, 0.4, delta)
    rng = np.random.RandomState(11)
    zP = (sc["P"] + sc["cp"])[:, 2]
    base = rng.choice(np.nonzero(np.abs(zP) < 0.15)[0], 4, replace=False).astype(np.int32)
    return sc, base, common.congruent_like_quads(sc, base, k, 5)


def _same(a, b):
    assert a["key"] == b["key"] and a["best_index"] == b["best_index"] and a["best_count"] == b["best_count"]
    assert a["n_gate_pass"] == b["n_gate_pass"] and a["n_q"] == b["n_q"]
    assert np.array_equal(common.bits(a["T"]), common.bits(b["T"])) and np.array_equal(a["quad"], b["quad"])
    assert np.array_equal(common.bits(a["centroid1"]), common.bits(b["centroid1"]))
    assert np.array_equal(common.bits(a["centroid2"]), common.bits(b["centroid2"]))
    assert np.float32(a["rms"]).view(np.uint32) == np.float32(b["rms"]).view(np.uint32)


def test_one_rank_communicator_runs_the_reduction_chain(s4g_lib):
    from super4pcs_b200 import Context, s4g
    sc, base, quads = _case()
    with Context(0) as ctx:
        _setup(ctx, sc)
        want = ctx.try_congruent_set(sc["P"][base], quads, 2 * sc["delta"])
        T = common.candidates_colmajor(sc, 300)
        counts = ctx.verify(T)
        assert ctx.comm_info()["ranks"] == 0
        plain_counts, plain_key = ctx.verify_best(T)                       # no communicator: the local maximum
        s4g.comm_init_all([ctx])
        info = ctx.comm_info()
        assert info["ranks"] == 1 and info["rank"] == 0 and info["nccl_version"] >= 22000
        got = ctx.try_congruent_set(sc["P"][base], quads, 2 * sc["delta"])
        _same(got, want)
        assert ctx.comm_info()["collectives"] == 2                         # key max + record sum
        empty = ctx.try_congruent_set(sc["P"][base], np.zeros((0, 4), np.int32), 2 * sc["delta"])
        assert empty["best_index"] == -1 and empty["key"] == 0 and empty["n_gate_pass"] == 0
        idx = np.arange(len(T), dtype=np.uint32) * 3 + 5
        c2, key = ctx.verify_best(T, idx)
        assert np.array_equal(c2, counts) and np.array_equal(plain_counts, counts)
        k = int(np.argmax(counts))                                          # first maximum
        assert key == (int(counts[k]) << 32) | (0xFFFFFFFF - int(idx[k]))
        assert plain_key == (int(counts[k]) << 32) | (0xFFFFFFFF - k)
        ctx.comm_destroy()
        assert ctx.comm_info()["ranks"] == 0
        _same(ctx.try_congruent_set(sc["P"][base], quads, 2 * sc["delta"]), want)


def test_shard_arguments_must_match_the_communicator(s4g_lib):
    from super4pcs_b200 import Context, S4GError, s4g
    sc, base, quads = _case(k=50)
    with Context(0) as ctx:
        _setup(ctx, sc)
        s4g.comm_init_all([ctx])
        with pytest.raises(S4GError, match="communicator"):
            ctx.try_congruent_set(sc["P"][base], quads, 2 * sc["delta"], shard_rank=1, shard_world=2)
        with pytest.raises(S4GError, match="already attached"):
            s4g.comm_init_all([ctx])


def _threads(fns):
    out, err = [None] * len(fns), [None] * len(fns)

    def run(i):
        try:
            out[i] = fns[i]()
        except Exception as e:  # noqa: BLE001 -- re-raised below
            err[i] = e

    ts = [threading.Thread(target=run, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    for e in err:
        if e is not None:
            raise e
    return out


def _world():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_world() < 2", reason="needs two GPUs")
def test_shards_on_several_gpus_return_the_global_record(s4g_lib):
    from super4pcs_b200 import Context, s4g
    W = min(_world(), 8)
    sc, base, quads = _case(n=20000, delta=0.02, k=6000)
    ctxs = [Context(d) for d in range(W)]
    try:
        for c in ctxs:
            _setup(c, sc)
        want = ctxs[0].try_congruent_set(sc["P"][base], quads, 2 * sc["delta"])
        assert want["best_index"] >= 0 and want["n_gate_pass"] > 0
        T = common.candidates_colmajor(sc, 1000)
        counts = ctxs[0].verify(T)
        s4g.comm_init_all(ctxs)
        for r, c in enumerate(ctxs):
            assert c.comm_info()["ranks"] == W and c.comm_info()["rank"] == r
            c.comm_set_timeout(60)
        for _ in range(3):                                                  # the communicator is reusable
            got = _threads([lambda c=c, r=r: c.try_congruent_set(sc["P"][base], quads, 2 * sc["delta"], shard_rank=r,
                                                                 shard_world=W) for r, c in enumerate(ctxs)])
            for g in got:
                _same(g, want)
        # no candidate anywhere: rank 0's empty record on every rank
        none = _threads([lambda c=c, r=r: c.try_congruent_set(sc["P"][base], np.array([[4, 4, 4, 4]] * 10, np.int32),
                                                              2 * sc["delta"], shard_rank=r, shard_world=W)
                         for r, c in enumerate(ctxs)])
        for g in none:
            assert g["best_index"] == -1 and g["key"] == 0 and g["n_gate_pass"] == 0
        # s4g_verify_best on index % W shards: every rank gets the key of the whole list
        k = int(np.argmax(counts))
        want_key = (int(counts[k]) << 32) | (0xFFFFFFFF - k)
        res = _threads([lambda c=c, r=r: c.verify_best(T[r::W], np.arange(r, len(T), W, dtype=np.uint32))
                        for r, c in enumerate(ctxs)])
        for r, (c_r, key) in enumerate(res):
            assert key == want_key and np.array_equal(c_r, counts[r::W])
    finally:
        for c in ctxs:
            c.close()


_LATE_PEER = '''
import os, sys, time
sys.path.insert(0, %r)
from super4pcs_b200 import Context, S4GError, s4g
from tests import common
rank, idfile = int(sys.argv[1]), sys.argv[2]
sc = common.scenario(3000, 0.4, 0.05)
ctx = Context(rank)
ctx.set_cloud_p(sc["P"], sc["delta"]); ctx.set_cloud_q(sc["Q"])
if rank == 0:
    with open(idfile + ".tmp", "wb") as f:
        f.write(s4g.comm_unique_id())
    os.rename(idfile + ".tmp", idfile)
t0 = time.time()
while not os.path.exists(idfile) and time.time() - t0 < 60:
    time.sleep(0.05)
ctx.comm_init_rank(open(idfile, "rb").read(), 2, rank)          # collective; runs NCCL's transport set-up too
if rank == 1:
    print("PEER_IDLE"); sys.stdout.flush()
    time.sleep(25)                                               # never reaches the reduction
    os._exit(0)
ctx.comm_set_timeout(2)
t0 = time.time()
try:
    ctx.verify_best(common.candidates_colmajor(sc, 8))
    print("RETURNED")
except S4GError as e:
    print("GAVE_UP after %%.1f s:" %% (time.time() - t0), e)
sys.stdout.flush()
os._exit(0)                                                      # (the communicator is gone; nothing to tidy up)
'''


_LATE_THREAD = '''
import os, sys, time
sys.path.insert(0, %r)
from super4pcs_b200 import Context, S4GError, s4g
from tests import common
sc = common.scenario(3000, 0.4, 0.05)
ctxs = [Context(0), Context(1)]
for c in ctxs:
    c.set_cloud_p(sc["P"], sc["delta"]); c.set_cloud_q(sc["Q"])
s4g.comm_init_all(ctxs)
ctxs[0].comm_set_timeout(2)
t0 = time.time()
try:
    ctxs[0].verify_best(common.candidates_colmajor(sc, 8))       # the context on device 1 never calls
    print("RETURNED")
except S4GError as e:
    print("GAVE_UP after %%.1f s:" %% (time.time() - t0), e)
sys.stdout.flush()
os._exit(0)
'''


@pytest.mark.skipif("_world() < 2", reason="needs two GPUs")
@pytest.mark.parametrize("form", ["processes", "threads"])
def test_a_peer_that_never_arrives_is_a_timeout_not_a_hang(s4g_lib, tmp_path, form):
    """both forms (s4g_comm_init_rank between two processes, s4g_comm_init_all inside one): the collective is a
    stream-ordered launch and the wait for it has a deadline -- taken BEFORE the result copies, which would block the host"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    idfile = str(tmp_path / "nccl_id")
    if form == "processes":
        cmds = [[sys.executable, "-c", _LATE_PEER % root, str(r), idfile] for r in (0, 1)]
    else:
        cmds = [[sys.executable, "-c", _LATE_THREAD % root]]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=root) for c in cmds]
    out0 = err0 = ""
    try:
        out0, err0 = procs[0].communicate(timeout=60)
    except subprocess.TimeoutExpired:
        pass
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert "GAVE_UP" in out0 and "time limit" in out0, (form, out0[-500:], err0[-1500:])
