"""Candidate-set sharding across the GPUs of one box.

Every rank holds the full clouds and grids (replicated in HBM) and processes the quads with
index % world == rank (`s4g_try_congruent_set*(..., shard_rank, shard_world)`).  The ONE collective
of the path is a max-allreduce of the packed 64-bit key

    key = (inlier_count << 32) | (0xFFFFFFFF - quad_index)

whose maximum is the highest count and, among equal counts, the SMALLEST quad index -- the
reference's first-maximum rule (strict '>' in quad order, match4pcsBase.hpp:468).  The owner's 4x4
(64 bytes) follows with one broadcast.  torch.distributed is plumbing only (NCCL over NVLink on
the GPUs, gloo in the CPU tests)."""
import numpy as np

KEY_NONE = 0


def pack_key(count, index):
    return (int(count) << 32) | (0xFFFFFFFF - int(index))


def unpack_key(key):
    key = int(key)
    if key == KEY_NONE:
        return 0, -1
    return key >> 32, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def shard_indices(K, rank, world):
    return np.arange(rank, K, world)


def reduce_best(local_key, local_T, device=None, group=None):
    """all ranks: returns (count, quad_index, T) of the global winner.  local_T: 16 floats."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        c, i = unpack_key(local_key)
        return c, i, np.asarray(local_T, np.float32)
    dev = device if device is not None else "cpu"
    k = torch.tensor([int(local_key)], dtype=torch.int64, device=dev)
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)            # the one collective of the path
    best = int(k.item())
    # the owner (exactly one rank: indices are disjoint) publishes its transform
    owner = torch.tensor([dist.get_rank(group) if (int(local_key) == best and best != KEY_NONE) else -1],
                         dtype=torch.int64, device=dev)
    dist.all_reduce(owner, op=dist.ReduceOp.MAX, group=group)
    T = torch.as_tensor(np.asarray(local_T, np.float32).copy(), device=dev)
    if int(owner.item()) >= 0:
        dist.broadcast(T, src=int(owner.item()), group=group)
    c, i = unpack_key(best)
    return c, i, T.cpu().numpy()
