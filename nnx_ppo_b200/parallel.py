"""Host-side plumbing of the env-sharded data-parallel path.

One process per GPU; `torch.distributed` carries the three real exchanges of the path (NCCL on
the GPU box, gloo in the CPU tests):
  * per update: all-reduce(SUM) of the two advantage moment sums (fp64) and of the flat gradient;
  * per iteration: all-gather of the per-rank Normalizer batch statistics, all-reduce of metrics.
Losses are divided by the GLOBAL sample count on every rank, so SUM yields the mean gradient and W
ranks are equivalent to one minibatch of W * mb envs.
"""
from __future__ import annotations

from . import prng


def dist_info():
    """(world_size, rank) — (1, 0) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except Exception:
        pass
    return 1, 0


def rank_keys(seed: int, rank: int):
    """Key derivation of new_training_state (ppo.py:544-548) for rank `rank` of a sharded run:
    rank 0 uses exactly the reference's keys; rank r > 0 folds r into both, so every rank is a
    bit-exact single-device run started from its own key."""
    key = prng.key(seed)
    key, training_key = prng.split(key)
    if rank > 0:
        key, training_key = prng.fold_in(key, rank), prng.fold_in(training_key, rank)
    return key, training_key


def all_reduce_sum(t, group=None):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_into(out, t, group=None):
    import torch.distributed as dist
    dist.all_gather_into_tensor(out, t, group=group)
    return out


def global_sample_count(rows_per_rank: int, world: int) -> int:
    return rows_per_rank * world
