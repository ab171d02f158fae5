"""Sharding of IFCB bins over the GPUs of one box.

Bins are independent (own .adc/.roi, own output CSV), so the path shards with no
collective: greedy longest-processing-time assignment by .roi size, one shard per GPU
(thread in `probability.main(devices=...)`, or rank under torchrun).  The host-side
merge is the union of the per-shard processed-sample sets (the value the reference's
`probability.main` returns, probability.py:105-115).
"""

import os
from pathlib import Path


def bin_cost(sample_path):
    try:
        return Path(sample_path).with_suffix(".roi").stat().st_size
    except OSError:
        return 0


def assign(costs, n_shards):
    """costs[i] -> list of index lists, one per shard; LPT greedy, deterministic."""
    n_shards = max(int(n_shards), 1)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * n_shards
    shards = [[] for _ in range(n_shards)]
    for i in order:
        j = min(range(n_shards), key=lambda s: (load[s], s))
        shards[j].append(i)
        load[j] += costs[i]
    for s in shards:
        s.sort()
    return shards


def assign_bins(sample_paths, n_shards):
    sample_paths = list(sample_paths)
    idx = assign([bin_cost(p) for p in sample_paths], n_shards)
    return [[sample_paths[i] for i in s] for s in idx]


def rank_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def merge_processed(local_set, group=None):
    """Union of every rank's processed-sample set on all ranks (torch.distributed all_gather_object;
    the only exchange of the multi-process path, off the hot path)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return set(local_set)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, sorted(local_set), group=group)
    out = set()
    for p in parts:
        out |= set(p)
    return out
