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


def gpu_numa_cpus(device):
    """CPUs of the NUMA node GPU `device` hangs off (sysfs: the PCI device's numa_node -> that node's cpulist), or None when
    the box does not say (single node, virtualised PCI topology)."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never an error
        return None


def pin_to_gpu_node(device):
    """Restrict this PROCESS (its loader / writer threads, its pinned allocations by first touch) to the CPUs of the GPU's
    NUMA node, within what it is allowed already ("watch NUMA placement and PCIe root sharing").  Only for the
    one-process-per-GPU modes (torchrun ranks, the spawned shard workers of `probability.main`); returns the CPU set or None.
    SYKEPIC_NO_PIN=1 leaves the affinity alone."""
    if os.environ.get("SYKEPIC_NO_PIN"):
        return None
    cpus = gpu_numa_cpus(device)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) >= 4:  # never squeeze the host pipeline onto a sliver of cores
            os.sched_setaffinity(0, allowed)
            return allowed
    except (AttributeError, OSError):
        pass
    return None
