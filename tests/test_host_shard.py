"""Multi-GPU path on the CPU: bins sharded over ranks (one process per GPU) with world_size 2 over
`gloo`.  There is no collective on the data path; the only exchange is the host-side merge of the processed-sample
sets (what the reference's `probability.main` returns, sykepic/compute/probability.py:105-115)."""

import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

from sykepic_b200 import shard

ROOT = Path(__file__).resolve().parent.parent


def test_lpt_assignment_is_a_balanced_partition():
    rng = np.random.default_rng(3)
    costs = [int(c) for c in rng.integers(1, 40_000_000, 1000)]
    for n in (1, 2, 4, 8):
        parts = shard.assign(costs, n)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(costs)))  # every bin exactly once
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)  # LPT bound
        assert all(p == sorted(p) for p in parts)
    assert shard.assign([], 4) == [[], [], [], []]
    assert shard.assign([5, 1], 0) == [[0, 1]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, sys.argv[1])
    import torch.distributed as dist
    from sykepic_b200 import shard
    rank, world, local = shard.rank_world()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [f"D20210523T{h:02d}{m:02d}00_IFCB114" for h in range(6) for m in (0, 20, 40)]
    costs = [1000 + 37 * ((7 * i) % 11) for i in range(len(names))]
    mine = [names[i] for i in shard.assign(costs, world)[rank]]
    # "process" the shard: one rank skips a faulty bin, as the reference does (probability.py:106-114)
    done = {n for n in mine if not (rank == 1 and n == mine[0])}
    merged = shard.merge_processed(done)
    print(json.dumps({"rank": rank, "mine": mine, "done": sorted(done), "merged": sorted(merged)}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_ranks_over_gloo_shard_and_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        import json

        outs.append(json.loads(out.strip().splitlines()[-1]))
    a, b = sorted(outs, key=lambda o: o["rank"])
    assert not set(a["mine"]) & set(b["mine"]) and len(a["mine"]) + len(b["mine"]) == 18  # disjoint cover
    assert abs(len(a["mine"]) - len(b["mine"])) <= 1
    assert a["merged"] == b["merged"] == sorted(set(a["done"]) | set(b["done"]))  # every rank holds the union
    assert len(a["merged"]) == 17  # the skipped bin is in nobody's set
