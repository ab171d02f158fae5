#!/usr/bin/env python
"""End-to-end archive run (BASELINE.json configs[4]): synthetic IFCB bins ON DISK -> `probability.main` (read ->
decode -> CNN -> CSV text -> files), ROIs/s = non-empty ROIs / wall time from the first file open to the last CSV closed.

    python tools/archive_bench.py --bins 64                                  # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/archive_bench.py --bins 512   # bins sharded over ranks

Under torchrun every rank writes / processes its own shard (LPT by .roi size, sykepic_b200/shard.py); the only
exchange is the host-side merge of the processed-sample sets and the max-over-ranks time (no data-path collective).
The default is a slice of a day (a full day is ~1000 bins x ~5000 ROIs = 24 GB of .roi); rates do not depend on it."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=48)
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--root", default=None, help="work directory (default: a temp dir on /dev/shm if present)")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--devices", type=int, default=0,
                    help="single process, one host thread + pipeline per GPU for the first N GPUs (what `sykepic prob --gpus N` does); "
                         "default: one process per GPU under torchrun")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from sykepic_b200 import pipeline, shard, synth
    from sykepic_b200.compute import probability

    rank, world, local = shard.rank_world()
    if args.devices > 0:
        return threads_mode(args)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    base = args.root or ("/dev/shm" if Path("/dev/shm").is_dir() else None)
    root = Path(tempfile.mkdtemp(prefix=f"spk_archive_r{rank}_", dir=base))
    raw, out = root / "raw", root / "out"
    raw.mkdir()
    mdir = synth.write_model_dir(root / "model", arch=args.arch, t=224, head=(256, 128), seed=0, border="mode",
                                 imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)
    # this rank's share of the day: bin i belongs to rank i % world (the synthetic bins have similar sizes)
    mine = [i for i in range(args.bins) if i % world == rank]
    n_rois, roi_bytes = 0, 0
    for i in mine:
        b = synth.synth_bin(1000 + i)
        synth.write_bin(raw, synth.bin_name(i), b)
        n_rois += int((b["w"] > 0).sum())
        roi_bytes += len(b["roi_bytes"])
    paths = sorted(p.with_suffix("") for p in raw.glob("*.roi"))
    # warm-up on one bin (engine construction, first-launch costs), not timed
    warm = root / "warm_out"
    probability.main(paths[:1], mdir, warm, batch_size=args.batch_size, progress_bar=False, precision=args.precision, devices=[local])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = probability.main(paths, mdir, out, batch_size=args.batch_size, progress_bar=False, precision=args.precision, devices=[local])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    csv_bytes = sum(p.stat().st_size for p in out.rglob("*.prob.csv"))
    st = pipeline.LAST_STATS[-1]
    assert len(done) == len(paths), (len(done), len(paths))
    t = torch.tensor([dt, float(n_rois), float(roi_bytes), float(csv_bytes), float(len(paths)), st["run_s"]], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        wall = float(tmax[0])
        print(json.dumps({
            "metric": "ifcb_rois_per_s_archive_e2e", "value": float(t[1]) / float(tmax[5]), "unit": "ROIs/s", "n_gpus": world,
            "bins": int(t[4]), "rois": int(t[1]), "first_open_to_last_csv_s": float(tmax[5]),
            "wall_s_incl_engine_construction": wall, "value_incl_engine_construction": float(t[1]) / wall,
            "stage_seconds_rank0": {k: round(st[k], 4) for k in ("load_s", "gpu_wait_s", "write_s")}, "roi_gb": float(t[2]) / 1e9, "csv_gb": float(t[3]) / 1e9,
            "config": {"workload": f"{args.arch} 3x224x224 {args.precision}, synthetic IFCB bins on {base or 'tmp'}, "
                                   f"-b {args.batch_size} (launches of max(b, SYKEPIC_MIN_BATCH = {os.environ.get('SYKEPIC_MIN_BATCH', '1024')}) ROIs), "
                                   f"probability.main (read -> GPU -> %.5f CSV files)"},
            "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not args.keep:
        shutil.rmtree(root, ignore_errors=True)


def threads_mode(args):
    """`probability.main(devices=[0..N-1])` in ONE process: bins sharded by .roi size, one engine + pipeline per GPU thread."""
    import torch

    from sykepic_b200 import pipeline, synth
    from sykepic_b200.compute import probability

    base = args.root or ("/dev/shm" if Path("/dev/shm").is_dir() else None)
    root = Path(tempfile.mkdtemp(prefix="spk_archive_thr_", dir=base))
    raw, out = root / "raw", root / "out"
    mdir = synth.write_model_dir(root / "model", arch=args.arch, t=224, head=(256, 128), seed=0, border="mode",
                                 imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)
    n_rois = 0
    for i in range(args.bins):
        b = synth.synth_bin(1000 + i)
        synth.write_bin(raw, synth.bin_name(i), b)
        n_rois += int((b["w"] > 0).sum())
    paths = sorted(p.with_suffix("") for p in raw.glob("*.roi"))
    devs = list(range(args.devices))
    probability.main(paths[:len(devs)], mdir, root / "warm", batch_size=args.batch_size, progress_bar=False, precision=args.precision, devices=devs)
    for d in devs:
        torch.cuda.synchronize(d)
    n0 = len(pipeline.LAST_STATS)
    t0 = time.perf_counter()
    done = probability.main(paths, mdir, out, batch_size=args.batch_size, progress_bar=False, precision=args.precision, devices=devs)
    dt = time.perf_counter() - t0
    assert len(done) == len(paths), (len(done), len(paths))
    runs = pipeline.LAST_STATS[n0:]
    pipe_s = max(r["run_s"] for r in runs)
    print(json.dumps({
        "metric": "ifcb_rois_per_s_archive_e2e", "value": n_rois / pipe_s, "unit": "ROIs/s", "n_gpus": len(devs), "mode": "probability.main(devices=N): " + (os.environ.get("SYKEPIC_MULTI") or "auto (threads up to 2 GPUs, else a process per GPU)"),
        "bins": len(paths), "rois": n_rois, "first_open_to_last_csv_s": pipe_s, "wall_s_incl_engine_construction": dt,
        "value_incl_engine_construction": n_rois / dt, "bins_per_gpu": [r["bins"] for r in runs], "host_cpus": os.cpu_count()}), flush=True)
    if not args.keep:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
