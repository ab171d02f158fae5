#!/usr/bin/env python
"""How busy is the GPU inside the product path?  `probability.main` over synthetic IFCB bins on tmpfs with the engine's
stamp profiler on: sum of the in-step kernel times per category against the wall time of the run (first file open to
last CSV closed).  The difference is GPU idle time: pipeline fill / drain, host-side gaps between launches, copies.

    python tools/e2e_gpu_busy.py [--bins 24] [--arch resnet18]"""
import argparse
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=16)  # (the stamp profiler holds 8192 launches: ~19 bins of 5000 ROIs)
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    import torch

    from sykepic_b200 import engine, pipeline, synth
    from sykepic_b200.compute import probability

    base = "/dev/shm" if Path("/dev/shm").is_dir() else None
    root = Path(tempfile.mkdtemp(prefix="spk_busy_", dir=base))
    raw, out = root / "raw", root / "out"
    raw.mkdir()
    mdir = synth.write_model_dir(root / "model", arch=args.arch, t=224, head=(256, 128), seed=0, border="mode",
                                 imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)
    n_rois = 0
    for i in range(args.bins):
        b = synth.synth_bin(1000 + i % 4)
        synth.write_bin(raw, synth.bin_name(i), b)
        n_rois += int((b["w"] > 0).sum())
    paths = sorted(p.with_suffix("") for p in raw.glob("*.roi"))
    eng = engine.Engine(mdir, precision="bf16", max_batch=args.batch, pre_chunk=4096)
    probability.main(paths[:4], mdir, root / "warm", batch_size=args.batch, force=True, progress_bar=False, precision="bf16", engine=eng)
    res = {}
    for mode in ("plain", "stamps"):
        if mode == "stamps":
            eng.profile_begin(stamps=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probability.main(paths, mdir, out / mode, batch_size=args.batch, force=True, progress_bar=False, precision="bf16", engine=eng)
        torch.cuda.synchronize()
        res[mode] = time.perf_counter() - t0
        if mode == "stamps":
            prof = eng.profile_read()
            eng.profile_end()
    busy = sum(v["ms"] for v in prof.values()) * 1e-3
    print(json.dumps({"bins": args.bins, "rois": n_rois, "wall_s": res, "rois_per_s": {k: n_rois / v for k, v in res.items()},
                      "gpu_kernel_s": busy, "gpu_busy_frac_of_stamped_run": busy / res["stamps"],
                      "per_category_s": {k: round(v["ms"] * 1e-3, 4) for k, v in prof.items()},
                      "launches": {k: v["launches"] for k, v in prof.items()}, "stage_seconds": pipeline.LAST_STATS[-1]}))
    eng.close()
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
