#!/usr/bin/env python
"""K1 (preprocess) alone, at batch and bin granularity: CUDA-event time per launch and the HBM fraction.

    python tools/k1_bench.py [--t 224] [--reps 20]

Algorithmic bytes per ROI = w*h read + 16 B descriptor + T*T u8 written."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=224)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sizes", default="256,1024,4096,16384")
    args = ap.parse_args()
    import torch

    from sykepic_b200 import _lib, synth
    from tests.gpu_util import RawCtx

    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    ctx = RawCtx()
    dev = ctx.device
    sizes = [int(s) for s in args.sizes.split(",")]
    nmax = max(sizes)
    bins = []
    total = 0
    seed = 3000
    while total < nmax:
        b = synth.synth_bin(seed)
        seed += 1
        keep = b["w"] > 0
        bins.append(b)
        total += int(keep.sum())
    # concatenate the bins into one byte stream
    ws, hs, starts, chunks, off = [], [], [], [], 0
    for b in bins:
        keep = np.flatnonzero(b["w"] > 0)
        ws.append(b["w"][keep]); hs.append(b["h"][keep]); starts.append(b["start"][keep] + off)
        chunks.append(b["roi_bytes"]); off += len(b["roi_bytes"])
    w = np.concatenate(ws).astype(np.int32)[:nmax]
    h = np.concatenate(hs).astype(np.int32)[:nmax]
    start = np.concatenate(starts).astype(np.int64)[:nmax]
    roi = np.concatenate(chunks)
    t = args.t
    with torch.cuda.device(dev), torch.cuda.stream(ctx.stream):
        roi_d = torch.from_numpy(roi).to(dev)
        w_d, h_d, s_d = (torch.from_numpy(a).to(dev) for a in (w, h, start))
        out = torch.empty((nmax, t, t), dtype=torch.uint8, device=dev)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        for n in sizes:
            in_bytes = float((w[:n].astype(np.int64) * h[:n]).sum())
            alg = in_bytes + 16.0 * n + float(n) * t * t
            ms = []
            for r in range(args.reps + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ctx.stream)
                ctx.ck(ctx.lib.spk_preprocess(ctx.ctx, roi_d.data_ptr(), roi.size, s_d.data_ptr(), w_d.data_ptr(), h_d.data_ptr(), n, t, t,
                                              _lib.BORDER["mode"], 1, _lib.DTYPE_U8, _lib.LAYOUT_NCHW, None, out.data_ptr()))
                e1.record(ctx.stream)
                ctx.stream.synchronize()
                if r >= 3:
                    ms.append(e0.elapsed_time(e1))
            m = float(np.median(ms))
            gbs = alg / (m * 1e-3) / 1e9
            print(json.dumps({"kernel": "preprocess_u8_kernel", "n_rois": n, "T": t, "ms": m, "rois_per_s": n / (m * 1e-3),
                              "algorithmic_bytes_per_roi": alg / n, "achieved_gbs": gbs, "peak_gbs": peaks["hbm_gbs"],
                              "frac": gbs / peaks["hbm_gbs"]}), flush=True)
    assert ctx.fault_count() == 0
    ctx.close()


if __name__ == "__main__":
    main()
