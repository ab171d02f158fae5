"""Host stages of `sykepic prob` one by one ("adc parse, file read, CSV: report ROI/s each; not part of a GPU
roofline").  No GPU needed: everything here is the C-ABI host code and the Python around it, on synthetic IFCB bins.

    python tools/host_stages.py [--bins 8] [--rois 5000] [--threads 1]

Prints one JSON line: ROIs/s (single thread unless --threads) of
  adc_parse      spk_adc_parse on the `.adc` text (replaces the per-line parsing of utils/ifcb.py:101-110)
  roi_read       `.roi` file -> (pinned-size) host buffer with readinto, page cache warm
  validate       spk_rois_validate (the geometry checks the reference performs implicitly)
  csv_format     spk_format_prob_csv, 50 classes (replaces probabilities_to_csv, compute/probability.py:200-206)
  csv_write      the formatted text to a file (tmpfs)
  png_decode     png.read_gray per file / png.read_gray_many (the image mode's batch call, --threads C threads)
and, for scale, what the reference's own Python does for the first and the last (str.split per line; f-string per value)."""
import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sykepic_b200 import engine, png, synth  # noqa: E402


def rate(fn, items, n_rois, threads, min_s=1.0):
    """ROIs/s of fn over items (each item = one bin), repeated until min_s has passed."""
    reps, t0 = 0, time.perf_counter()
    while True:
        if threads > 1:
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(fn, items))
        else:
            for it in items:
                fn(it)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_s:
            return n_rois * reps / dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=8)
    ap.add_argument("--rois", type=int, default=5000)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--classes", type=int, default=50)
    args = ap.parse_args()
    tmp = Path(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    bins, names = [], []
    for i in range(args.bins):
        b = synth.synth_bin(1000 + i, args.rois)
        names.append(synth.bin_name(i))
        synth.write_bin(tmp, names[-1], b)
        bins.append(b)
    parsed = [engine.parse_adc(b["adc_text"]) for b in bins]
    n_rois = sum(len(p[0]) for p in parsed)
    roi_bytes = sum(int(b["roi_bytes"].size) for b in bins)
    classes = [f"class_{k}" for k in range(args.classes)]
    rng = np.random.default_rng(0)
    probs = [rng.dirichlet(np.ones(args.classes), len(p[0])).astype(np.float32) for p in parsed]
    texts = [b["adc_text"].encode() if isinstance(b["adc_text"], str) else b["adc_text"] for b in bins]
    bufs = [np.empty(int(b["roi_bytes"].size), np.uint8) for b in bins]

    def read_roi(i):
        with open(tmp / f"{names[i]}.roi", "rb") as fh:
            fh.readinto(memoryview(bufs[i]))

    def validate(i):
        _, w, h, start = parsed[i]
        engine.validate_rois(w, h, start, bufs[i].size, 224, 224)

    csv = [None] * args.bins

    def fmt(i):
        csv[i] = engine.format_prob_csv(classes, parsed[i][0], probs[i])

    def write(i):
        with open(tmp / f"{names[i]}.prob.csv", "wb") as fh:
            fh.write(csv[i])

    def ref_parse(i):  # what utils/ifcb.py:101-110 does per line
        out = []
        for k, line in enumerate(texts[i].decode().splitlines(), start=1):
            f = line.split(",")
            w, h, s = int(f[15]), int(f[16]), int(f[17])
            if w < 1 or h < 1:
                continue
            out.append((k, w, h, s))
        return out

    def ref_csv(i):  # what compute/probability.py:200-206 does
        rows = ["roi," + ",".join(classes)]
        for r, p in zip(parsed[i][0].tolist(), probs[i].tolist()):
            rows.append(f"{r}," + ",".join(f"{x:.5f}" for x in p))
        return "\n".join(rows) + "\n"

    idx = list(range(args.bins))
    th = args.threads
    res = {
        "adc_parse": rate(lambda i: engine.parse_adc(texts[i]), idx, n_rois, th),
        "roi_read": rate(read_roi, idx, n_rois, th),
        "validate": rate(validate, idx, n_rois, th),
        "csv_format": rate(fmt, idx, n_rois, th),
        "csv_write": rate(write, idx, n_rois, th),
        "reference_python_adc_parse": rate(ref_parse, idx[:2], sum(len(parsed[i][0]) for i in idx[:2]), 1, 0.5),
        "reference_python_csv_format": rate(ref_csv, idx[:1], len(parsed[0][0]), 1, 0.5),
    }
    # image mode: the first bin's ROIs as PNG files
    from tests.test_host_png import write_png_up

    rid, w, h, start = parsed[0]
    pdir = tmp / "png"
    pdir.mkdir()
    paths = []
    for i, ww, hh, s in zip(rid[:1500], w, h, start):
        p = pdir / f"{names[0]}_{int(i):05d}.png"
        write_png_up(p, bins[0]["roi_bytes"][int(s): int(s) + int(ww) * int(hh)].reshape(int(hh), int(ww)))
        paths.append(p)
    res["png_decode"] = rate(png.read_gray, paths, len(paths), 1)  # one file at a time (Python container parsing + C filters)
    res["png_decode_batch"] = rate(lambda _: png.read_gray_many(paths, th), [0], len(paths), 1)  # the image mode's call
    line = {"tool": "host_stages", "unit": "ROIs/s", "threads": th, "host_cpus": os.cpu_count(), "bins": args.bins, "rois": n_rois,
            "mean_roi_bytes": roi_bytes / n_rois, "csv_bytes_per_roi": len(csv[0]) / len(parsed[0][0]),
            "roi_read_gb_s": res["roi_read"] * roi_bytes / n_rois / 1e9, **{k: round(v) for k, v in res.items()}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
