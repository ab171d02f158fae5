"""Compute class probabilities for raw IFCB data on B200 GPUs.

Drop-in for the reference's `sykepic.compute.probability` (file:line below are the
reference's): same entry points, argument meaning, CSV layout, skip / --force
behaviour and per-bin error handling, but no PNG round trip, no DataLoader workers
and no PyTorch execution -- each bin is decoded, transformed and classified by the
CUDA library (sykepic_b200/csrc) through its C ABI.

  call                  probability.py:27-64
  main                  probability.py:67-115   (returns the set of processed samples)
  prepare_model         probability.py:118-130
  process_sample        probability.py:133-162
  process_images        probability.py:165-177
  net_pass              probability.py:180-197
  probabilities_to_csv  probability.py:200-206

Extensions (keyword-only, defaults keep the reference's behaviour): `precision`
("fp32" | "bf16") and `devices` (GPU ids; bins are sharded over them, one engine
and one host thread per GPU, no collective -- SURVEY.md 8e).
"""

import os
import threading
from collections import namedtuple
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from .. import engine as _engine
from .. import shard
from ..utils import files, ifcb, logger

SOFTMAX_EXP = _engine.SOFTMAX_EXP
FILE_SUFFIX = ".prob"
log = logger.get_logger("prob")
EvalParams = namedtuple(
    "EvalParams",
    ["batch_size", "num_workers", "classes", "img_shape", "transform", "device"],
)
# "fp32_tc": fp32-level accuracy (within 1e-4 of the reference, measured <= 2e-5) on the tcgen05 tensor cores; "fp32" = the
# exact CUDA-core FFMA path (<= 1e-6, ~7x slower); "bf16" = within 2e-2, ~4x faster again
DEFAULT_PRECISION = os.environ.get("SYKEPIC_PRECISION", "fp32_tc")


ROI_FILE_LIMIT = 1e9  # bins whose `.roi` is larger are not processed (probability.py:45-51)


def _samples_from_images(img_files):
    """{sample: [its image files]} -- the sample is the file name up to its last `_` (probability.py:28-36)."""
    grouped = {}
    for img in img_files:
        grouped.setdefault(img.name.rpartition("_")[0], []).append(img)
    return grouped


def _samples_from_bins(sample_paths):
    """The bins to process: those whose `.roi` is within ROI_FILE_LIMIT; the others are reported and left out."""
    kept = []
    for sample_path in sample_paths:
        if sample_path.with_suffix(".roi").stat().st_size > ROI_FILE_LIMIT:
            log.warning(f"{sample_path.name} is over 1G, skipping")
        else:
            kept.append(sample_path)
    return kept


def call(args):
    """Entry point of `sykepic prob` (argparse namespace, or any object with the same attributes): one of
    `raw` (directory of bins) / `samples` (bin paths without suffix) / `image_dir` / `images` (ROI PNG files)."""
    image_dir, images = getattr(args, "image_dir", None), getattr(args, "images", None)
    as_images = bool(image_dir or images)
    if as_images:
        pngs = Path(image_dir).rglob("*.png") if image_dir else (Path(p) for p in images)
        work = _samples_from_images(sorted(pngs))
    else:
        raw_dir = getattr(args, "raw", None)
        work = _samples_from_bins(files.list_sample_paths(raw_dir) if raw_dir else [Path(p) for p in args.samples])
    return main(work, args.model, args.out, args.batch_size, args.num_workers, args.force, progress_bar=True,
                samples_as_images=as_images, precision=getattr(args, "precision", None), devices=getattr(args, "devices", None))


def _devices(devices):
    import torch

    if devices is None:
        env = os.environ.get("SYKEPIC_DEVICES")
        if env:
            devices = [int(d) for d in env.split(",") if d.strip()]
        else:
            devices = [int(os.environ.get("LOCAL_RANK", 0))] if torch.cuda.is_available() else [0]
    elif isinstance(devices, int):
        devices = list(range(devices))
    devices = list(devices)
    if not devices:
        raise ValueError("sykepic prob: at least one GPU is needed (devices / --gpus / SYKEPIC_DEVICES is empty or 0)")
    return devices


def main(
    sample_paths,
    model_dir,
    out_dir,
    batch_size=64,
    num_workers=2,
    force=False,
    progress_bar=True,
    samples_as_images=False,
    *,
    precision=None,
    devices=None,
    engine=None,
):
    """`num_workers` is accepted for compatibility; there are no loader processes here.
    `engine`: an already built `Engine` for `model_dir` on the (single) device -- callers that process several batches
    of bins with one model (services, bench.py) skip the per-call model construction the reference pays (:77)."""
    devices = [engine.device.index] if engine is not None else _devices(devices)
    precision = precision or DEFAULT_PRECISION
    spec = engine.spec if engine is not None else _engine.ModelSpec.from_dir(model_dir)
    # `batch_size` is the reference's DataLoader batch (CLI default 64).  Results do not depend on how ROIs are batched
    # (tests/test_gpu_network.py::test_batch_split_and_partial_batches), and the GPU wants large launches: ResNet-18 runs
    # at 263 k ROI/s with 256 ROIs per launch sequence, 280 k with 512, 284 k with 1024 -- so smaller requests are raised.
    max_batch = max(int(batch_size), int(os.environ.get("SYKEPIC_MIN_BATCH", "1024")), 1)
    if engine is not None:
        max_batch = engine.max_batch

    def make_params(dev):
        net = engine if engine is not None else _engine.Engine(spec, device=dev, precision=precision, max_batch=max_batch)
        return net, EvalParams(batch_size=max_batch, num_workers=num_workers, classes=spec.classes,
                               img_shape=spec.img_shape, transform=None, device=net.device)

    if samples_as_images:
        net, params = make_params(devices[0])
        try:
            items = list(sample_paths.items())
            csv_of = lambda sample: Path(out_dir) / f"{sample}{FILE_SUFFIX}.csv"  # noqa: E731
            # the next sample's files are decoded (host library threads, no interpreter lock) while this one is on the GPU
            with ThreadPoolExecutor(1) as pool:
                def ahead(i):
                    if i < len(items) and (force or not csv_of(items[i][0]).is_file()):
                        return pool.submit(decode_images, items[i][1])
                    return None

                nxt = ahead(0)
                for i, (sample, img_paths) in enumerate(_progress(items, progress_bar)):
                    cur, nxt = nxt, ahead(i + 1)
                    process_images(img_paths, net, params, csv_of(sample), force, decoded=cur.result() if cur else None)
        finally:
            if engine is None:
                net.close()
        return None

    sample_paths = list(sample_paths)
    samples_processed = set()
    bar = _progress_bar(len(sample_paths), progress_bar)

    def run_on(dev, paths):
        """One GPU: read -> GPU -> CSV pipeline over its bins (sykepic_b200/pipeline.py)."""
        from .. import pipeline

        net, params = make_params(dev)
        try:
            pipe = pipeline.BinPipeline(net, params.classes, out_dir, batch_size=params.batch_size, force=force, suffix=FILE_SUFFIX)
            return pipe.run(paths, progress=bar)
        finally:
            if engine is None:
                net.close()

    try:
        if len(devices) <= 1:
            return run_on(devices[0], sample_paths)

        # ---- bins sharded over the GPUs of the box; host-side merge = union of the per-GPU sets.
        # One PROCESS per GPU: with one thread per GPU in a single interpreter the host stages contend for the GIL
        # (measured on 8 GPUs, 128 bins: 0.52 M ROI/s with threads, 1.09 M with processes).
        shards = shard.assign_bins(sample_paths, len(devices))
        # (spawning costs ~4 s of start-up per job: threads for small jobs and up to 2 GPUs)
        mode = os.environ.get("SYKEPIC_MULTI") or ("processes" if len(devices) > 2 and len(sample_paths) >= 8 * len(devices) else "threads")
        if mode == "threads":
            results = [set() for _ in devices]
            errors = []

            def worker(i):
                try:
                    results[i] = run_on(devices[i], shards[i])
                except Exception as e:  # engine construction failed: report, do not hang the others
                    errors.append(e)

            threads = [threading.Thread(target=worker, args=(i,), name=f"spk-gpu{devices[i]}") for i in range(len(devices))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
            for r in results:
                samples_processed |= r
            return samples_processed
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        queue = ctx.Queue()
        procs = [ctx.Process(target=_shard_process, name=f"spk-gpu{dev}",
                             args=(dev, [str(p) for p in shards[i]], str(model_dir), str(out_dir), batch_size, num_workers, force, precision, queue))
                 for i, dev in enumerate(devices)]
        for pr in procs:
            pr.start()
        failures = []
        reported = set()
        import queue as _queue

        def take(msg):
            dev, done, err, stats = msg
            reported.add(dev)
            if err:
                failures.append(f"GPU {dev}: {err}")
            samples_processed.update(done)
            if stats:
                from .. import pipeline

                pipeline.LAST_STATS.append(stats)
            if bar is not None:
                bar.update(len(shards[devices.index(dev)]))  # every bin of the shard: processed, skipped or failed

        # A worker that dies without reporting (segfault, OOM kill, CUDA abort) must not hang the parent: poll the queue
        # and watch the children; one that has exited without a message is recorded as a failure.
        while len(reported) < len(procs):
            try:
                take(queue.get(timeout=1.0))
                continue
            except _queue.Empty:
                pass
            for dev, pr in zip(devices, procs):
                if dev not in reported and not pr.is_alive():
                    try:  # its message may have arrived between the time-out and the liveness check
                        while True:
                            take(queue.get_nowait())
                    except _queue.Empty:
                        pass
                    if dev not in reported:
                        reported.add(dev)
                        failures.append(f"GPU {dev}: worker process exited with code {pr.exitcode} without reporting")
        for pr in procs:
            pr.join()
        if failures:
            raise RuntimeError("; ".join(failures))
        return samples_processed
    finally:
        if bar is not None:
            bar.close()


def _shard_process(dev, paths, model_dir, out_dir, batch_size, num_workers, force, precision, queue):
    """Worker process of one GPU: its shard of the bins through `main` on that device; reports the processed set."""
    try:
        from .. import pipeline

        shard.pin_to_gpu_node(dev)  # this process serves one GPU: keep its host threads on that GPU's NUMA node
        done = main([Path(p) for p in paths], model_dir, out_dir, batch_size, num_workers, force, progress_bar=False,
                    precision=precision, devices=[dev])
        queue.put((dev, sorted(done), None, pipeline.LAST_STATS[-1] if pipeline.LAST_STATS else None))
    except Exception as e:  # engine construction / CUDA failure: the parent raises after the others finish
        queue.put((dev, [], repr(e), None))


class _LockedBar:
    """tqdm shared by the writer threads of every GPU pipeline."""

    def __init__(self, bar):
        self.bar, self.lock = bar, threading.Lock()

    def update(self, n):
        with self.lock:
            self.bar.update(n)

    def close(self):
        self.bar.close()


def _progress_bar(total, progress_bar):
    if not progress_bar:
        return None
    try:
        from tqdm import tqdm
    except ImportError:
        return None
    return _LockedBar(tqdm(total=total, desc="Processing samples"))


def _progress(items, progress_bar):
    if not progress_bar:
        return items
    try:
        from tqdm import tqdm
    except ImportError:
        return items
    return tqdm(items, desc="Processing samples")


def prepare_model(model_dir, precision=None, device=None, max_batch=256):
    """-> (net, classes, img_shape, eval_transform, device) like the reference; `net` is an Engine,
    the eval transform is part of it (K1) and reported as None."""
    devs = _devices(None if device is None else [device])
    net = _engine.Engine(model_dir, device=devs[0], precision=precision or DEFAULT_PRECISION, max_batch=max_batch)
    return net, net.spec.classes, net.spec.img_shape, None, net.device


def process_sample(sample_path, net, params, out_dir, force=False):
    sample_path = Path(sample_path)
    sample = sample_path.name
    csv_path = files.sample_csv_path(sample_path, out_dir, suffix=FILE_SUFFIX)
    if csv_path.is_file():
        if force:
            log.warning(f"{csv_path.name} already exists, overwriting")
        else:
            log.warning(f"{csv_path.name} already exists, skipping")
            return sample
    log.debug(f"Computing probabilities for {sample}")
    # appending the suffix (not with_suffix) would be the literal mirror; bin names have no dots
    with open(sample_path.with_suffix(".adc"), "rb") as fh:
        adc = fh.read()
    roi = np.fromfile(sample_path.with_suffix(".roi"), dtype=np.uint8)
    roi_id, probs = net.run_bin(adc, roi, batch_size=params.batch_size)
    _write_csv(roi_id, probs, params.classes, csv_path)
    return sample


def process_images(img_paths, net, params, csv_path, force=False, decoded=None):
    """`--image-dir` / `--images` mode: ROIs come from PNG files named <sample>_<roi>.png
    (probability.py:165-177, :190).  The PNGs are decoded on the host (gray, as cv2.imread's three
    equal planes) and packed into one byte stream, then take the same device path as a raw bin.
    `decoded`: the result of `decode_images(img_paths)` if the caller already has it (main decodes one sample ahead)."""
    csv_path = Path(csv_path)
    if csv_path.is_file():
        if force:
            log.warning(f"{csv_path.name} already exists, overwriting")
        else:
            log.warning(f"{csv_path.name} already exists, skipping")
            return
    ids, w, h, start, data = decoded if decoded is not None else decode_images(img_paths)
    if len(ids):
        probs = net.run_rois(ids, w, h, start, data, batch_size=params.batch_size)
        _write_csv(ids, probs, params.classes, csv_path)  # sorted by ROI id there (probability.py:197)
    else:
        probabilities_to_csv([], params.classes, csv_path)


def decode_images(img_paths):
    """All PNG files of a sample -> (roi ids, w, h, start, data): one byte stream laid out like a `.roi` file, decoded by
    the host library's threads (sykepic_b200/png.py).  ROI id = last `_` field of the file stem (probability.py:190)."""
    from .. import png

    img_paths = list(img_paths)
    ids = np.array([int(Path(p).stem.split("_")[-1]) for p in img_paths], np.int32)
    w, h, start, data = png.read_gray_many(img_paths)
    return ids, w, h, start, data


def net_pass(net, rois, device=None, batch_size=None):
    """[(roi_id, (h,w) uint8 array), ...] -> [(roi_id, [p, ...]), ...] sorted by ROI id
    (probability.py:180-197; the reference takes a DataLoader over PNG files here)."""
    rois = list(rois)
    if not rois:
        return []
    w = np.array([r.shape[1] for _, r in rois], np.int32)
    h = np.array([r.shape[0] for _, r in rois], np.int32)
    area = w.astype(np.int64) * h
    start = np.concatenate([[0], np.cumsum(area)[:-1]]).astype(np.int64)
    data = np.concatenate([np.ascontiguousarray(r, dtype=np.uint8).ravel() for _, r in rois])
    ids = np.array([i for i, _ in rois], np.int32)
    probs = net.run_rois(ids, w, h, start, data, batch_size=batch_size)
    return sorted(zip(ids.tolist(), probs.tolist()))


def probabilities_to_csv(probabilities, classes, csv_path):
    """[(roi, [p...]), ...] -> csv_path; header `roi,<classes>`, `%.5f` values (probability.py:200-206)."""
    probabilities = list(probabilities)
    roi_id = np.array([r for r, _ in probabilities], np.int32)
    probs = np.array([p for _, p in probabilities], np.float32).reshape(len(probabilities), len(classes))
    _write_csv(roi_id, probs, classes, csv_path)


def _write_csv(roi_id, probs, classes, csv_path):
    csv_path = Path(csv_path)
    csv_path.parent.mkdir(parents=True, exist_ok=True)
    # results sorted by ROI id (probability.py:197); .adc order is already ascending
    if len(roi_id) > 1 and np.any(np.diff(roi_id) < 0):
        order = np.argsort(roi_id, kind="stable")
        roi_id, probs = roi_id[order], probs[order]
    with open(csv_path, "wb") as fh:
        fh.write(_engine.format_prob_csv(classes, roi_id, probs))
