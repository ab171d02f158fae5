"""Per-GPU host pipeline for `sykepic prob` on raw bins: read -> GPU -> CSV, overlapped.

The reference handles one bin at a time on one thread (sykepic/compute/probability.py:105-115,
133-162): extract PNGs, run the DataLoader / network, format and write the CSV.  Here the three
host stages of a bin run on different threads so that the GPU never waits for the disk or for the
text formatter:

    loader threads     .adc + .roi -> parsed descriptors + the byte stream in a pinned buffer, ONE C-ABI call per
                       bin (`spk_bin_load`: file reads, parse, geometry checks) that holds no interpreter lock
    GPU thread         H2D, K1, K2 + K3 per batch, D2H into pinned buffers -- bin i+1 is submitted
                       before bin i's results are awaited, so the stream always has work queued
    writer threads     `%.5f` CSV text + file write, one C-ABI call per bin (`spk_prob_csv_write`)

Per-bin error policy is the reference's (probability.py:106-114): a bin that raises is logged
("Faulty raw data" for ValueError, "Unexpected error" otherwise) and skipped, the others go on.
There is no collective and no cross-bin state: N of these pipelines (one per GPU) run side by side.
"""

import os
import queue
import threading
import time
from pathlib import Path

import numpy as np

from . import engine as _engine
from .utils import files, logger

log = logger.get_logger("prob")
_STOP = object()
LAST_STATS = []  # stats dict of every finished BinPipeline.run in this process (benchmarks read it)


class _PinnedPool:
    """Reusable pinned host byte buffers (cudaHostAlloc per bin would cost more than the copy)."""

    def __init__(self, torch):
        self.torch = torch
        self.free = []
        self.lock = threading.Lock()

    def get(self, nbytes):
        with self.lock:
            for i, t in enumerate(self.free):
                if t.numel() >= nbytes:
                    return self.free.pop(i)
        cap = max(1 << 20, int(nbytes * 1.25))
        try:
            return self.torch.empty(cap, dtype=self.torch.uint8, pin_memory=True)
        except RuntimeError:  # no CUDA runtime (host-only unit tests): an ordinary buffer serves the same interface
            return self.torch.empty(cap, dtype=self.torch.uint8)

    def put(self, t):
        with self.lock:
            self.free.append(t)
            self.free.sort(key=lambda x: x.numel())
            del self.free[8:]


class BinPipeline:
    """Processes raw bins on one Engine.  `run(sample_paths)` -> set of processed sample names."""

    def __init__(self, net, classes, out_dir, batch_size=None, force=False, suffix=".prob", loaders=None, writers=None, depth=None,
                 want_labels=False):
        # one B200 takes ~0.27 M ROI/s = 1.6 GB/s of .roi bytes; a loader thread copies 2.8-3.4 GB/s out of the page cache
        # when it has a core to itself, a writer thread formats ~1 M ROI/s.  Four loaders keep the GPU fed with slack for
        # cold files and busy hosts (SYKEPIC_LOADERS / SYKEPIC_WRITERS override).
        loaders = loaders or int(os.environ.get("SYKEPIC_LOADERS", "4"))
        writers = writers or int(os.environ.get("SYKEPIC_WRITERS", "2"))
        depth = depth or 3
        self.net = net
        self.classes = list(classes)
        self.out_dir = out_dir
        self.batch_size = batch_size
        self.force = force
        self.suffix = suffix
        self.n_loaders = max(1, loaders)
        self.n_writers = max(1, writers)
        self.depth = max(1, depth)
        self.want_labels = want_labels
        self.stats = {"bins": 0, "rois": 0, "load_s": 0.0, "gpu_wait_s": 0.0, "write_s": 0.0, "roi_bytes": 0, "csv_bytes": 0}
        self._stat_lock = threading.Lock()

    # ------------------------------------------------------------------ stages
    def _load(self, sample_path, pool):
        """-> dict(sample, csv_path, roi_id, w, h, start, buf (pinned tensor), roi_len) or {"skip": True}."""
        sample_path = Path(sample_path)
        sample = sample_path.name
        csv_path = files.sample_csv_path(sample_path, self.out_dir, suffix=self.suffix)
        if csv_path.is_file():
            if self.force:
                log.warning(f"{csv_path.name} already exists, overwriting")
            else:
                log.warning(f"{csv_path.name} already exists, skipping")
                return {"sample": sample, "skip": True}
        t0 = time.perf_counter()
        n_bytes = sample_path.with_suffix(".roi").stat().st_size
        buf = pool.get(max(n_bytes, 16))
        try:
            try:
                roi_id, w, h, start, roi_len = _engine.load_bin(sample_path, buf, self.net.th, self.net.tw)
            except _engine.CapacityError as e:  # the file grew between stat and read
                pool.put(buf)
                buf = pool.get(e.needed)
                roi_id, w, h, start, roi_len = _engine.load_bin(sample_path, buf, self.net.th, self.net.tw)
        except BaseException:
            pool.put(buf)
            raise
        with self._stat_lock:
            self.stats["load_s"] += time.perf_counter() - t0
            self.stats["roi_bytes"] += roi_len
        return {"sample": sample, "csv_path": csv_path, "roi_id": roi_id, "w": w, "h": h, "start": start, "buf": buf,
                "roi_len": roi_len, "skip": False}

    def _write(self, item):
        t0 = time.perf_counter()
        roi_id, probs = item["roi_id"], item["probs"]
        if len(roi_id) > 1 and np.any(np.diff(roi_id) < 0):  # results sorted by ROI id (probability.py:197)
            order = np.argsort(roi_id, kind="stable")
            roi_id, probs = roi_id[order], probs[order]
        csv_path = Path(item["csv_path"])
        csv_path.parent.mkdir(parents=True, exist_ok=True)
        n_text = _engine.write_prob_csv(csv_path, self.classes, roi_id, probs)
        with self._stat_lock:
            self.stats["write_s"] += time.perf_counter() - t0
            self.stats["csv_bytes"] += n_text
            self.stats["bins"] += 1
            self.stats["rois"] += len(roi_id)

    # ------------------------------------------------------------------ driver
    def run(self, sample_paths, progress=None):
        torch = self.net.torch
        pool = _PinnedPool(torch)
        sample_paths = list(sample_paths)
        processed = set()
        plock = threading.Lock()
        todo = iter(enumerate(sample_paths))
        todo_lock = threading.Lock()
        stop = threading.Event()  # set when the driver loop leaves early: loaders stop taking work
        loaded = {}  # index -> item | exception marker, consumed in order so that bins finish in submission order
        loaded_cv = threading.Condition()
        slots = threading.Semaphore(self.depth + self.n_loaders)  # bounds the pinned memory in flight
        write_q = queue.Queue(maxsize=self.depth + 2)

        def loader():
            while not stop.is_set():
                # the slot is taken BEFORE the index: the lowest outstanding index always holds a slot, so the in-order
                # consumer below can never be starved by later bins occupying every slot
                if not slots.acquire(timeout=0.2):
                    continue
                with todo_lock:
                    nxt = next(todo, None)
                if nxt is None:
                    slots.release()
                    return
                i, sp = nxt
                try:
                    item = self._load(sp, pool)
                except ValueError:
                    log.exception(f"Faulty raw data for {Path(sp).name}")
                    item = None
                except Exception:
                    log.exception(f"Unexpected error for {Path(sp).name}:")
                    item = None
                if item is None or item.get("skip"):
                    slots.release()
                with loaded_cv:
                    loaded[i] = item
                    loaded_cv.notify_all()

        def writer():
            while True:
                item = write_q.get()
                if item is _STOP:
                    return
                try:
                    self._write(item)
                    with plock:
                        processed.add(item["sample"])
                except Exception:
                    log.exception(f"Unexpected error for {item['sample']}:")
                if progress is not None:
                    progress.update(1)

        loaders = [threading.Thread(target=loader, name=f"spk-load{k}", daemon=True) for k in range(self.n_loaders)]
        writers = [threading.Thread(target=writer, name=f"spk-write{k}", daemon=True) for k in range(self.n_writers)]
        t_run = time.perf_counter()  # before the first file is opened
        for t in loaders + writers:
            t.start()

        def finish(h):
            """Await a submitted bin, release its input buffer, hand the results to the writers."""
            item, handle = h
            t0 = time.perf_counter()
            try:
                out = handle.result()
            except ValueError:
                log.exception(f"Faulty raw data for {item['sample']}")
                out = None
            except Exception:
                log.exception(f"Unexpected error for {item['sample']}:")
                out = None
            with self._stat_lock:
                self.stats["gpu_wait_s"] += time.perf_counter() - t0
            pool.put(item.pop("buf"))
            slots.release()
            if out is None:
                if progress is not None:
                    progress.update(1)
                return
            item["probs"] = out[0] if isinstance(out, tuple) else out
            if isinstance(out, tuple):
                item["label"], item["classified"] = out[1], out[2]
            write_q.put(item)

        pending = None
        try:
            for i in range(len(sample_paths)):
                with loaded_cv:
                    while i not in loaded:
                        loaded_cv.wait()
                    item = loaded.pop(i)
                if item is None:
                    if progress is not None:
                        progress.update(1)
                    continue
                if item.get("skip"):  # existing CSV, not forced: counted as processed, like the reference (:141)
                    with plock:
                        processed.add(item["sample"])
                    if progress is not None:
                        progress.update(1)
                    continue
                try:
                    handle = self.net.submit_rois(item["w"], item["h"], item["start"], item["buf"], item["roi_len"],
                                                  batch_size=self.batch_size, want_labels=self.want_labels)
                except Exception:
                    log.exception(f"Unexpected error for {item['sample']}:")
                    pool.put(item.pop("buf"))
                    slots.release()
                    if progress is not None:
                        progress.update(1)
                    continue
                if pending is not None:
                    finish(pending)
                pending = (item, handle)
            if pending is not None:
                finish(pending)
        finally:
            stop.set()
            for _ in writers:
                write_q.put(_STOP)
            for t in loaders + writers:
                t.join()
            self.stats["run_s"] = time.perf_counter() - t_run
            LAST_STATS.append(dict(self.stats))
            del LAST_STATS[:-64]
        return processed
