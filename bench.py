#!/usr/bin/env python
"""Benchmark of the syke-pic prob hot path on B200 (BASELINE.json metric: IFCB ROIs/s, bin -> class probs).

    python bench.py --gpus N --steps K --warmup W            # the B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU arm (oracle port of the reference)

A step = one pass of the hot path (K1 decode/resize/normalise -> K2 CNN -> K3 pool/head/softmax/
threshold) over one batch of synthetic IFCB ROIs.  Default workload = BASELINE.json configs[1]:
ResNet-18, 3x224x224, batch 256 per GPU, BF16, random-init weights (seeded), synthetic ROIs with the
IFCB size distribution (sykepic_b200/synth.py).  Multi-GPU = bins sharded over ranks, no collective
on the data path (weak scaling); NCCL is used only for the barrier and the max-over-ranks time.

Prints ONE JSON line (rank 0).  `value`: device-timed throughput with the batch resident in HBM;
`e2e`: the same through the host API (`Engine.run_rois`: pinned host buffers -> H2D -> kernels -> D2H);
`roofline`: the tcgen05 convolution kernels against the measured BF16 peak; `cpu_baseline`: the oracle
port timed on this box's host cores on a bounded sample.
"""

import argparse
import json
import os
import shutil
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ifcb_rois_per_s"
UNIT = "ROIs/s"
FLUSH_BYTES = 256 << 20

# conv FLOPs / ROI at 224x224 with un-padded logical shapes -- informational
CONV_GFLOP = {"resnet18": 3.627, "resnet50": 8.174, "densenet121": 5.666}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--batch", type=int, default=None, help="ROIs per step per GPU (default 256; 512 for resnet50/densenet121)")
    ap.add_argument("--chunk-batches", type=int, default=16,
                    help="batches per bin chunk: K1 decodes a chunk per launch (as Engine.run_bin_device does), K2+K3 run per batch")
    ap.add_argument("--target", type=int, default=224)
    ap.add_argument("--precision", choices=("bf16", "fp32", "fp32_tc"), default="bf16")
    ap.add_argument("--conv-impl", choices=("auto", "simt", "tcgen05", "taps"), default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-detail", metavar="FILE", help="write the per-launch timing table of the roofline pass")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--e2e-bins", type=int, default=64, help="IFCB bins per GPU of the end-to-end (files -> CSV) leg")
    ap.add_argument("--profile-events", metavar="FILE", help="also write the per-launch CUDA-event table (serialised launches)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload(args, seed, n_chunks, rois=None):
    """n_chunks synthetic bins of `rois` non-empty ROIs (default: one batch) -> list of (w, h, start, roi_bytes)."""
    from sykepic_b200 import synth

    rois = rois or args.batch
    out = []
    for i in range(n_chunks):
        b = synth.synth_bin(seed + i, int(rois * 1.01) + 8)
        keep = np.flatnonzero(b["w"] > 0)[:rois]
        assert len(keep) == rois
        out.append((b["w"][keep].astype(np.int32), b["h"][keep].astype(np.int32), b["start"][keep].astype(np.int64), b["roi_bytes"]))
    return out


BENCH_CASE = {"resnet18": "bench_r18", "resnet50": "bench_r50", "densenet121": "bench_d121"}


def model_dir(args, root):
    """The checkpoint of the parity case of this architecture (tests/cases.py BIG_CASES: seeded random-init weights,
    BatchNorm statistics calibrated on synthetic ROIs, logit spread of a trained checkpoint), so that the network
    benchmarked here is the one tests/test_gpu_bench_parity.py holds against the reference's own output."""
    from sykepic_b200 import synth

    if args.arch in BENCH_CASE and args.target == 224:
        from tests.cases import case_model_dir

        return case_model_dir(BENCH_CASE[args.arch], root)
    return synth.write_model_dir(Path(root) / f"model_{args.arch}", arch=args.arch, t=args.target, head=(256, 128), seed=0,
                                 border="mode", imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                break
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(args, mdir, batches, budget_s, max_rois=4096, chunk=64):
    """The oracle (numpy + torch-CPU restatement of the reference path) on this box's host cores."""
    import torch

    from oracle import network, pipeline

    model = pipeline.prepare_model(mdir)
    w, h, start, roi = batches[0]
    done, t_used = 0, 0.0
    i = 0
    # one untimed chunk to page everything in
    with torch.no_grad():
        while done < max_rois and (t_used < budget_s or done == 0):
            lo = (i * chunk) % len(w)
            idx = range(lo, min(lo + chunk, len(w)))
            t0 = time.perf_counter()
            imgs = [roi[start[k]:start[k] + int(w[k]) * int(h[k])].reshape(int(h[k]), int(w[k])) for k in idx]
            x = torch.from_numpy(pipeline.preprocess_rois(model, imgs))
            probs = network.probabilities(network.forward_logits(model.state_dict, x))
            probs.tolist()
            dt = time.perf_counter() - t0
            if i > 0:  # chunk 0 is warm-up
                done += len(imgs)
                t_used += dt
            i += 1
    return done / t_used, done, torch.get_num_threads()


def parity_check(args, mdir, chunk, lo, probs, label, n=64):
    """The outputs of the last timed step against the oracle (checker only): `n` ROIs spread evenly over that batch.
    The full gates (every ROI of 1200-ROI bins against the reference's own output) are tests/test_gpu_bench_parity.py."""
    import torch

    from oracle import network, pipeline

    model = pipeline.prepare_model(mdir)
    w, h, start, roi = chunk
    pick = np.unique(np.linspace(0, len(probs) - 1, n).astype(int))
    imgs = [roi[start[lo + k]:start[lo + k] + int(w[lo + k]) * int(h[lo + k])].reshape(int(h[lo + k]), int(w[lo + k])) for k in pick]
    want = network.probabilities(network.forward_logits(model.state_dict, torch.from_numpy(pipeline.preprocess_rois(model, imgs)))).numpy()
    err = np.abs(probs[pick] - want)
    tol = 2e-2 if args.precision == "bf16" else 1e-4
    top2 = np.sort(want, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 2 * tol
    same = label[pick] == want.argmax(axis=1)  # thresholds 0.5 everywhere: the label is the arg max either way
    return {"max_dp": float(err.max()), "mean_dp": float(err.mean()), "n": int(len(pick)), "tol": tol, "ok": bool(err.max() <= tol),
            "argmax_agree": float(same.mean()), "argmax_agree_decided": float(same[decided].mean()) if decided.any() else None,
            "against": "oracle (numpy cv2-exact transform + torch-CPU fp32 forward) on ROIs spread over the last timed batch"}


def _import_reference():
    """The UNMODIFIED reference from baseline/_ref (installed by baseline/install_ref.sh; git-ignored, travels to the GPU
    box) -> its `sykepic.compute.probability` module, or None.  Shim of SURVEY 8c: `sykepic/utils/ifcb.py:9` imports an
    unused `pytz`, absent from this image -- a dummy module is injected after pandas has been imported."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "sykepic" / "compute" / "probability.py").exists():
        return None
    import types

    import pandas  # noqa: F401  (before the stub: pandas probes pytz itself)

    sys.modules.setdefault("pytz", types.ModuleType("pytz")).timezone = lambda name: None
    sys.path.insert(0, str(ref))
    try:
        from sykepic.compute import probability as ref_probability
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: reference import failed ({type(e).__name__}: {e}); falling back to the oracle port", file=sys.stderr)
        sys.path.remove(str(ref))
        return None
    return ref_probability


def run_reference(args):
    """CPU arm.  kind "reference": the reference's own `probability.main` (baseline/_ref) on bins written to local
    disk -- .adc/.roi -> PNG round trip -> DataLoader workers -> torch-CPU fp32 forward -> CSV, its stock code path --
    one bin of `batch` ROIs per step, all host cores (best effort: batch_size = batch, num_workers = min(16, cores)).
    kind "port" (only when baseline/_ref is absent): the oracle restatement."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference picks cuda:0 when it sees one (probability.py:127); this is the CPU arm
    import torch

    # all the host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)
    try:
        cores = max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        cores = max(1, os.cpu_count() or 1)
    torch.set_num_threads(cores)
    args.batch = args.batch or (256 if args.arch == "resnet18" else 512)
    tmp = tempfile.mkdtemp(prefix="spk_bench_ref_")
    mdir = model_dir(args, tmp)
    ref_probability = _import_reference()
    config = {"workload": f"{args.arch} 3x{args.target}x{args.target} synthetic IFCB ROIs, batch {args.batch}",
              "arch": args.arch, "batch": args.batch, "target": args.target}
    if ref_probability is not None:
        from sykepic_b200 import synth

        # a step = one bin of `batch` ROIs; the CPU reference does ~100-250 ROIs/s, so the run is bounded by a time budget
        sample = min(args.batch, 256)
        raw = Path(tmp) / "raw"
        n_bins = max(args.warmup, 1) + max(args.steps, 1)
        paths = []
        for i in range(n_bins):
            b = synth.synth_bin(2000 + i, int(sample * 1.01) + 8)
            keep = np.flatnonzero(b["w"] > 0)[:sample]
            area = b["w"][keep].astype(np.int64) * b["h"][keep]
            start = np.concatenate([[0], np.cumsum(area)[:-1]]).astype(np.int64)
            roi = np.concatenate([b["roi_bytes"][s:s + a] for s, a in zip(b["start"][keep], area)])
            paths.append(synth.write_bin(raw, synth.bin_name(i), {"adc_text": synth.adc_text(b["w"][keep], b["h"][keep], start), "roi_bytes": roi}))
        workers = min(16, cores)
        out = Path(tmp) / "out"
        nw = max(args.warmup, 1)
        ref_probability.main(paths[:nw], mdir, out, batch_size=args.batch, num_workers=workers, force=True, progress_bar=False)
        budget = 150.0
        done = 0
        t0 = time.perf_counter()
        # one `main` call per slice of bins so that the time budget can end the run; model construction is inside (as in the CLI)
        for i in range(nw, n_bins):
            got = ref_probability.main(paths[i:i + 1], mdir, out, batch_size=args.batch, num_workers=workers, force=True, progress_bar=False)
            assert got == {paths[i].name}, got
            done += 1
            if time.perf_counter() - t0 > budget:
                break
        dt = time.perf_counter() - t0
        n_csv = len(list(out.glob("**/*.prob.csv")))
        assert n_csv >= done
        kind = "reference"
        note = ("CPU arm: the reference's own sykepic.compute.probability.main from baseline/_ref (unmodified; stock path: PNG round "
                f"trip, DataLoader with {workers} workers, torch-CPU fp32, CSV), one bin of {sample} ROIs per step, model built per call")
        sample_txt = f"{done} steps x one bin of {sample} ROIs through probability.main (files -> CSV)"
    else:
        from oracle import network, pipeline

        batches = workload(args, 2000, 1)
        model = pipeline.prepare_model(mdir)
        w, h, start, roi = batches[0]
        sample = min(64, args.batch)

        def step(i):
            lo = (i * sample) % (len(w) - sample + 1)
            imgs = [roi[start[k]:start[k] + int(w[k]) * int(h[k])].reshape(int(h[k]), int(w[k])) for k in range(lo, lo + sample)]
            x = torch.from_numpy(pipeline.preprocess_rois(model, imgs))
            with torch.no_grad():
                network.probabilities(network.forward_logits(model.state_dict, x)).tolist()

        for i in range(max(args.warmup, 1) if args.steps > 0 else 0):
            step(i)
        t0 = time.perf_counter()
        done = 0
        for i in range(args.steps):
            step(i)
            done += 1
            if time.perf_counter() - t0 > 150.0:
                break
        dt = time.perf_counter() - t0
        kind = "port"
        note = "CPU arm: oracle port of the reference path (baseline/_ref is absent): numpy cv2-exact transform + torch-CPU fp32 forward"
        sample_txt = f"{done} steps x {sample} ROIs of the batch"
    value = done * sample / dt
    config["note"] = note
    config["rois_per_step"] = sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_txt, "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    return 0


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from sykepic_b200 import _lib, engine

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.batch = args.batch or (256 if args.arch == "resnet18" else 512)
    dev = torch.device("cuda", local)

    tmp = tempfile.mkdtemp(prefix=f"spk_bench_r{rank}_")
    mdir = model_dir(args, tmp)
    G = max(1, args.chunk_batches)
    chunk = G * args.batch
    eng = engine.Engine(mdir, device=local, precision=args.precision, max_batch=args.batch, conv_impl=args.conv_impl, pre_chunk=chunk)
    thr = {name: 0.5 for name in eng.spec.classes}
    eng.set_thresholds(thr)
    n_pool = 3
    # synthetic bins of `chunk` ROIs: the engine decodes a whole chunk of a bin per K1 launch and walks it in batches
    chunks = workload(args, 2000 + 100 * rank, n_pool, chunk)
    in_bytes = [int((b[0].astype(np.int64) * b[1]).sum()) for b in chunks]
    K = eng.k

    # ---- device-resident inputs (value) and pinned host inputs (e2e)
    stream = eng.stream
    dev_in = []
    with torch.cuda.stream(stream):
        for w, h, start, roi in chunks:
            dev_in.append((torch.from_numpy(roi).to(dev), torch.from_numpy(start).to(dev), torch.from_numpy(w).to(dev),
                           torch.from_numpy(h).to(dev), len(roi)))
        probs = torch.empty((chunk, K), dtype=torch.float32, device=dev)
        label = torch.empty(chunk, dtype=torch.int32, device=dev)
        cls = torch.empty(chunk, dtype=torch.uint8, device=dev)
        flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream.synchronize()

    def step_device(i):
        """One batch through K2 + K3; every G-th step first runs K1 on the next chunk of G batches (all of K1's work
        for the ROIs of the timed steps is inside the timed region)."""
        j = i % G
        if j == 0:
            roi_d, start_d, w_d, h_d, roi_len = dev_in[(i // G) % n_pool]
            eng.preprocess(roi_d, roi_len, start_d, w_d, h_d, chunk, eng._x)
        lo = j * args.batch
        eng.forward(eng._x[lo:lo + args.batch], args.batch, probs[lo:lo + args.batch], label[lo:lo + args.batch], cls[lo:lo + args.batch])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_l2=True):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(stream):
            for i in range(steps):
                if flush_l2:
                    flush.zero_()  # evict L2 between steps, outside the timed events
                evs[i][0].record(stream)
                fn(i)
                evs[i][1].record(stream)
        stream.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 1)):
            step_device(i)  # step 0 decodes chunk 0
    launches0 = eng.launches
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ms_total = timed(step_device, args.steps)  # starts at step 0 again: K1 runs ceil(steps / G) times inside
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    gpu_launches = eng.launches - launches0
    assert eng.fault_count() == 0
    # what the LAST timed step produced (checked against the oracle below, outside every timed region)
    last = args.steps - 1
    last_chunk, last_lo = (last // G) % n_pool, (last % G) * args.batch
    last_probs = probs[last_lo:last_lo + args.batch].cpu().numpy()
    last_label = label[last_lo:last_lo + args.batch].cpu().numpy()

    # ---- e2e: the PRODUCT path.  `probability.main` over IFCB bins on local disk (tmpfs): .adc/.roi files -> loader threads
    # (spk_bin_load into pinned buffers) -> H2D -> K1/K2/K3 -> D2H -> writer threads (%.5f CSV files), first file open to last
    # CSV closed, with the engine built before the clock (a service builds it once; the reference rebuilds its model per call).
    from sykepic_b200 import pipeline, shard, synth
    from sykepic_b200.compute import probability

    shard.pin_to_gpu_node(local)
    base = "/dev/shm" if Path("/dev/shm").is_dir() else None
    froot = Path(tempfile.mkdtemp(prefix=f"spk_bench_files_r{rank}_", dir=base))
    raw, out_dir = froot / "raw", froot / "out"
    raw.mkdir()
    n_files = max(8, args.e2e_bins)
    distinct = [synth.synth_bin(3000 + 10 * rank + i) for i in range(min(4, n_files))]  # ~5000 ROIs each, IFCB geometry
    free = shutil.disk_usage(froot).free
    per_bin = max(len(d["roi_bytes"]) for d in distinct) * 1.15
    n_files = int(max(8, min(n_files, 0.5 * free / world / per_bin)))
    files_rois = files_bytes = 0
    for i in range(n_files):
        d = distinct[i % len(distinct)]
        name = synth.bin_name(72 * rank + i)
        if i < len(distinct):
            synth.write_bin(raw, name, d)
        else:  # same content under another name: a different file to open, read and answer
            first = synth.bin_name(72 * rank + i % len(distinct))
            shutil.copyfile(raw / f"{first}.adc", raw / f"{name}.adc")
            shutil.copyfile(raw / f"{first}.roi", raw / f"{name}.roi")
        files_rois += int((d["w"] > 0).sum())
        files_bytes += len(d["roi_bytes"])
    paths = sorted(q.with_suffix("") for q in raw.glob("*.roi"))
    warm = probability.main(paths[:4], mdir, froot / "warm", batch_size=args.batch, force=True, progress_bar=False,
                            precision=args.precision, engine=eng)
    assert len(warm) == 4
    barrier()
    t0 = time.perf_counter()
    done = probability.main(paths, mdir, out_dir, batch_size=args.batch, force=True, progress_bar=False, precision=args.precision,
                            engine=eng)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert len(done) == len(paths), (len(done), len(paths))
    stage = dict(pipeline.LAST_STATS[-1])
    n_csv = len(list(out_dir.rglob("*.prob.csv")))
    assert n_csv == len(paths)
    barrier()
    h2d = int(files_bytes / files_rois * args.batch) + 16 * args.batch
    d2h = args.batch * K * 4
    shutil.rmtree(froot, ignore_errors=True)

    # ---- roofline pass: IN-STEP time of every launch from in-kernel global-timer stamps (launches stay back to back, PDL on)
    prof_steps = G  # one full chunk cycle: K1 once, K2 + K3 G times
    eng.profile_begin(stamps=True)
    with torch.cuda.stream(stream):
        for i in range(prof_steps):
            flush.zero_()
            step_device(i)
    prof = eng.profile_read(detail=True)
    eng.profile_end()
    detail = prof.pop("detail")
    if args.profile_detail and rank == 0:
        agg = {}
        for cat, ms, fl, by, what in detail:
            what, _, span = what.partition(" |span_ms=")
            a = agg.setdefault((cat, what), [0, 0.0, fl, by, 0.0])
            a[0] += 1
            a[1] += ms
            a[4] += float(span or 0.0)
        with open(args.profile_detail, "w") as fh:
            fh.write("category\tlaunches\tin_step_ms\tTFLOP/s\tGB/s(algorithmic)\tspan_ms(first CTA start -> last CTA end)\twhat\n")
            for (cat, what), (cnt, ms, fl, by, span) in agg.items():
                avg = max(ms / cnt, 1e-9)
                fh.write(f"{cat}\t{cnt}\t{avg:.4f}\t{fl / avg / 1e9:.1f}\t{by / avg / 1e6:.1f}\t{span / cnt:.4f}\t{what}\n")
    if args.profile_events and rank == 0:  # the old per-launch CUDA-event table (serialises the launches; for comparison)
        eng.profile_begin(stamps=False)
        with torch.cuda.stream(stream):
            for i in range(prof_steps):
                flush.zero_()
                step_device(i)
        ev = eng.profile_read(detail=True)
        eng.profile_end()
        agg = {}
        for cat, ms, fl, by, what in ev.pop("detail"):
            a = agg.setdefault((cat, what), [0, 0.0, fl, by])
            a[0] += 1
            a[1] += ms
        with open(args.profile_events, "w") as fh:
            fh.write("category\tlaunches\tavg_ms\tTFLOP/s\tGB/s(algorithmic)\twhat\n")
            for (cat, what), (cnt, ms, fl, by) in agg.items():
                fh.write(f"{cat}\t{cnt}\t{ms / cnt:.4f}\t{fl / (ms / cnt) / 1e9:.1f}\t{by / (ms / cnt) / 1e6:.1f}\t{what}\n")

    # ---- K1 alone (a kernel timed in isolation: CUDA events on its stream, L2 flushed before each launch)
    k1_ms = []
    with torch.cuda.stream(stream):
        for r in range(6):
            roi_d, start_d, w_d, h_d, roi_len = dev_in[r % n_pool]
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.preprocess(roi_d, roi_len, start_d, w_d, h_d, chunk, eng._x)
            e1.record(stream)
            k1_ms.append((e0, e1, in_bytes[r % n_pool]))
    stream.synchronize()
    k1 = sorted((a.elapsed_time(b), nb) for a, b, nb in k1_ms[1:])[len(k1_ms[1:]) // 2]

    # ---- reduce over ranks: max time, summed ROIs
    t = torch.tensor([ms_total, e2e_s, wall, -float(files_rois)], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(files_rois), float(len(paths))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_s, wall, _ = t.tolist()
    all_rois, all_bins = tot.tolist()

    if rank == 0:
        pk, pk_kind = peaks()
        step_ms = {c: v["ms"] / prof_steps for c, v in prof.items()}
        total_prof_ms = sum(step_ms.values())
        tc = prof.get("conv_tc")
        mean_in = float(np.mean(in_bytes))
        burst, sustained = float(pk["bf16_tflops"]), float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
        # which regime did THIS run see?  The sustained figure belongs to seconds of tensor load under the power cap
        # (MEASURED_PEAKS: 1312 MHz); a run whose SM clock stayed near the maximum is held to the burst figure.
        sm, sm_max = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
        regime = "sustained" if (sm and sm_max and sm < 0.8 * sm_max) else "burst"
        peak = sustained if regime == "sustained" else burst
        traffic_file = ROOT / "profiles" / "r2_traffic.json"
        traffic = json.loads(traffic_file.read_text()) if traffic_file.exists() else {}
        if tc:
            ach = tc["flops"] / (tc["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolutions, all launches of a step (conv3x3_hp / conv_pair / conv_tc kernels)",
                        "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "peak_kind": f"{pk_kind} bf16 {regime} (cuBLAS); SM clock median {sm} of {sm_max} MHz during the timed region",
                        "frac_of_burst": ach / burst, "frac_of_sustained": ach / sustained, "traffic": None,
                        "launches_per_step": tc["launches"] // prof_steps, "share_of_step": step_ms["conv_tc"] / total_prof_ms,
                        "timing": "in-step time per launch from in-kernel global-timer stamps (end - max(previous end, own start)); "
                                  "launches back to back with programmatic dependent launch, as in the timed region",
                        "in_step_ms_sum": total_prof_ms}
            dom = [(ms, fl) for cat, ms, fl, by, what in detail if "[halo pair]" in what and "64->64" in what]
            if dom and args.arch == "resnet18":
                d_ms, d_fl = sum(m for m, _ in dom), sum(f for _, f in dom)
                tr = traffic.get("conv3x3_hp_kernel<64>")
                roofline["dominant"] = {"kernel": "conv3x3_hp_kernel<64> (3x3 64->64 @56x56, 4 launches per step)",
                                        "achieved": d_fl / (d_ms * 1e-3) / 1e12, "frac": d_fl / (d_ms * 1e-3) / 1e12 / peak,
                                        "frac_of_burst": d_fl / (d_ms * 1e-3) / 1e12 / burst,
                                        "ms_per_launch": d_ms / len(dom), "share_of_step": d_ms / prof_steps / total_prof_ms,
                                        "traffic": tr["dram_bytes_per_launch"] if tr else None,
                                        "algorithmic_bytes": tr["algorithmic_bytes_per_launch"] if tr else None,
                                        "traffic_source": tr["source"] if tr else None}
                roofline["traffic"] = roofline["dominant"]["traffic"]
        else:
            cs = prof.get("conv_simt", {"flops": 0.0, "ms": 1.0, "launches": 0})
            ach = cs["flops"] / (cs["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "conv_simt_kernel (CUDA-core FFMA; no tensor-core path in this precision)",
                        "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst, "peak_kind": pk_kind, "traffic": None}
        value = args.gpus * args.batch * args.steps / (ms_total * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.arch} 3x{args.target}x{args.target} synthetic IFCB ROIs, batch {args.batch} per GPU, "
                                   f"{args.precision}, random-init weights",
                       "arch": args.arch, "batch_per_gpu": args.batch, "target": args.target, "conv_impl": args.conv_impl,
                       "parallelism": f"bins sharded over {args.gpus} GPU(s), no collective",
                       "l2": "flushed between steps (256 MiB memset outside the timed events)",
                       "mean_roi_bytes": mean_in / chunk,
                       "k1_chunk": f"K1 decodes {chunk} ROIs ({G} batches) per launch, every {G}th step; K2+K3 per batch",
                       "checkpoint": f"tests/cases.py {BENCH_CASE.get(args.arch, 'seeded random init')}",
                       "conv_gflop_per_roi_logical": CONV_GFLOP.get(args.arch) if args.target == 224 else None},
            "e2e": {"value": all_rois / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": f"probability.main(paths, model_dir, out, batch_size={args.batch}, engine=<prebuilt>) per GPU process: "
                           f"{int(all_bins)} IFCB bins ({int(all_rois)} ROIs) as .adc/.roi files on {base or 'tmp'} -> .prob.csv files; "
                           "first file open to last CSV closed, max over ranks",
                    "bins": int(all_bins), "rois": int(all_rois), "seconds": e2e_s,
                    "stage_seconds_rank0": {k: round(float(stage.get(k, 0.0)), 4) for k in ("load_s", "gpu_wait_s", "write_s", "run_s")},
                    "host_threads_rank0": {"loaders": int(os.environ.get("SYKEPIC_LOADERS", "4")), "writers": int(os.environ.get("SYKEPIC_WRITERS", "2"))}},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_preprocess": None,
            "kernel_ms_per_step": step_ms,
            "wall_s_timed_region": wall,
        }
        k1_bytes = k1[1] + chunk * (16 + args.target * args.target)  # ROI bytes read + descriptors + the u8 planes written
        k1_gbs = k1_bytes / (k1[0] * 1e-3) / 1e9
        pre_tr = traffic.get("preprocess_u8_kernel") or {}
        line["roofline_preprocess"] = {"bound": "hbm", "kernel": "preprocess_u8_kernel (+ classify / heavy-tail cluster kernel)", "achieved": k1_gbs,
                                       "peak": float(pk["hbm_gbs"]), "unit": "GB/s", "frac": k1_gbs / float(pk["hbm_gbs"]),
                                       "bytes_per_roi": k1_bytes / chunk, "rois_per_launch": chunk, "ms_per_launch": k1[0],
                                       "traffic": pre_tr.get("dram_bytes_per_launch"), "traffic_source": pre_tr.get("source"),
                                       "timing": "CUDA events around one spk_preprocess call, L2 flushed, median of 5"}
        line["parity"] = parity_check(args, mdir, chunks[last_chunk], last_lo, last_probs, last_label)
        if not args.no_cpu_baseline and world == 1:  # reported at N = 1 only (under torchrun OMP_NUM_THREADS is 1)
            rate, n_done, cores = cpu_port_rate(args, mdir, chunks, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_done} ROIs of the same synthetic batch (oracle: numpy transform + torch-CPU fp32 forward)",
                                    "host_cpus": os.cpu_count()}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
