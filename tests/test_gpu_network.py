"""K2 + K3 (convolutions, head, softmax, thresholds) and the whole `sykepic prob` path on the GPU,
against what the REFERENCE produced on the same bins and checkpoints (tests/golden).

Tolerances (BASELINE.json north_star): FP32 probabilities within 1e-4 absolute, BF16 within 2e-2,
thresholded labels agree on >= 99.9 % of ROIs (FP32)."""

import json

import numpy as np
import pytest

from oracle import prediction as o_pred
from sykepic_b200 import engine, synth
from sykepic_b200.compute import prediction, probability
from tests.cases import CASES, FIXTURE, GOLDEN, case_bins

pytestmark = pytest.mark.gpu

FP32_PROB_TOL = 1e-4
BF16_PROB_TOL = 2e-2


@pytest.fixture(scope="module")
def engines(model_dirs):
    cache = {}

    def get(case, precision, impl="auto", max_batch=64):
        key = (case, precision, impl, max_batch)
        if key not in cache:
            cache[key] = engine.Engine(model_dirs(case), precision=precision, conv_impl=impl, max_batch=max_batch)
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


def _bf16_tol(golden_logits):
    """Per-ROI BF16 gate.  2e-2 absolute (north_star) for ROIs whose logits are in the range a trained
    checkpoint produces: the reference's only real output, tests/data/prob/*.prob.csv, has a per-ROI logit
    standard deviation of 7-9 (max probability 0.46 / 0.23).  The synthetic stress checkpoints (logit gain 24,
    uniform-noise ROIs) reach 3-6x that; a bf16 mantissa gives a RELATIVE logit error, so there the gate
    scales with the ROI's logit spread."""
    spread = golden_logits.std(axis=1)
    return BF16_PROB_TOL * np.maximum(1.0, spread / 10.0)


def _thresholds(tname):
    return 0.5 if tname.startswith("scalar") else o_pred.threshold_dictionary(FIXTURE / f"{tname}.txt")


@pytest.mark.parametrize("case", list(CASES))
def test_fp32_probabilities_and_labels(engines, case):
    eng = engines(case, "fp32")
    labels = json.loads((GOLDEN / f"case_{case}.labels.json").read_text())
    total = agree = 0
    for bname, b in case_bins(case):
        g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
        for tname in ("thresholds-2021", "thresholds-zero", "scalar-0.5"):
            eng.set_thresholds(_thresholds(tname))
            rid, probs, label, classified = eng.run_bin(b["adc_text"], b["roi_bytes"], want_labels=True)
            assert rid.tolist() == g["roi_id"].tolist()
            assert np.abs(probs - g["probs"]).max() <= FP32_PROB_TOL
            logits = eng.last_logits(len(rid))
            scale = max(1.0, float(np.abs(g["logits"]).max()))
            assert np.abs(logits - g["logits"]).max() <= 2e-4 * scale
            # the fused label rule == the host rule on the CSV decimals of the same probabilities
            vals = o_pred.parse_prob_csv_text(engine.format_prob_csv(eng.spec.classes, rid, probs).decode())[2]
            idx, flag = prediction.predict_array(vals, eng.spec.classes, _thresholds(tname))
            assert label.tolist() == idx.tolist() and classified.tolist() == flag.tolist()
            # ... and agrees with the labels the reference derived from ITS OWN csv
            want = labels[bname][tname]
            names = [eng.spec.classes[i] for i in label]
            same = [(a == b2 and bool(c) == d) for a, b2, c, d in zip(names, want["prediction"], classified, want["classified"])]
            total += len(same)
            agree += sum(same)
    assert total > 0 and agree / total >= 0.999, (agree, total)


@pytest.mark.parametrize("case", list(CASES))
def test_bf16_probabilities(engines, case):
    eng = engines(case, "bf16")
    eng.set_thresholds(_thresholds("thresholds-zero"))
    labels = json.loads((GOLDEN / f"case_{case}.labels.json").read_text())
    total = agree = clear = 0
    for bname, b in case_bins(case):
        g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
        rid, probs, label, classified = eng.run_bin(b["adc_text"], b["roi_bytes"], want_labels=True)
        assert np.isfinite(probs).all()
        err = np.abs(probs - g["probs"]).max(axis=1)
        tol = _bf16_tol(g["logits"])
        assert (err <= tol).all(), (float(err.max()), float(g["logits"].std(axis=1).max()))
        want = labels[bname]["thresholds-zero"]["prediction"]
        same = np.array([eng.spec.classes[i] == n for i, n in zip(label, want)])
        # a ROI whose reference winner leads the runner-up by more than twice the gate cannot change its label inside the
        # gate: those must all agree.  (Random-init checkpoints leave many near-ties, which a perturbation of either sign
        # flips; they only count towards the overall rate.)
        top2 = np.sort(g["probs"], axis=1)[:, -2:]
        decided = (top2[:, 1] - top2[:, 0]) > 2.0 * tol
        assert same[decided].all(), (bname, np.flatnonzero(decided & ~same).tolist())
        clear += int(decided.sum())
        total += len(want)
        agree += int(same.sum())
    assert agree / total >= 0.90, (agree, total)
    assert clear > 0


@pytest.mark.parametrize("case", ["r18_224n", "r50_224"])
def test_tcgen05_and_cuda_core_paths_agree(engines, case):
    """Whole network, bf16 activations: tensor-core path (bf16 weights, fused stem) against the CUDA-core
    path (fp32 weights).  Both must sit inside the BF16 gate of the reference's probabilities; the tight
    layer-by-layer comparison is tests/test_gpu_conv.py."""
    tc = engines(case, "bf16", "auto")
    simt = engines(case, "bf16", "simt")
    bname, b = case_bins(case)[-1]
    g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
    _, p_tc = tc.run_bin(b["adc_text"], b["roi_bytes"])
    _, p_simt = simt.run_bin(b["adc_text"], b["roi_bytes"])
    tol = _bf16_tol(g["logits"])
    assert (np.abs(p_tc - g["probs"]).max(axis=1) <= tol).all()
    assert (np.abs(p_simt - g["probs"]).max(axis=1) <= tol).all()
    assert np.abs(p_tc - p_simt).max() <= 3e-2


def test_batch_split_and_partial_batches(engines):
    """Batches of 7 (ragged last batch) give the same probabilities as one batch."""
    eng = engines("r18_180", "fp32")
    bname, b = case_bins("r18_180")[2]
    _, p_all = eng.run_bin(b["adc_text"], b["roi_bytes"])
    _, p_7 = eng.run_bin(b["adc_text"], b["roi_bytes"], batch_size=7)
    assert np.abs(p_all - p_7).max() <= 1e-6


def test_prob_main_writes_reference_layout(engines, model_dirs, tmp_path):
    """`probability.main` end to end: same relative path, header, ROI ids and (within 1e-4) values as the
    CSV the reference wrote; existing CSV is skipped unless force; faulty / empty bins behave like the reference."""
    case = "r18_180"
    raw = tmp_path / "raw"
    paths = [synth.write_bin(raw, bname, b) for bname, b in case_bins(case)]
    # a truncated bin ("Faulty raw data": no CSV) and an empty bin (header-only CSV)
    bad = case_bins(case)[1][1]
    bad_path = synth.write_bin(raw, "D20210601T000000_IFCB114", {"adc_text": bad["adc_text"], "roi_bytes": bad["roi_bytes"][:-5]})
    empty_path = synth.write_bin(raw, "D20210601T002000_IFCB114", {"adc_text": "", "roi_bytes": np.zeros(0, np.uint8)})
    out = tmp_path / "out"
    done = probability.main(paths + [bad_path, empty_path], model_dirs(case), out, batch_size=16, num_workers=2,
                            force=False, progress_bar=False, precision="fp32")
    assert done == {p.name for p in paths} | {empty_path.name}
    labels = json.loads((GOLDEN / f"case_{case}.labels.json").read_text())
    for p in paths:
        rel = labels[p.name]["csv_relpath"]
        got = o_pred.parse_prob_csv_text((out / rel).read_text())
        want = o_pred.parse_prob_csv_text((GOLDEN / f"case_{case}__{p.name}.prob.csv").read_text())
        assert got[0] == want[0] and got[1].tolist() == want[1].tolist()
        assert np.abs(got[2] - want[2]).max() <= FP32_PROB_TOL + 1e-5
    assert not list(out.glob(f"**/{bad_path.name}*"))
    assert (out / "2021/06/01" / f"{empty_path.name}.prob.csv").read_text().count("\n") == 1
    # skip / force
    target = out / labels[paths[0].name]["csv_relpath"]
    target.write_text("sentinel")
    probability.main(paths[:1], model_dirs(case), out, 16, 2, False, progress_bar=False, precision="fp32")
    assert target.read_text() == "sentinel"
    probability.main(paths[:1], model_dirs(case), out, 16, 2, True, progress_bar=False, precision="fp32")
    assert target.read_text().startswith("roi,")


def test_cli_prob_then_class(model_dirs, tmp_path):
    from sykepic_b200.__main__ import main as cli

    case = "r18_180"
    raw = tmp_path / "raw"
    for bname, b in case_bins(case):
        synth.write_bin(raw, bname, b)
    out = tmp_path / "prob"
    assert cli(["prob", "-r", str(raw), "-m", str(model_dirs(case)), "-o", str(out), "-b", "32"]) == 0
    assert len(list(out.glob("**/*.prob.csv"))) == 3
    thr = tmp_path / "thr.txt"
    thr.write_text((FIXTURE / "thresholds-zero.txt").read_text() +
                   "Dolichospermum-Anabaenopsis_coiled 0\nNodularia_spumigena-coiled 0\n")
    assert cli(["class", str(out), "-t", str(thr), "-o", str(tmp_path / "class.csv")]) == 0
    lines = (tmp_path / "class.csv").read_text().splitlines()
    assert lines[0].startswith("Time,") and len(lines) == 4


@pytest.mark.parametrize("mode", ["thread", "process"])
def test_prob_on_two_gpus_writes_the_same_files(model_dirs, tmp_path, monkeypatch, mode):
    """`sykepic prob --gpus 2` (probability.main(devices=2): bins sharded by .roi size over the GPUs, one host thread or one
    spawned process per GPU, no collective): every bin processed once, the CSV files byte-identical to the
    one-GPU run.  Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = "r18_180"
    raw = tmp_path / "raw"
    names = []
    for rep in range(3):  # nine bins: the three of the case under three names each
        for bname, b in case_bins(case):
            name = bname[:-3] + f"{100 + rep:03d}"
            synth.write_bin(raw, name, b)
            names.append(name)
    paths = sorted(p.with_suffix("") for p in raw.glob("*.roi"))
    assert len(paths) == 9
    one = probability.main(paths, model_dirs(case), tmp_path / "one", batch_size=64, progress_bar=False, precision="bf16", devices=1)
    monkeypatch.setenv("SYKEPIC_MULTI", mode)
    two = probability.main(paths, model_dirs(case), tmp_path / "two", batch_size=64, progress_bar=False, precision="bf16", devices=2)
    assert set(one) == set(two) and len(two) == 9
    a = sorted((tmp_path / "one").rglob("*.prob.csv"))
    b = sorted((tmp_path / "two").rglob("*.prob.csv"))
    assert [f.relative_to(tmp_path / "one") for f in a] == [f.relative_to(tmp_path / "two") for f in b] and len(a) == 9
    for fa, fb in zip(a, b):
        assert fa.read_bytes() == fb.read_bytes(), fa.name


def test_image_mode_matches_raw_mode(engines, model_dirs, tmp_path):
    """SURVEY 8f rank 1, `sykepic prob --image-dir`: the ROIs of two bins written as `<sample>_<roi>.png` (what
    ifcb.raw_to_png produces, sykepic/utils/ifcb.py:76-118) go through `probability.call` and give the raw mode's
    probabilities (same kernels, same bytes), one CSV per sample directly under OUT (probability.py:96)."""
    from types import SimpleNamespace

    from tests.test_host_png import write_png_up

    case = "r18_180"
    eng = engines(case, "fp32")
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    expected = {}
    for bname, b in case_bins(case)[:2]:  # the reference's fixture bin and the edge-geometry bin
        rid, w, h, start = engine.parse_adc(b["adc_text"])
        for i, ww, hh, s in zip(rid, w, h, start):
            px = np.asarray(b["roi_bytes"][int(s): int(s) + int(ww) * int(hh)]).reshape(int(hh), int(ww))
            write_png_up(img_dir / f"{bname}_{int(i):05d}.png", px)
        rid2, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
        assert np.array_equal(rid2, rid)
        expected[bname] = (rid2, probs)
    out = tmp_path / "out"
    args = SimpleNamespace(image_dir=str(img_dir), images=None, raw=None, samples=None, model=str(model_dirs(case)), out=str(out),
                           batch_size=64, num_workers=2, force=False, precision="fp32", devices=None)
    assert probability.call(args) is None  # the reference returns nothing in image mode (probability.py:94-97)
    assert sorted(p.name for p in out.iterdir()) == sorted(f"{b}.prob.csv" for b in expected)
    for bname, (rid, probs) in expected.items():
        lines = (out / f"{bname}.prob.csv").read_text().splitlines()
        assert lines[0] == "roi," + ",".join(eng.spec.classes)
        got = np.array([[float(x) for x in l.split(",")[1:]] for l in lines[1:]], np.float64)
        assert [int(l.split(",")[0]) for l in lines[1:]] == rid.tolist()
        assert np.abs(got - probs).max() <= 1e-5 + 5e-6  # 5 decimals of the same fp32 values (launch size differs: 1024 vs 64)


def test_densenet121_matches_oracle(tmp_path):
    """BASELINE config 4.  The reference raises for DenseNet at these sizes, so there is no golden
    from it: the defined behaviour (torchvision's forward + the syke-pic head) is checked against the oracle's
    torch-CPU fp32 restatement of it, in FP32 (1e-4) and BF16 (spread-aware 2e-2 gate)."""
    from oracle import ifcb as o_ifcb
    from oracle import pipeline

    mdir = synth.write_model_dir(tmp_path / "model_d121", arch="densenet121", t=224, head=(256, 128), seed=3, border="mode",
                                 imagenet_normalization=False, randomize_bn=True, logit_gain=8.0)
    model = pipeline.prepare_model(mdir)
    b = synth.synth_bin(1011, 20)
    rows = o_ifcb.parse_adc_text(b["adc_text"])
    want = pipeline.net_pass(model, rows, np.asarray(b["roi_bytes"], np.uint8), batch_size=32)
    wp = np.array([p for _, p in want], np.float32)
    for precision in ("fp32", "bf16"):
        eng = engine.Engine(mdir, precision=precision, max_batch=32)
        try:
            rid, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
            assert rid.tolist() == [r for r, _ in want]
            err = np.abs(probs - wp).max(axis=1)
            if precision == "fp32":
                assert err.max() <= FP32_PROB_TOL, float(err.max())
            else:
                logits = np.log(np.maximum(wp, 1e-30)) / np.log(engine.SOFTMAX_EXP)  # up to a per-ROI constant
                assert (err <= _bf16_tol(logits)).all(), float(err.max())
        finally:
            eng.close()


def test_fp32_tc_overflow_is_loud(model_dirs, tmp_path):
    """The fp16 hi + lo split format of FP32_TC holds |x| <= 65504.  A checkpoint whose activations leave that range must
    not produce quiet garbage: the convolution epilogue counts the tile, `.result()` raises; `fp32` runs the same model."""
    import shutil

    import torch

    src = model_dirs("r18_180")
    mdir = tmp_path / "huge"
    shutil.copytree(src, mdir)
    sd = torch.load(mdir / "best_state.pth", map_location="cpu", weights_only=False)
    key = next(k for k in sd if k.endswith("weight") and sd[k].ndim == 4 and sd[k].shape[1] in (1, 3))  # the stem convolution
    sd[key] = sd[key] * 3.0e6
    torch.save(sd, mdir / "best_state.pth")
    (bname, b), = case_bins("r18_180")[:1]
    eng = engine.Engine(mdir, precision="fp32_tc", max_batch=64)
    try:
        with pytest.raises(ArithmeticError, match="fp16 range"):
            eng.run_bin(b["adc_text"], b["roi_bytes"])
    finally:
        eng.close()
    eng = engine.Engine(mdir, precision="fp32", max_batch=64)
    try:
        rid, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
        assert np.isfinite(probs).all()
    finally:
        eng.close()


@pytest.mark.parametrize("case", ["r18_180", "r18_224n", "r50_224"])
def test_fp32_tc_probabilities(engines, case):
    """FP32-level accuracy on the bf16 tensor cores (precision "fp32_tc": bf16 hi + lo activations and weights, fp32
    accumulation in chunks of 4 k blocks summed in registers): within the FP32 gate of 1e-4 on ResNet-18 AND ResNet-50,
    labels identical to the fused rule on its own CSV decimals."""
    eng = engines(case, "fp32_tc")
    eng.set_thresholds(_thresholds("thresholds-zero"))
    for bname, b in case_bins(case):
        g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
        rid, probs, label, classified = eng.run_bin(b["adc_text"], b["roi_bytes"], want_labels=True)
        assert rid.tolist() == g["roi_id"].tolist()
        assert np.abs(probs - g["probs"]).max() <= FP32_PROB_TOL, float(np.abs(probs - g["probs"]).max())
        vals = o_pred.parse_prob_csv_text(engine.format_prob_csv(eng.spec.classes, rid, probs).decode())[2]
        idx, flag = prediction.predict_array(vals, eng.spec.classes, _thresholds("thresholds-zero"))
        assert label.tolist() == idx.tolist() and classified.tolist() == flag.tolist()
