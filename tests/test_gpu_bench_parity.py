"""Parity at the BENCHMARKED configurations and launch sizes (BASELINE.json configs 2-4).

tests/cases.py BIG_CASES: the checkpoints bench.py runs (ResNet-18 / ResNet-50 / DenseNet-121 at 3x224x224, logit
spread of a trained checkpoint), one synthetic IFCB bin of 1200 ROIs each.  Goldens: the REFERENCE's own output
(`probability.main` + `prediction_dataframe`, tests/golden/make_golden.py) for the ResNets, the oracle's torchvision
restatement for DenseNet-121 (the reference raises there).

Gates (north_star), with NO spread scaling: BF16 probabilities within 2e-2 absolute on EVERY ROI; FP32 within 1e-4;
thresholded labels agree on >= 99.9 % of the ROIs whose reference winner leads by more than twice the gate (a ROI
inside that margin may legitimately flip under any perturbation allowed by the gate; the overall rate is printed).
Launches of 256 and of 1024 ROIs (multi-wave persistent scheduling, ragged tail 1200 = 1024 + 176 = 4 x 256 + 176)
must give bit-identical results: per-ROI arithmetic does not depend on the launch size.
"""

import json

import numpy as np
import pytest

from oracle import prediction as o_pred
from sykepic_b200 import engine
from tests.cases import BIG_CASES, FIXTURE, GOLDEN, case_bins

pytestmark = pytest.mark.gpu

FP32_PROB_TOL = 1e-4
BF16_PROB_TOL = 2e-2


def _golden(case):
    (bname, b), = case_bins(case)
    g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
    labels = json.loads((GOLDEN / f"case_{case}.labels.json").read_text())[bname]
    return b, g, labels


def _run(model_dir, b, precision, max_batch):
    eng = engine.Engine(model_dir, precision=precision, max_batch=max_batch, pre_chunk=4096)
    try:
        eng.set_thresholds(o_pred.threshold_dictionary(FIXTURE / "thresholds-zero.txt"))
        rid, probs, label, classified = eng.run_bin(b["adc_text"], b["roi_bytes"], batch_size=max_batch, want_labels=True)
        assert eng.fault_count() == 0
        return rid, probs, [eng.spec.classes[i] for i in label], classified
    finally:
        eng.close()


def _label_rates(names, want, golden_probs, tol):
    same = np.array([a == b for a, b in zip(names, want)])
    top2 = np.sort(golden_probs, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 2.0 * tol
    return same, decided


@pytest.mark.parametrize("case", list(BIG_CASES))
def test_bf16_strict_gate_at_launch_sizes(model_dirs, case):
    b, g, labels = _golden(case)
    assert len(g["roi_id"]) >= 1100
    out = {}
    for max_batch in (256, 1024):
        rid, probs, names, classified = _run(model_dirs(case), b, "bf16", max_batch)
        assert rid.tolist() == g["roi_id"].tolist()
        assert np.isfinite(probs).all()
        err = np.abs(probs - g["probs"]).max(axis=1)
        worst = int(err.argmax())
        assert err.max() <= BF16_PROB_TOL, (case, max_batch, float(err.max()), int(rid[worst]), float(g["logits"][worst].std()))
        same, decided = _label_rates(names, labels["thresholds-zero"]["prediction"], g["probs"], BF16_PROB_TOL)
        assert decided.sum() >= 100  # the checkpoint is peaky enough for the label gate to mean something
        assert same[decided].mean() >= 0.999, (case, max_batch, np.flatnonzero(decided & ~same).tolist())
        print(f"{case} bf16 launch {max_batch}: max |dp| {err.max():.2e}, mean {np.abs(probs - g['probs']).mean():.2e}, labels "
              f"{same.mean():.4f} overall, {same[decided].mean():.4f} on the {int(decided.sum())} decided ROIs")
        out[max_batch] = (probs, names, classified)
    assert np.array_equal(out[256][0], out[1024][0])  # bit-identical whatever the launch size
    assert out[256][1] == out[1024][1] and np.array_equal(out[256][2], out[1024][2])


@pytest.mark.parametrize("case", list(BIG_CASES))
def test_fp32_gate_at_launch_size_256(model_dirs, case):
    b, g, labels = _golden(case)
    rid, probs, names, classified = _run(model_dirs(case), b, "fp32", 256)
    assert rid.tolist() == g["roi_id"].tolist()
    err = np.abs(probs - g["probs"]).max(axis=1)
    assert err.max() <= FP32_PROB_TOL, (case, float(err.max()))
    want = labels["thresholds-zero"]
    same = np.array([a == b2 and bool(c) == d for a, b2, c, d in zip(names, want["prediction"], classified, want["classified"])])
    assert same.mean() >= 0.999, (case, float(same.mean()))
    print(f"{case} fp32: max |dp| {err.max():.2e}, labels {same.mean():.4f}")


@pytest.mark.parametrize("case", list(BIG_CASES))
def test_fp32_tc_gate_at_launch_size_256(model_dirs, case):
    """FP32-level accuracy on the tensor cores (precision "fp32_tc", the CLI default): the FP32 gate of 1e-4 and the label
    gate on the benchmark checkpoints, every layer on tcgen05 (no CUDA-core fallback)."""
    b, g, labels = _golden(case)
    eng = engine.Engine(model_dirs(case), precision="fp32_tc", max_batch=256, pre_chunk=4096)
    try:
        assert int(eng.lib.spk_net_simt_layers(eng.ctx)) == 0
        eng.set_thresholds(o_pred.threshold_dictionary(FIXTURE / "thresholds-zero.txt"))
        rid, probs, label, classified = eng.run_bin(b["adc_text"], b["roi_bytes"], batch_size=256, want_labels=True)
        names = [eng.spec.classes[i] for i in label]
    finally:
        eng.close()
    err = np.abs(probs - g["probs"]).max(axis=1)
    print(f"{case} fp32_tc: max |dp| {err.max():.2e}")
    assert err.max() <= FP32_PROB_TOL, (case, float(err.max()))
    want = labels["thresholds-zero"]
    same = np.array([a == b2 and bool(c) == d for a, b2, c, d in zip(names, want["prediction"], classified, want["classified"])])
    assert same.mean() >= 0.999, (case, float(same.mean()))
