"""the reference's downstream tools (`sykepic class`, abundance, class_stats, features_per_prediction,
analyze/evaluation.py:60, analyze/frequency.py:125) all read `.prob.csv` files through `prediction_dataframe`
(sykepic/compute/prediction.py:8-28).  They keep working on files written here because

  1. the bytes are identical: the probabilities the REFERENCE computed (tests/golden, 1200 ROIs x 50 classes per case),
     written by this repo's writer (`spk_prob_csv_write`), give exactly the file the reference wrote; and
  2. the unmodified reference readers (baseline/_ref, installed by baseline/install_ref.sh; skipped when absent), run on a
     file written here, return the frames / counts this repo's mirrors return.
"""
import json
import sys
import types

import numpy as np
import pytest

from sykepic_b200 import engine
from sykepic_b200.compute import classification, prediction
from tests.cases import BIG_CASES, FIXTURE, GOLDEN, ROOT_DIR, case_bins


def _golden(case):
    (bname, _), = case_bins(case)
    return bname, np.load(GOLDEN / f"case_{case}__{bname}.npz"), (GOLDEN / f"case_{case}__{bname}.prob.csv").read_bytes()


@pytest.mark.parametrize("case", list(BIG_CASES))
def test_written_csv_is_byte_identical_to_the_references(tmp_path, case):
    bname, g, want = _golden(case)
    classes = (FIXTURE / "class_names.txt").read_text().splitlines()
    path = tmp_path / f"{bname}.prob.csv"
    engine.write_prob_csv(path, classes, g["roi_id"], g["probs"])
    assert path.read_bytes() == want


@pytest.fixture(scope="module")
def reference():
    ref = ROOT_DIR / "baseline" / "_ref"
    if not (ref / "sykepic" / "compute" / "prediction.py").exists():
        pytest.skip("baseline/_ref is not installed (bash baseline/install_ref.sh)")
    import pandas  # noqa: F401  (before the stub: pandas probes pytz itself)

    sys.modules.setdefault("pytz", types.ModuleType("pytz")).timezone = lambda name: None
    sys.path.insert(0, str(ref))
    try:
        from sykepic.compute import classification as ref_classification
        from sykepic.compute import prediction as ref_prediction
    finally:
        sys.path.remove(str(ref))
    return ref_prediction, ref_classification


@pytest.mark.parametrize("tname", ["thresholds-2021", "thresholds-zero", "scalar"])
def test_reference_readers_on_a_csv_written_here(tmp_path, reference, tname):
    """A CSV of perturbed probabilities (what a GPU run within the gates produces), written here, read back by the
    reference's `prediction_dataframe` / `class_df_probs_only` and by this repo's mirrors: same labels, flags, counts."""
    ref_prediction, ref_classification = reference
    bname, g, _ = _golden("bench_r18")
    classes = (FIXTURE / "class_names.txt").read_text().splitlines()
    rng = np.random.default_rng(1)
    probs = np.clip(g["probs"] + rng.normal(0, 2e-3, g["probs"].shape).astype(np.float32), 0, 1)
    out = tmp_path / "2021" / "05" / "23"
    out.mkdir(parents=True)
    path = out / f"{bname}.prob.csv"
    engine.write_prob_csv(path, classes, g["roi_id"], probs)
    thr = 0.5 if tname == "scalar" else FIXTURE / f"{tname}.txt"
    want = ref_prediction.prediction_dataframe(path, thr)
    got = prediction.prediction_dataframe(path, thr)
    assert list(want.columns) == list(got.columns) and want.index.tolist() == got.index.tolist()
    assert [str(p) for p in want["prediction"]] == [str(p) for p in got["prediction"]]
    assert want["classified"].tolist() == got["classified"].tolist()
    assert np.array_equal(want[classes].to_numpy(), got[classes].to_numpy())
    if tname != "scalar":
        a = ref_classification.class_df_probs_only([path], thr)
        b = classification.class_df_probs_only([path], thr)
        assert a.index.tolist() == b.index.tolist() and list(a.columns) == list(b.columns)
        assert np.array_equal(a.to_numpy(), b.to_numpy())
        assert int(a["Total"].iloc[0]) == len(g["roi_id"])
