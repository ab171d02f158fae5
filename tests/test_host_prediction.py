"""Host mirror of sykepic.compute.prediction / classification against the labels and counts the
REFERENCE produced (tests/golden/*.labels.json) and its own fixture CSV."""

import json
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest

from sykepic_b200.compute import classification, prediction
from sykepic_b200.utils import files, ifcb
from tests.cases import CASES, FIXTURE, GOLDEN, VALID_BIN


def _thr(tname):
    return 0.5 if tname.startswith("scalar") else FIXTURE / f"{tname}.txt"


@pytest.mark.parametrize("case", list(CASES))
def test_prediction_dataframe_matches_reference_labels(case):
    labels = json.loads((GOLDEN / f"case_{case}.labels.json").read_text())
    for bname, lab in labels.items():
        csv = GOLDEN / f"case_{case}__{bname}.prob.csv"
        for tname in ("thresholds-2021", "thresholds-zero", "scalar-0.5"):
            df = prediction.prediction_dataframe(csv, _thr(tname))
            assert list(df.columns[:2]) == ["prediction", "classified"]
            assert [str(p) for p in df["prediction"]] == lab[tname]["prediction"]
            assert [bool(c) for c in df["classified"]] == lab[tname]["classified"]
            assert str(df["prediction"].dtype) == "category"
            if "counts" in lab[tname]:
                counts = classification.class_df_probs_only([csv], _thr(tname))
                assert {k: int(v) for k, v in counts.iloc[0].items()} == lab[tname]["counts"]
                assert counts.index.tolist() == [csv.with_suffix("").stem]


def test_reference_fixture_labels_and_counts():
    """zero thresholds => Uroglenopsis_sp / Licmophora_sp classified, Total 2
    (the reference's tests/test_classification.py:61-65)."""
    csv = FIXTURE / f"{VALID_BIN}.prob.csv"
    real = json.loads((GOLDEN / "ref_fixture_labels.json").read_text())
    for tname, want in real.items():
        df = prediction.prediction_dataframe(csv, FIXTURE / f"{tname}.txt")
        assert [str(p) for p in df["prediction"]] == want["prediction"]
        assert [bool(c) for c in df["classified"]] == want["classified"]
        assert df.index.tolist() == want["roi"]
        counts = classification.class_df_probs_only([csv], FIXTURE / f"{tname}.txt")
        assert {k: int(v) for k, v in counts.iloc[0].items()} == want["counts"]
    counts = classification.class_df_probs_only([csv], FIXTURE / "thresholds-zero.txt")
    assert counts.loc[VALID_BIN, "Uroglenopsis_sp"] == 1 and counts.loc[VALID_BIN, "Total"] == 2


def test_row_prediction_and_list_input():
    csv = FIXTURE / f"{VALID_BIN}.prob.csv"
    df = pd.read_csv(csv, index_col=0)
    thr = prediction.threshold_dictionary(FIXTURE / "thresholds-zero.txt")
    assert prediction.row_prediction(df.iloc[0], thr) == ("Uroglenopsis_sp", True)
    assert prediction.row_prediction(df.iloc[0], 0.9) == ("Uroglenopsis_sp", False)
    multi = prediction.prediction_dataframe([csv, csv], thr)
    assert multi.index.names == ["sample", "roi"] and len(multi) == 4
    with pytest.raises(ValueError):
        prediction.prediction_dataframe(3.0)


def test_threshold_dictionary_missing_value(tmp_path):
    p = tmp_path / "t.txt"
    p.write_text("A 0.5\nB\n")
    with pytest.raises(ValueError):
        prediction.threshold_dictionary(p)
    assert prediction.threshold_dictionary(p, default=0.7) == {"A": 0.5, "B": 0.7}


def test_predict_array_ties_and_missing_classes():
    classes = ["a", "b", "c"]
    vals = np.array([[0.4, 0.4, 0.2], [0.1, 0.6, 0.3], [0.2, 0.3, 0.5]])
    idx, flag = prediction.predict_array(vals, classes, {"a": 0.3, "b": 0.3})
    assert idx.tolist() == [0, 1, 1] and flag.tolist() == [True, True, True]
    idx, flag = prediction.predict_array(vals, classes, {"a": 0.9})
    assert idx.tolist() == [0, 1, 2] and not flag.any()
    idx, flag = prediction.predict_array(vals, classes, 0.4)
    assert idx.tolist() == [0, 1, 2] and flag.tolist() == [False, True, True]


def test_class_main_writes_summary(tmp_path):
    """`sykepic class` end to end on the reference's fixture CSV.  The reference's own thresholds lack the
    coiled classes swell_df hard-codes, so the file is extended with them here."""
    thr = tmp_path / "thr.txt"
    thr.write_text((FIXTURE / "thresholds-zero.txt").read_text() +
                   "Dolichospermum-Anabaenopsis_coiled 0\nNodularia_spumigena-coiled 0\n")
    probs = tmp_path / "prob" / "2018" / "07" / "12"
    probs.mkdir(parents=True)
    (probs / f"{VALID_BIN}.prob.csv").write_text((FIXTURE / f"{VALID_BIN}.prob.csv").read_text())
    out = tmp_path / "out.csv"
    args = SimpleNamespace(probabilities=str(tmp_path / "prob"), feat=None, thresholds=str(thr), divisions=None,
                           out=str(out), value_column="biomass_ugl", append=False, force=False, exclusion_list=None)
    classification.main(args)
    df = pd.read_csv(out, index_col=0)
    assert df.index.name == "Time" and df.index.tolist() == ["2018-07-12T06:56:00+00:00"]
    assert df.columns[-1] == "Total" and df.columns[-2] == "Filamentous cyanobacteria"
    assert df.iloc[0]["Uroglenopsis sp"] == 1 and df.iloc[0]["Licmophora sp"] == 1 and df.iloc[0]["Total"] == 2
    with pytest.raises(FileExistsError):
        classification.main(args)
    args.append = True
    classification.main(args)
    assert len(pd.read_csv(out, index_col=0)) == 2
    # the reference's own thresholds file: KeyError in swell_df, as in the reference
    args.thresholds, args.force = str(FIXTURE / "thresholds-zero.txt"), True
    with pytest.raises(KeyError):
        classification.main(args)
    args.out = str(tmp_path / "out.txt")
    with pytest.raises(ValueError):
        classification.main(args)


def test_paths_and_listing(tmp_path):
    assert str(files.sample_csv_path("/x/D20180712T065600_IFCB114", "/out", ".prob")) == \
        "/out/2018/07/12/D20180712T065600_IFCB114.prob.csv"
    assert ifcb.sample_to_datetime("D20180712T065600_IFCB114", True) == "2018-07-12T06:56:00+00:00"
    (tmp_path / "a").mkdir()
    for n in ("a/D20180712T065600_IFCB114", "D20180713T065600_IFCB114"):
        (tmp_path / f"{n}.roi").write_bytes(b"")
    assert sorted(p.name for p in files.list_sample_paths(tmp_path)) == ["D20180712T065600_IFCB114", "D20180713T065600_IFCB114"]
    excl = tmp_path / "excl.txt"
    excl.write_text("D20180713T065600\n")
    kept = ifcb.filter_out_quality_flagged_samples(files.list_sample_paths(tmp_path), excl)
    assert [p.name for p in kept] == ["D20180712T065600_IFCB114"]


def test_raw_to_numpy_matches_fixture():
    rois = list(ifcb.raw_to_numpy(FIXTURE / f"{VALID_BIN}.adc", FIXTURE / f"{VALID_BIN}.roi"))
    assert [(i, r.shape) for i, r in rois] == [(2, (42, 56)), (3, (53, 128))]
