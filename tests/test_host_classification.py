"""`sykepic class --feat` (labels joined with feature CSVs) against what the REFERENCE returned
for the same files (tests/golden/make_class_feat_golden.py), and against the reference's own fixture value."""
from collections import namedtuple
from pathlib import Path

import pytest

from sykepic_b200.compute import classification

G = Path(__file__).resolve().parent / "golden" / "class_feat"
FIX = Path(__file__).resolve().parent / "golden" / "ref_fixture"
Args = namedtuple("Args", "probabilities feat thresholds divisions out value_column append force exclusion_list")


def _lists():
    return sorted((G / "prob").glob("*.csv")), sorted((G / "feat").glob("*.csv"))


@pytest.mark.parametrize("tag", ["nodiv", "div"])
@pytest.mark.parametrize("column", ["biomass_ugl", "biovolume_um3", "frequency"])
def test_class_df_matches_reference(tag, column):
    probs, feats = _lists()
    div = G / "divisions.txt" if tag == "div" else None
    got = classification.class_df(probs, feats, G / "thresholds.txt", div, column).to_csv()
    assert got == (G / f"ref_class_df_{tag}_{column}.csv").read_text()


def test_main_with_features_and_probs_only(tmp_path):
    out = tmp_path / "summary.csv"
    classification.main(Args(str(G / "prob"), str(G / "feat"), str(G / "thresholds.txt"), None, out, "biomass_ugl", False, False, None))
    assert out.read_text() == (G / "ref_main_biomass.csv").read_text()
    with pytest.raises(FileExistsError):  # classification.py:32-34
        classification.main(Args(str(G / "prob"), str(G / "feat"), str(G / "thresholds.txt"), None, out, "biomass_ugl", False, False, None))
    out2 = tmp_path / "counts.csv"
    classification.main(Args(str(G / "prob"), None, str(G / "thresholds.txt"), None, out2, None, False, False, None))
    assert out2.read_text() == (G / "ref_main_probs_only.csv").read_text()


def test_sample_volume_is_the_last_comment_line():
    assert classification.sample_volume(G / "feat" / "D20210523T000000_IFCB114.feat.csv") == "4.25"


def test_divide_row_names():
    d = {"A": [10, 20]}
    mk = lambda v: classification.divide_row({"prediction": "A", "x": v}, d, "x")["prediction"]  # noqa: E731
    assert mk(5) == "A_10_20"  # the last limit decides, as in the reference's loop
    assert mk(15) == "A_10_20"
    assert mk(25) == "A_over_20"
    assert classification.divide_row({"prediction": "B", "x": 1}, d, "x")["prediction"] == "B"
    assert sorted(classification.names_of_divisions(d)) == ["A_10_20", "A_over_20", "A_under_10"]


def test_reference_fixture_total_biomass():
    """The reference's own test pins Total biomass 1.782 for its fixture bin (tests/test_classification.py:12-37)."""
    df = classification.class_df([FIX / "D20180712T065600_IFCB114.prob.csv"], [FIX / "D20180712T065600_IFCB114.feat.csv"],
                                 FIX / "thresholds-2021.txt")
    assert df.shape == (1, 51)
    assert float(df["Total"].iloc[0]) == pytest.approx(1.7823066272440204, rel=1e-12)
