"""`read_prob_csv` (C-ABI parser behind prediction_dataframe) against pandas.read_csv, which is what the reference uses
(sykepic/compute/prediction.py:8-28): identical frames on every golden CSV, on all 100001 five-decimal values, and the
fall-back to pandas for anything that is not the plain layout."""
import io

import numpy as np
import pandas as pd
import pytest

from sykepic_b200 import engine
from sykepic_b200.compute import prediction
from tests.cases import FIXTURE, GOLDEN, VALID_BIN


def _same(a, b):
    pd.testing.assert_frame_equal(a, b, check_exact=True)


@pytest.mark.parametrize("path", sorted(GOLDEN.glob("*.prob.csv")) + [FIXTURE / f"{VALID_BIN}.prob.csv"], ids=lambda p: p.name[:40])
def test_golden_files_read_like_pandas(path):
    _same(prediction.read_prob_csv(path, index_col=0), pd.read_csv(path, index_col=0))
    _same(prediction.read_prob_csv(path), pd.read_csv(path))


def test_every_five_decimal_value(tmp_path):
    classes = [f"c{i}" for i in range(50)]
    flat = [f"{i / 1e5:.5f}" for i in range(100001)]
    flat += ["1.00000"] * (-len(flat) % 50)  # pad the last row
    text = "roi," + ",".join(classes) + "\n"
    lines = [f"{r + 1}," + ",".join(flat[50 * r: 50 * r + 50]) for r in range(len(flat) // 50)]
    k = len(flat)
    p = tmp_path / "all.prob.csv"
    p.write_text(text + "\n".join(lines) + "\n")
    a, b = prediction.read_prob_csv(p, index_col=0), pd.read_csv(p, index_col=0)
    _same(a, b)
    want = np.array([float(x) for x in flat[:k]]).reshape(-1, 50)
    assert (a.to_numpy() == want).all()  # == Python float(): correctly rounded


def test_round_trip_of_the_writer(tmp_path):
    rng = np.random.default_rng(3)
    probs = rng.dirichlet(np.ones(7) * 0.2, 300).astype(np.float32)
    ids = np.sort(rng.choice(10000, 300, replace=False)).astype(np.int32) + 1
    p = tmp_path / "w.prob.csv"
    p.write_bytes(engine.format_prob_csv(["a b", "Dolichospermum-Anabaenopsis_coiled", "c", "d", "e", "f", "g"], ids, probs))
    a = prediction.read_prob_csv(p, index_col=0)
    _same(a, pd.read_csv(p, index_col=0))
    assert a.index.tolist() == ids.tolist() and a.index.name == "roi" and a.columns[1] == "Dolichospermum-Anabaenopsis_coiled"
    # CRLF line ends and blank lines, as pandas takes them
    crlf = tmp_path / "crlf.prob.csv"
    crlf.write_bytes(p.read_bytes().replace(b"\n", b"\r\n") + b"\r\n\r\n")
    _same(prediction.read_prob_csv(crlf, index_col=0), pd.read_csv(crlf, index_col=0))


@pytest.mark.parametrize("text", [
    "roi,a,b\n",  # header only (empty bin)
    'roi,"a,x",b\n1,0.5,0.5\n',  # quoted header
    "roi,a,b\n1,0.5\n2,0.1,0.9\n",  # ragged
    "roi,a,b\n1,0.5,abc\n",  # not a number
    "roi,a,a\n1,0.5,0.5\n",  # duplicate names (pandas renames)
    "roi,a,b\n1,5e-05,1E-3\n2,nan,inf\n",  # exponents and specials: strtod path, same values as pandas
    "roi,a,b\n1,0.1234567890123456789,12345678901234567890.5\n",  # long mantissas
    "roi,a,b\nx1,0.5,0.5\n",  # non-numeric id
], ids=lambda t: repr(t)[:28])
def test_odd_files_behave_like_pandas(tmp_path, text):
    p = tmp_path / "odd.prob.csv"
    p.write_text(text)
    for kw in ({"index_col": 0}, {}):
        try:
            want = pd.read_csv(p, **kw)
        except Exception as e:  # whatever pandas raises, the wrapper raises too
            with pytest.raises(type(e)):
                prediction.read_prob_csv(p, **kw)
            continue
        _same(prediction.read_prob_csv(p, **kw), want)
