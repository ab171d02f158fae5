"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and its
host entry points (.adc parser, geometry validation, CSV formatter, threshold quantiser) agree with the
oracle / the reference goldens.  No compute call needs a GPU here."""

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import ifcb as o_ifcb
from oracle import prediction as o_pred
from oracle import preprocess as o_pre
from sykepic_b200 import _lib, engine
from tests.cases import CASES, FIXTURE, GOLDEN, VALID_BIN, case_bins

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "sykepic_b200.h").read_text()
    declared = set(re.findall(r"\b(spk_[a-z0-9_]+)\s*\(", header))
    declared -= {"spk_ctx"}
    assert len(declared) >= 25
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.spk_abi_version() == 1


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = _lib.load().spk_create(0, None, C.byref(ctx))
    assert rc == _lib.SPK_ERR_CUDA and not ctx.value
    assert "no CPU fallback" in _lib.last_error()
    with pytest.raises(_lib.SpkError):
        engine.Engine(object.__new__(engine.ModelSpec))


@pytest.mark.parametrize("case", list(CASES))
def test_adc_parse_matches_oracle(case):
    for bname, b in case_bins(case):
        want = o_ifcb.parse_adc_text(b["adc_text"])
        rid, w, h, start = engine.parse_adc(b["adc_text"])
        assert list(zip(rid.tolist(), w.tolist(), h.tolist(), start.tolist())) == want


def test_adc_parse_reference_fixture_and_newlines():
    text = (FIXTURE / f"{VALID_BIN}.adc").read_bytes()
    rid, w, h, start = engine.parse_adc(text)
    assert list(zip(rid.tolist(), w.tolist(), h.tolist(), start.tolist())) == [(2, 56, 42, 0), (3, 128, 53, 2352)]
    row = ",".join(["0"] * 15 + [" 8 ", "+4", "16"] + ["x"] * 6)
    for nl in ("\n", "\r\n", "\r"):
        txt = nl.join([row, row.replace(" 8 ", "0"), row]) + nl
        got = engine.parse_adc(txt)
        assert got[0].tolist() == [1, 3] and got[1].tolist() == [8, 8] and got[3].tolist() == [16, 16]
        assert [r[0] for r in o_ifcb.parse_adc_text(txt)] == [1, 3]
    # no trailing newline, empty text
    assert engine.parse_adc(row)[0].tolist() == [1]
    assert len(engine.parse_adc("")[0]) == 0


@pytest.mark.parametrize("bad", ["1,2,3\n", ",".join(["0"] * 15 + ["a", "1", "2"]) + "\n",
                                 ",".join(["0"] * 15 + ["1.5", "1", "2"]) + "\n"])
def test_adc_parse_errors(bad):
    with pytest.raises((ValueError, IndexError)):
        o_ifcb.parse_adc_text(bad)
    with pytest.raises(ValueError):
        engine.parse_adc(bad)


def test_validate_faulty_and_empty_resize():
    w = np.array([56, 128], np.int32)
    h = np.array([42, 53], np.int32)
    s = np.array([0, 2352], np.int64)
    engine.validate_rois(w, h, s, 9136, 180, 180)
    with pytest.raises(_lib.FaultyBin):  # the reference: reshape ValueError -> "Faulty raw data"
        engine.validate_rois(w, h, s, 9135, 180, 180)
    assert issubclass(_lib.FaultyBin, ValueError)
    with pytest.raises(_lib.EmptyResize):  # aspect ratio > T:1 -> 0-pixel side -> cv2.error in the reference
        engine.validate_rois(np.array([1000], np.int32), np.array([2], np.int32), np.array([0], np.int64), 2000, 180, 180)
    with pytest.raises(ValueError):
        o_pre.resize_with_border_u8(np.zeros((2, 1000), np.uint8), 180, 180)


def test_new_dims_matches_oracle():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(2000):
        h, w = int(rng.integers(1, 1400)), int(rng.integers(1, 1400))
        t = int(rng.choice([180, 224, 299]))
        nh, nw = C.c_int(), C.c_int()
        lib.spk_new_dims(h, w, t, t, C.byref(nh), C.byref(nw))
        assert (nh.value, nw.value) == o_pre.get_new_dims(h, w, t, t)


def test_csv_format_matches_python_formatting():
    rng = np.random.default_rng(1)
    probs = rng.random((257, 50), dtype=np.float32)
    # values that sit on or next to a rounding boundary of the 5th decimal
    edge = np.array([0.0, 1.0, 0.5, 0.000005, 0.000015, 0.000025, 0.123455, 0.123465, 0.999995, 0.9999949,
                     1e-9, 0.00000499999, 0.30000001192092896], np.float32)
    probs[0, :len(edge)] = edge
    probs[1] = np.float32(1) / np.arange(1, 51, dtype=np.float32)
    ids = np.arange(2, 259, dtype=np.int32)
    classes = [f"c{i}" for i in range(50)]
    got = engine.format_prob_csv(classes, ids, probs).decode()
    want = o_pred.probabilities_to_csv_text(zip(ids.tolist(), probs.tolist()), classes)
    assert got == want
    assert engine.format_prob_csv(classes, ids[:0], probs[:0]).decode() == "roi," + ",".join(classes) + "\n"


@pytest.mark.parametrize("case", list(CASES))
def test_csv_format_reproduces_reference_csv(case):
    """The reference's CSV text from the reference's own fp32 probabilities (goldens)."""
    for bname, _ in case_bins(case):
        g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
        classes = (FIXTURE / "class_names.txt").read_text().splitlines()
        got = engine.format_prob_csv(classes, g["roi_id"], g["probs"]).decode()
        want = (GOLDEN / f"case_{case}__{bname}.prob.csv").read_text()
        a, b = o_pred.parse_prob_csv_text(got), o_pred.parse_prob_csv_text(want)
        assert a[0] == b[0] and a[1].tolist() == b[1].tolist()
        # the goldens' probabilities were recomputed in batches of 16 (make_golden.py), the CSV in the
        # reference's own batches: identical up to one unit of the last printed decimal
        assert np.abs(a[2] - b[2]).max() <= 1.0001e-5


def test_threshold_quantize_is_the_decimal_comparison():
    lib = _lib.load()
    rng = np.random.default_rng(2)
    thr = np.concatenate([rng.random(300), [0.0, 1.0, 0.5, 0.9, 0.95, 0.99, 0.1, 0.3, 0.70000001, 1e-5, 0.99999, 2.0, -1.0]])
    for t in thr:
        for strict in (0, 1):
            q = lib.spk_threshold_quantize(float(t), strict)
            for cand in (q - 1, q, q + 1):
                if cand < 0 or cand > 100000:
                    continue
                value = float(f"{cand / 1e5:.5f}")  # what pandas reads back from the CSV
                above = value > t if strict else value >= t
                assert above == (cand >= q), (t, strict, q, cand)


def test_adc_parse_agrees_with_python_int_on_random_fields():
    """Fields 15-17 go through Python's int() in the reference (utils/ifcb.py:104-106): optional whitespace, sign, digits with
    single underscores.  Random strings over that alphabet (plus letters and dots) either parse to the same numbers or fail in both."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    field = st.text(alphabet=" \t+-_0123456789.a", min_size=0, max_size=8)

    @settings(max_examples=400, deadline=None)
    @given(st.lists(st.tuples(field, field, field), min_size=1, max_size=4), st.sampled_from(["\n", "\r\n"]))
    def check(rows, nl):
        text = nl.join(",".join(["0"] * 15 + list(r) + ["x"] * 6) for r in rows) + nl
        try:
            want = o_ifcb.parse_adc_text(text)
        except (ValueError, IndexError):
            want = None
        try:
            rid, w, h, start = engine.parse_adc(text)
            got = list(zip(rid.tolist(), w.tolist(), h.tolist(), start.tolist()))
        except ValueError:
            got = None
        if want is not None and any(abs(v) >= 2 ** 31 for r in want for v in r[1:3]):
            return  # widths beyond int32: Python's ints are unbounded, the C ABI reports a parse error
        assert got == want, (text, got, want)

    check()


def test_bin_load_reads_parses_and_validates(tmp_path):
    """`spk_bin_load`: one call = both file reads + .adc parse + geometry checks; same descriptors as `spk_adc_parse`,
    the .roi bytes land in the caller's buffer, and the reference's per-bin error cases map to the same exceptions."""
    from sykepic_b200 import synth

    b = synth.synth_bin(77, 60)
    path = synth.write_bin(tmp_path, "D20210601T000000_IFCB114", b)
    buf = np.full(len(b["roi_bytes"]) + 64, 0xAB, np.uint8)
    rid, w, h, start, roi_len = engine.load_bin(path, buf, 224, 224)
    want = engine.parse_adc(b["adc_text"])
    for got, exp in zip((rid, w, h, start), want):
        assert np.array_equal(got, exp)
    assert roi_len == len(b["roi_bytes"]) and np.array_equal(buf[:roi_len], b["roi_bytes"]) and (buf[roi_len:] == 0xAB).all()
    # buffer too small: the needed size is reported, nothing is written past the buffer
    small = np.zeros(100, np.uint8)
    with pytest.raises(engine.CapacityError) as ei:
        engine.load_bin(path, small, 224, 224)
    assert ei.value.needed == roi_len
    # truncated .roi -> FaultyBin (a ValueError: "Faulty raw data", probability.py:111-112)
    (tmp_path / "D20210601T000000_IFCB114.roi").write_bytes(b["roi_bytes"][:-3].tobytes())
    with pytest.raises(_lib.FaultyBin):
        engine.load_bin(path, buf, 224, 224)
    # missing file -> OSError ("Unexpected error", :113-114); malformed .adc -> AdcParseError (a ValueError)
    with pytest.raises(OSError):
        engine.load_bin(tmp_path / "nope", buf, 224, 224)
    bad = synth.write_bin(tmp_path, "D20210601T002000_IFCB114", {"adc_text": "1,2,3\n", "roi_bytes": np.zeros(4, np.uint8)})
    with pytest.raises(_lib.AdcParseError):
        engine.load_bin(bad, buf, 224, 224)
    # empty bin: no ROIs, no bytes
    empty = synth.write_bin(tmp_path, "D20210601T004000_IFCB114", {"adc_text": "", "roi_bytes": np.zeros(0, np.uint8)})
    rid, w, h, start, roi_len = engine.load_bin(empty, buf, 224, 224)
    assert len(rid) == 0 and roi_len == 0


def test_prob_csv_write_equals_format(tmp_path):
    rng = np.random.default_rng(3)
    classes = [f"c{i}" for i in range(50)]
    probs = rng.random((321, 50), dtype=np.float32)
    rid = np.arange(1, 322, dtype=np.int32)
    n = engine.write_prob_csv(tmp_path / "x.prob.csv", classes, rid, probs)
    data = (tmp_path / "x.prob.csv").read_bytes()
    assert n == len(data) and data == engine.format_prob_csv(classes, rid, probs)
    with pytest.raises(OSError):
        engine.write_prob_csv(tmp_path / "no_such_dir" / "x.csv", classes, rid, probs)
