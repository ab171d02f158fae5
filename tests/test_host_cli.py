"""`python -m sykepic_b200 prob|class`: the reference's flags, defaults and exclusivity (sykepic/__main__.py:63-99, 134-190)."""
import re
from pathlib import Path

import pytest

from sykepic_b200.__main__ import build_parser

REF_MAIN = Path("/root/reference/sykepic/__main__.py")


def _sub(parser, name):
    for a in parser._actions:
        if getattr(a, "choices", None) and name in a.choices:
            return a.choices[name]
    raise AssertionError(name)


def test_prob_flags_and_defaults():
    p = build_parser()
    a = p.parse_args(["prob", "-r", "RAW", "-m", "MODEL", "-o", "OUT"])
    assert (a.raw, a.samples, a.image_dir, a.images, a.model, a.out) == ("RAW", None, None, None, "MODEL", "OUT")
    assert (a.batch_size, a.num_workers, a.force) == (64, 2, False)
    a = p.parse_args(["prob", "--samples", "a", "b", "-m", "M", "-o", "O", "-b", "256", "-w", "8", "-f"])
    assert a.samples == ["a", "b"] and (a.batch_size, a.num_workers, a.force) == (256, 8, True)
    assert p.parse_args(["prob", "--image-dir", "D", "-m", "M", "-o", "O"]).image_dir == "D"
    assert p.parse_args(["prob", "--images", "x.png", "y.png", "-m", "M", "-o", "O"]).images == ["x.png", "y.png"]
    for bad in (["prob", "-m", "M", "-o", "O"],  # one input source is required
                ["prob", "-r", "R", "--image-dir", "D", "-m", "M", "-o", "O"],  # ... and only one
                ["prob", "-r", "R", "-o", "O"], ["prob", "-r", "R", "-m", "M"]):
        with pytest.raises(SystemExit):
            p.parse_args(bad)


def test_class_flags_and_defaults():
    p = build_parser()
    a = p.parse_args(["class", "PROBS", "-t", "THR", "-o", "out.csv"])
    assert (a.probabilities, a.thresholds, a.out, a.feat, a.divisions) == ("PROBS", "THR", "out.csv", None, None)
    assert (a.value_column, a.append, a.force) == ("biomass_ugl", False, False)
    a = p.parse_args(["class", "P", "--feat", "F", "-t", "T", "-d", "DIV", "-o", "o.csv", "-v", "volume", "-a", "-f", "-exc", "EX"])
    assert (a.feat, a.divisions, a.value_column, a.append, a.force, a.exclusion_list) == ("F", "DIV", "volume", True, True, "EX")
    for bad in (["class", "P", "-o", "o.csv"], ["class", "P", "-t", "T"], ["class", "-t", "T", "-o", "o.csv"]):
        with pytest.raises(SystemExit):
            p.parse_args(bad)


@pytest.mark.skipif(not REF_MAIN.is_file(), reason="the reference checkout is only present in the build container")
def test_every_option_string_of_the_reference_is_accepted():
    text = REF_MAIN.read_text()
    parser = build_parser()
    for sub, var in (("prob", r"prob_(?:parser|raw)"), ("class", r"class_parser")):
        ours = set(_sub(parser, sub)._option_string_actions)
        theirs = set()
        for m in re.finditer(var + r"\.add_argument\((.*?)\)", text, re.S):
            theirs |= set(re.findall(r'"(--?[A-Za-z][\w-]*)"', m.group(1)))
        assert theirs and theirs <= ours, (sub, sorted(theirs - ours))


def test_prob_call_hands_main_the_work_list(tmp_path, monkeypatch):
    """`probability.call` (sykepic/compute/probability.py:22-64 of the reference): the four input modes -> what `main` gets."""
    from types import SimpleNamespace

    from sykepic_b200.compute import probability

    seen = {}
    monkeypatch.setattr(probability, "main", lambda *a, **k: seen.update(args=a, kw=k) or "done")
    raw = tmp_path / "raw" / "2021" / "05"
    raw.mkdir(parents=True)
    for name, size in (("D20210523T000000_IFCB114", 10), ("D20210523T002000_IFCB114", 20)):
        (raw / f"{name}.roi").write_bytes(bytes(size))
        (raw / f"{name}.adc").write_text("")
    base = dict(raw=None, samples=None, image_dir=None, images=None, model="M", out="O", batch_size=64, num_workers=2, force=False)
    assert probability.call(SimpleNamespace(**{**base, "raw": str(tmp_path / "raw")})) == "done"
    work = seen["args"][0]
    assert sorted(p.name for p in work) == ["D20210523T000000_IFCB114", "D20210523T002000_IFCB114"] and all(p.suffix == "" for p in work)
    assert seen["args"][1:] == ("M", "O", 64, 2, False) and seen["kw"]["samples_as_images"] is False and seen["kw"]["progress_bar"] is True
    # explicit sample paths; a bin over the size limit is left out (probability.py:45-51)
    monkeypatch.setattr(probability, "ROI_FILE_LIMIT", 15)
    probability.call(SimpleNamespace(**{**base, "samples": [str(raw / "D20210523T000000_IFCB114"), str(raw / "D20210523T002000_IFCB114")]}))
    assert [p.name for p in seen["args"][0]] == ["D20210523T000000_IFCB114"]
    # image modes: grouped by sample, files sorted
    imgs = tmp_path / "imgs" / "sub"
    imgs.mkdir(parents=True)
    for n in ("A_IFCB1_00002.png", "A_IFCB1_00001.png", "B_IFCB1_00007.png"):
        (imgs / n).write_bytes(b"")
    probability.call(SimpleNamespace(**{**base, "image_dir": str(tmp_path / "imgs")}))
    work = seen["args"][0]
    assert list(work) == ["A_IFCB1", "B_IFCB1"] and [p.name for p in work["A_IFCB1"]] == ["A_IFCB1_00001.png", "A_IFCB1_00002.png"]
    assert seen["kw"]["samples_as_images"] is True
    probability.call(SimpleNamespace(**{**base, "images": [str(imgs / "B_IFCB1_00007.png"), str(imgs / "A_IFCB1_00002.png")]}))
    assert {k: [p.name for p in v] for k, v in seen["args"][0].items()} == {"A_IFCB1": ["A_IFCB1_00002.png"], "B_IFCB1": ["B_IFCB1_00007.png"]}


def test_default_precision_is_fp32_accuracy_on_the_tensor_cores(monkeypatch):
    """The reference computes in fp32 (compute/probability.py:189-194).  The CLI's default precision is the one that keeps the
    1e-4 gate AND runs on tcgen05 ("fp32_tc"); `--precision` and SYKEPIC_PRECISION select the exact CUDA-core path or bf16."""
    import importlib

    from sykepic_b200.compute import probability

    monkeypatch.delenv("SYKEPIC_PRECISION", raising=False)
    importlib.reload(probability)
    assert probability.DEFAULT_PRECISION == "fp32_tc"
    p = build_parser()
    assert p.parse_args(["prob", "-r", "R", "-m", "M", "-o", "O"]).precision is None  # -> DEFAULT_PRECISION
    assert p.parse_args(["prob", "-r", "R", "-m", "M", "-o", "O", "--precision", "fp32"]).precision == "fp32"
    with pytest.raises(SystemExit):
        p.parse_args(["prob", "-r", "R", "-m", "M", "-o", "O", "--precision", "fp16"])
    monkeypatch.setenv("SYKEPIC_PRECISION", "bf16")
    importlib.reload(probability)
    assert probability.DEFAULT_PRECISION == "bf16"
    monkeypatch.delenv("SYKEPIC_PRECISION")
    importlib.reload(probability)
