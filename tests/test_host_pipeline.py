"""Host pipeline (sykepic_b200/pipeline.py) without a GPU: a stand-in engine checks the order of work, the per-bin error
policy of the reference (sykepic/compute/probability.py:106-114: log and carry on) and the skip / force rule (:136-141)."""
import numpy as np
import torch

from sykepic_b200 import engine, pipeline, synth


class _Handle:
    def __init__(self, value, error=None):
        self.value, self.error = value, error

    def result(self):
        if self.error:
            raise self.error
        return self.value


class FakeNet:
    """Same surface as Engine for BinPipeline: th / tw / torch / submit_rois -> handle."""

    th = tw = 64
    torch = torch

    def __init__(self, k, fail_on=()):
        self.k, self.fail_on, self.calls = k, set(fail_on), []

    def submit_rois(self, w, h, start, roi, roi_len=None, batch_size=None, want_labels=False):
        n = len(w)
        self.calls.append(n)
        if n in self.fail_on:
            return _Handle(None, RuntimeError("device fault"))
        buf = roi.numpy() if torch.is_tensor(roi) else roi
        # "probabilities": a deterministic function of the ROI bytes, so that a mixed-up buffer would show
        first = np.array([buf[int(s)] if int(ww) * int(hh) else 0 for s, ww, hh in zip(start, w, h)], np.float32)
        probs = np.tile((first[:, None] + 1.0) / 512.0, (1, self.k)).astype(np.float32)
        return _Handle(probs)


def test_pipeline_order_errors_and_skip(tmp_path):
    raw, out = tmp_path / "raw", tmp_path / "out"
    classes = ["a", "b", "c"]
    names, bins = [], []
    for i in range(7):
        b = synth.synth_bin(500 + i, 30 + i)
        names.append(synth.bin_name(i))
        bins.append(b)
        synth.write_bin(raw, names[-1], b)
    # bin 2: truncated .roi -> "Faulty raw data" (ValueError in the loader); bin 4: the device call fails -> "Unexpected error"
    (raw / f"{names[2]}.roi").write_bytes(bins[2]["roi_bytes"][:-9].tobytes())
    n4 = int((bins[4]["w"] > 0).sum())
    # bin 5 already has a CSV: skipped but reported as processed
    existing = engine_csv_path(raw / names[5], out)
    existing.parent.mkdir(parents=True, exist_ok=True)
    existing.write_text("sentinel")
    net = FakeNet(len(classes), fail_on={n4})
    pipe = pipeline.BinPipeline(net, classes, out, batch_size=16, loaders=2, writers=2, depth=2)
    done = pipe.run([raw / n for n in names])
    assert done == {names[i] for i in (0, 1, 3, 5, 6)}
    assert existing.read_text() == "sentinel"
    assert not engine_csv_path(raw / names[2], out).exists() and not engine_csv_path(raw / names[4], out).exists()
    # submissions reach the engine in the order given (the faulty and the skipped bin never do)
    assert net.calls == [int((bins[i]["w"] > 0).sum()) for i in (0, 1, 3, 4, 6)]
    for i in (0, 1, 3, 6):
        text = engine_csv_path(raw / names[i], out).read_text().splitlines()
        rid, w, h, start = engine.parse_adc(bins[i]["adc_text"])
        assert text[0] == "roi,a,b,c" and len(text) == len(rid) + 1
        first = (float(bins[i]["roi_bytes"][start[0]]) + 1.0) / 512.0
        assert text[1] == f"{rid[0]}," + ",".join([f"{np.float32(first):.5f}"] * 3)
    # force overwrites
    pipe = pipeline.BinPipeline(FakeNet(len(classes)), classes, out, force=True)
    assert pipe.run([raw / names[5]]) == {names[5]}
    assert existing.read_text().startswith("roi,a,b,c")
    assert pipe.stats["bins"] == 1 and pipeline.LAST_STATS[-1]["rois"] == pipe.stats["rois"]


def engine_csv_path(sample_path, out):
    from sykepic_b200.utils import files

    return files.sample_csv_path(sample_path, out, suffix=".prob")
