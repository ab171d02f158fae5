"""Generate tests/golden/* by running the REFERENCE itself (build container only).

    python tests/golden/make_golden.py [case ...]      (default: every case of tests/cases.py)

Imports `/root/reference` read-only (sykepic.compute.probability /
prediction / classification, sykepic.utils.ifcb, sykepic.train.*) with the shims
of SURVEY.md section 8c: a dummy `pytz` module injected after pandas, synthetic
model dirs with `weights =` empty and a seeded `best_state.pth`, raw bins copied
to a writable scratch dir.  Nothing from the reference's sources is copied; only
its data fixtures (tests/data/raw/valid, tests/data/prob, tests/model/*.txt) and
the OUTPUTS of its code are committed:

  ref_fixture/           the reference's own fixture files for this path
  invalid_adc_geometry.npz   columns 15/16/17 of tests/data/raw/invalid/*.adc
  case_<name>.npz        per case: adc rows the reference decoded, decoded ROI
                         bytes digest, padded/resized uint8 taps, fp32 tensor
                         digests, logits/probabilities of `TorchVisionNet`
  case_<name>.prob.csv   the CSV `probability.main` wrote
  case_<name>.labels.json  `prediction_dataframe` labels + `class_df_probs_only`
                         counts for thresholds-2021 / thresholds-zero / scalar 0.5

The GPU box has no /root/reference: tests only read these files.
"""

import hashlib
import json
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

import pandas  # noqa: E402,F401  (must be imported before the pytz stub)

sys.modules.setdefault("pytz", types.ModuleType("pytz")).timezone = lambda name: None
sys.path.insert(0, str(REF))

import cv2  # noqa: E402
import torch  # noqa: E402
from sykepic.compute import classification, prediction, probability  # noqa: E402
from sykepic.utils import ifcb as ref_ifcb  # noqa: E402

from sykepic_b200 import synth  # noqa: E402
from tests.cases import ALL_CASES, LOGIT_GAIN, case_bins  # noqa: E402

torch.set_num_threads(8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _reference_net(arch, n_classes):
    """The reference's module for this architecture.  DenseNet: `TorchVisionNet` drops torchvision's functional
    ReLU + pooling and raises at 224x224, so the defined behaviour -- torchvision's own forward with the
    classifier replaced by the syke-pic head -- is built here from torchvision directly."""
    from sykepic.train.network import TorchVisionNet

    net = TorchVisionNet(arch, n_classes, None, [256, 128], [])
    if arch.startswith("densenet"):
        class TvForward(torch.nn.Module):
            def __init__(self, inner):
                super().__init__()
                self.base, self.head = inner.base, inner.head

            def forward(self, x):
                f = torch.nn.functional.relu(self.base(x))
                return self.head(torch.nn.functional.adaptive_avg_pool2d(f, 1).flatten(1))

        return TvForward(net)
    return net


def calibrate_bn(name, arch, t, border, norm, classes, seed, gain=LOGIT_GAIN, res_gamma=1.0):
    """BatchNorm running statistics measured on synthetic ROIs (one train-mode pass, momentum 1).

    A random-init network with random running stats maps every ROI to nearly
    the same feature vector; with statistics that match its own activations it
    is as input-sensitive as a trained checkpoint, so the probability / label
    gates are exercised.  The result is committed (calib_<case>.npz) so the
    checkpoint is reproducible from the seed + this file anywhere.
    """
    sd = synth.synth_state_dict(arch, len(classes), (256, 128), seed, True, gain, res_gamma=res_gamma)
    net = _reference_net(arch, len(classes))
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    b = synth.synth_bin(500 + seed, 96)
    tmp = Path(tempfile.mkdtemp(prefix="calib_"))
    try:
        sp = synth.write_bin(tmp, "D20200101T000000_IFCB114", b)
        md = synth.write_model_dir(tmp / "m", arch=arch, t=t, seed=seed, border=border, imagenet_normalization=norm,
                                   classes=classes)
        transform = probability.prepare_model(md)[3]
        xs = [transform(cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)) for _, img in
              ref_ifcb.raw_to_numpy(sp.with_suffix(".adc"), sp.with_suffix(".roi"))]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    net.train()
    with torch.no_grad():
        net(torch.stack(xs))
    stats = {k: v.numpy().astype(np.float32) for k, v in net.state_dict().items()
             if k.endswith("running_mean") or k.endswith("running_var")}
    np.savez_compressed(HERE / f"calib_{name}.npz", **stats)
    return stats


def run_case(name, arch, t, border, norm, bins, classes, thresholds_files, tap_limit, seed, gain=LOGIT_GAIN, source="reference",
             res_gamma=1.0):
    """source "reference": everything below is produced by the reference's own code.  source "oracle" (DenseNet, which
    the reference cannot run): decode / transform still come from the reference; network, CSV and labels from the oracle
    restatement of torchvision's forward, cross-checked here against the torchvision module itself."""
    print(f"== case {name} ({source})")
    stats = calibrate_bn(name, arch, t, border, norm, classes, seed, gain, res_gamma)
    tmp = Path(tempfile.mkdtemp(prefix="golden_"))
    try:
        model_dir = synth.write_model_dir(tmp / "model", arch=arch, t=t, head=(256, 128), seed=seed, border=border,
                                          imagenet_normalization=norm, classes=classes, logit_gain=gain,
                                          bn_stats=stats, res_gamma=res_gamma)
        raw = tmp / "raw"
        sample_paths = []
        for bname, b in bins:
            sample_paths.append(synth.write_bin(raw, bname, b))
        out = tmp / "out"
        if source == "reference":
            processed = probability.main(sample_paths, model_dir, out, batch_size=16, num_workers=0, force=False,
                                         progress_bar=False)
            assert processed == {p.name for p in sample_paths}, processed
            net, cls, img_shape, transform, device = probability.prepare_model(model_dir)
        else:
            from oracle import pipeline as o_pipeline

            from sykepic.train import config as ref_config
            from configparser import ConfigParser

            cfg = ConfigParser()
            cfg.read(model_dir / "config.ini")
            img_shape = ref_config.get_img_shape(cfg)
            transform = ref_config.get_transforms(cfg, img_shape)[1]
            net = _reference_net(arch, len(classes))
            net.load_state_dict(torch.load(model_dir / "best_state.pth"), strict=True)
            o_model = o_pipeline.prepare_model(model_dir)
            for sp in sample_paths:
                csv_path = ref_ifcb_csv_path(sp, out)
                csv_path.parent.mkdir(parents=True, exist_ok=True)
                csv_path.write_text(o_pipeline.process_bin(o_model, sp.with_suffix(".adc"), sp.with_suffix(".roi"), 16))
        net.eval()
        result = {}
        labels = {}
        for (bname, b), sp in zip(bins, sample_paths):
            csv_path = next(out.glob(f"**/{bname}.prob.csv"))
            rel = csv_path.relative_to(out)
            shutil.copy(csv_path, HERE / f"case_{name}__{bname}.prob.csv")
            ids, ws, hs, u8_taps, f32_digest, roi_digest, xs = [], [], [], [], [], [], []
            for rid, img in ref_ifcb.raw_to_numpy(sp.with_suffix(".adc"), sp.with_suffix(".roi")):
                ids.append(rid)
                hs.append(img.shape[0])
                ws.append(img.shape[1])
                roi_digest.append(sha(img))
                # what cv2.imread + cvtColor give for the PNG the reference writes (lossless, 3 equal planes)
                img3 = cv2.cvtColor(cv2.cvtColor(img, cv2.COLOR_GRAY2BGR), cv2.COLOR_BGR2RGB)
                x = transform(img3)
                xs.append(x)
                f32_digest.append(sha(x.numpy()))
                if len(u8_taps) < tap_limit:
                    # uint8 stage of the same transform (Resize only, no ToTensor)
                    from sykepic.train.image import Compose, Resize

                    u8 = Compose([Resize()], img_shape[1:], border)(img3)
                    assert (u8[..., 0] == u8[..., 1]).all() and (u8[..., 0] == u8[..., 2]).all()
                    u8_taps.append(u8[..., 0].copy())
            with torch.no_grad():
                logits = torch.cat([net(torch.stack(xs[i : i + 16])) for i in range(0, len(xs), 16)])
                probs = torch.softmax(logits * np.log(probability.SOFTMAX_EXP), dim=1)
            np.savez_compressed(
                HERE / f"case_{name}__{bname}.npz",
                roi_id=np.array(ids, np.int32), w=np.array(ws, np.int32), h=np.array(hs, np.int32),
                roi_sha=np.array(roi_digest), f32_sha=np.array(f32_digest),
                u8_taps=np.stack(u8_taps) if u8_taps else np.zeros((0, t, t), np.uint8),
                logits=logits.numpy(), probs=probs.numpy(),
                roi_bytes_sha=np.array(sha(b["roi_bytes"])), adc_sha=np.array(hashlib.sha256(b["adc_text"].encode()).hexdigest()),
            )
            lab = {"csv_relpath": str(rel)}
            for tname, thr in thresholds_files.items():
                df = prediction.prediction_dataframe(csv_path, thr)
                entry = {"prediction": [str(p) for p in df["prediction"]] if len(df) else [],
                         "classified": [bool(c) for c in df["classified"]] if len(df) else []}
                if not isinstance(thr, float) and len(df):
                    counts = classification.class_df_probs_only([csv_path], thr)
                    entry["counts"] = {k: int(v) for k, v in counts.iloc[0].items()}
                lab[tname] = entry
            labels[bname] = lab
        (HERE / f"case_{name}.labels.json").write_text(json.dumps(labels, indent=1))
        return result
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_ifcb_csv_path(sample_path, out_dir):
    from sykepic.utils import files as ref_files

    return Path(ref_files.sample_csv_path(sample_path, out_dir, suffix=".prob"))


def main():
    wanted = sys.argv[1:] or list(ALL_CASES)
    fx = HERE / "ref_fixture"
    fx.mkdir(exist_ok=True)
    for rel in ("tests/data/raw/valid/D20180712T065600_IFCB114.adc", "tests/data/raw/valid/D20180712T065600_IFCB114.roi",
                "tests/data/prob/D20180712T065600_IFCB114.prob.csv", "tests/model/thresholds-2021.txt",
                "tests/model/thresholds-zero.txt", "tests/model/resnet18_20201022/class_names.txt",
                "tests/model/resnet18_20201022/config.ini"):
        shutil.copy(REF / rel, fx / Path(rel).name)
        (fx / Path(rel).name).chmod(0o644)
    # geometry of the one real-world .adc (its .roi is a missing blob)
    rows = []
    with open(REF / "tests/data/raw/invalid/D20210523T053149_IFCB114.adc") as fh:
        for line in fh:
            f = line.split(",")
            rows.append((int(f[15]), int(f[16]), int(f[17])))
    np.savez_compressed(HERE / "invalid_adc_geometry.npz", whs=np.array(rows, np.int64))

    # labels of the reference's own golden CSV (real checkpoint) under its own thresholds
    real = {}
    for tname in ("thresholds-2021", "thresholds-zero"):
        df = prediction.prediction_dataframe(fx / "D20180712T065600_IFCB114.prob.csv", fx / f"{tname}.txt")
        counts = classification.class_df_probs_only([fx / "D20180712T065600_IFCB114.prob.csv"], fx / f"{tname}.txt")
        real[tname] = {"prediction": [str(p) for p in df["prediction"]], "classified": [bool(c) for c in df["classified"]],
                       "roi": [int(i) for i in df.index], "counts": {k: int(v) for k, v in counts.iloc[0].items()}}
    (HERE / "ref_fixture_labels.json").write_text(json.dumps(real, indent=1))

    classes = (fx / "class_names.txt").read_text().splitlines()
    thr = {"thresholds-2021": fx / "thresholds-2021.txt", "thresholds-zero": fx / "thresholds-zero.txt", "scalar-0.5": 0.5}
    for name in wanted:
        c = ALL_CASES[name]
        run_case(name, c["arch"], c["t"], c["border"], c["norm"], case_bins(name), classes, thr,
                 tap_limit=c["tap_limit"], seed=c["seed"], gain=c.get("gain", LOGIT_GAIN), source=c.get("source", "reference"), res_gamma=c.get("res_gamma", 1.0))


if __name__ == "__main__":
    main()
