"""BF16 probability error of the CUDA path against the oracle (torch-CPU fp32) for synthetic ResNet checkpoints as a function
of how "trained-like" they are: `res_gamma` scales the last BatchNorm of every residual branch (torchvision's
zero_init_residual is the limit 0), `gain` the last Linear (logit spread).  BatchNorm statistics are calibrated on synthetic
ROIs for every variant (train-mode pass, like tests/golden/make_golden.py `calibrate_bn`, restated functionally here).

    python tools/bf16_sweep.py resnet50:0.25:48 resnet50:0.1:64 ... [--rois 256]

Prints per variant: logit spread, relative bf16 error of the pooled features, max / median |dp| and the ROIs beyond 2e-2.
Also: per-ROI errors of the committed benchmark cases (--cases)."""
import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import ifcb, network, pipeline, preprocess  # noqa: E402
from sykepic_b200 import engine, synth  # noqa: E402


def calibrated_state_dict(arch, res_gamma, gain, seed=0, t=224):
    def build(stats=None):
        return synth.synth_state_dict(arch, 50, (256, 128), seed, True, gain, stats, res_gamma=res_gamma)

    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in build().items()}
    b = synth.synth_bin(500 + seed, 96)
    rows = ifcb.parse_adc_text(b["adc_text"])
    x = torch.from_numpy(np.stack([preprocess.eval_transform(img, t, t, "mode", False) for _, img in ifcb.decode_rois(rows, b["roi_bytes"])]))
    stats = {}

    def bn(y, p):
        m, v = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=True)
        stats[p + ".running_mean"], stats[p + ".running_var"] = m.numpy().copy(), v.numpy().copy()
        return F.batch_norm(y, None, None, sd[p + ".weight"], sd[p + ".bias"], True, eps=1e-5)

    with torch.no_grad():
        y = F.max_pool2d(F.relu(bn(F.conv2d(x, sd["base.0.weight"], stride=2, padding=3), "base.1")), 3, 2, 1)
        for stage in (4, 5, 6, 7):
            for bi in sorted({int(m.group(1)) for k in sd if (m := re.match(rf"base\.{stage}\.(\d+)\.", k))}):
                p, stride, idt = f"base.{stage}.{bi}", 2 if (stage > 4 and bi == 0) else 1, y
                if (p + ".conv3.weight") in sd:
                    o = F.relu(bn(F.conv2d(y, sd[p + ".conv1.weight"]), p + ".bn1"))
                    o = F.relu(bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=stride, padding=1), p + ".bn2"))
                    o = bn(F.conv2d(o, sd[p + ".conv3.weight"]), p + ".bn3")
                else:
                    o = F.relu(bn(F.conv2d(y, sd[p + ".conv1.weight"], stride=stride, padding=1), p + ".bn1"))
                    o = bn(F.conv2d(o, sd[p + ".conv2.weight"], padding=1), p + ".bn2")
                if (p + ".downsample.0.weight") in sd:
                    idt = bn(F.conv2d(y, sd[p + ".downsample.0.weight"], stride=stride), p + ".downsample.1")
                y = F.relu(o + idt)
    return build(stats)


def report(tag, probs, want, logits):
    err = np.abs(probs - want).max(axis=1)
    spread = logits.std(axis=1)
    print(f"{tag}: n {len(err)}  logit spread med {np.median(spread):.2f} max {spread.max():.1f}  pmax med {np.median(want.max(1)):.2f}  "
          f"|dp| max {err.max():.3e} p99 {np.percentile(err, 99):.3e} med {np.median(err):.3e}  >2e-2: {(err > 2e-2).sum()}", flush=True)
    return err


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(sys.argv[sys.argv.index("--rois") + 1]) if "--rois" in sys.argv else 256
    root = Path(tempfile.mkdtemp())
    torch.set_num_threads(max(1, torch.get_num_threads()))
    if "--cases" in sys.argv:
        from tests.cases import BIG_CASES, GOLDEN, case_bins, case_model_dir

        for case in BIG_CASES:
            (bname, b), = case_bins(case)
            g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
            eng = engine.Engine(case_model_dir(case, root), precision="bf16", max_batch=256)
            _, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
            eng.close()
            err = report(case, probs, g["probs"], g["logits"])
            np.save(f"gpurun_out/bf16_err_{case}.npy", err)
    b = synth.synth_bin(1101, n)
    rows = ifcb.parse_adc_text(b["adc_text"])
    for spec in args:
        arch, rg, gain = spec.split(":")
        sd = calibrated_state_dict(arch, float(rg), float(gain))
        mdir = root / spec.replace(":", "_")
        synth.write_model_dir(mdir, arch=arch, t=224, seed=0, border="mode", logit_gain=float(gain),
                              bn_stats={k: v for k, v in sd.items() if "running_" in k}, res_gamma=float(rg))
        model = pipeline.prepare_model(mdir)
        want = pipeline.net_pass(model, rows, np.asarray(b["roi_bytes"], np.uint8), batch_size=64)
        wp = np.array([p for _, p in want], np.float32)
        logits = np.log(np.maximum(wp, 1e-30)) / np.log(1.3)
        eng = engine.Engine(mdir, precision="bf16", max_batch=256)
        _, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
        eng.close()
        report(spec, probs, wp, logits)


if __name__ == "__main__":
    main()
