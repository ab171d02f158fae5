"""Synthetic IFCB bins and seeded random-init checkpoints.

Used by bench.py, the parity tests and tests/golden/make_golden.py.  Everything
is driven by numpy's PCG64 (`default_rng(seed)`), so the same seed gives the
same bytes on the build container and on the GPU box; no torch RNG, no
torchvision.

Bin geometry follows the one real-world .adc in the reference
(tests/data/raw/invalid/D20210523T053149_IFCB114.adc: width median 88 /
p99 199 / max 956, height median 50, widths multiples of 8, contiguous
monotone start bytes, a few 0x0 rows, duplicated trigger numbers).
The checkpoint layout is the state_dict `TorchVisionNet` produces
(sykepic/train/network.py:48-64): `base.<i>...` for the torchvision children,
`head.<i>.{weight,bias}` for the Linear chain.
"""

from collections import OrderedDict
from pathlib import Path

import numpy as np

ADC_FIELDS = 24

RESNET_SPECS = {
    # arch: (bottleneck, blocks per stage)
    "resnet18": (False, (2, 2, 2, 2)),
    "resnet34": (False, (3, 4, 6, 3)),
    "resnet50": (True, (3, 4, 6, 3)),
    "resnet101": (True, (3, 4, 23, 3)),
    "resnet152": (True, (3, 8, 36, 3)),
}

DENSENET_SPECS = {
    # arch: (growth, block config, init features, bn_size)
    "densenet121": (32, (6, 12, 24, 16), 64, 4),
    "densenet169": (32, (6, 12, 32, 32), 64, 4),
}


# --------------------------------------------------------------------------- geometry
def synth_geometry(rng, n_rows, empty_frac=0.0005, tail_frac=0.01):
    """-> (w[int32], h[int32]) per .adc row; empty rows have w = h = 0."""
    w = 8 * np.rint(rng.lognormal(np.log(88.0), 0.25, n_rows) / 8.0)
    h = 2 * np.rint(rng.lognormal(np.log(50.0), 0.30, n_rows) / 2.0)
    w = np.clip(w, 16, 1380).astype(np.int32)
    h = np.clip(h, 8, 1034).astype(np.int32)
    tail = rng.random(n_rows) < tail_frac
    w = np.where(tail, 8 * (rng.integers(200, 961, n_rows) // 8), w).astype(np.int32)
    h = np.where(tail, 2 * (rng.integers(100, 401, n_rows) // 2), h).astype(np.int32)
    empty = rng.random(n_rows) < empty_frac
    w = np.where(empty, 0, w).astype(np.int32)
    h = np.where(empty, 0, h).astype(np.int32)
    return w, h


def synth_roi_pixels(rng, w, h):
    """Background level + sigma=3 noise (well-defined mode) with a dark ellipse."""
    b = float(rng.integers(160, 211))
    img = b + 3.0 * rng.standard_normal((h, w), dtype=np.float32)
    cy, cx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w
    ry, rx = rng.uniform(0.15, 0.45) * h + 1.0, rng.uniform(0.15, 0.45) * w + 1.0
    yy, xx = np.ogrid[:h, :w]
    d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
    dark = float(rng.integers(20, 151))
    img = np.where(d < 1.0, dark + (b - dark) * d * d + 6.0 * rng.standard_normal((h, w), dtype=np.float32), img)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def adc_text(w, h, start, rng=None, newline="\r\n"):
    """24 comma-separated fields per row; width/height/start byte in columns 15/16/17."""
    n = len(w)
    trig = np.arange(1, n + 1)
    if rng is not None and n > 2:
        dup = rng.random(n) < 0.02  # duplicated trigger numbers must be ignored
        trig = np.where(dup, np.maximum(trig - 1, 1), trig)
    lines = []
    t = 0.047
    for i in range(n):
        t += 0.2 + (i % 7) * 0.013
        f = [str(int(trig[i])), f"{t:.6f}"] + ["0.00074"] * 8 + ["50.00000", f"{t:.6f}", f"{t + 2.059:.6f}"]
        f += [str(100 + (i * 37) % 900), str(50 + (i * 53) % 800)]
        f += [str(int(w[i])), str(int(h[i])), str(int(start[i]))]
        f += ["-999.000000", "0", "0", "0", f"{t + 2.0:.6f}", f"{0.75 * (i + 1):.6f}"]
        assert len(f) == ADC_FIELDS
        lines.append(",".join(f))
    return newline.join(lines) + (newline if lines else "")


def synth_bin(seed, n_rows=None, uniform_pixels=False):
    """-> dict(adc_text, roi_bytes uint8[...], w, h, start) for one synthetic bin.

    `uniform_pixels=True` gives the adversarial variant (uniform random bytes:
    mode ties, full dynamic range).
    """
    rng = np.random.default_rng(seed)
    if n_rows is None:
        n_rows = int(np.clip(np.rint(rng.normal(5000, 500)), 500, 9000))
    w, h = synth_geometry(rng, n_rows)
    area = w.astype(np.int64) * h.astype(np.int64)
    start = np.concatenate([[0], np.cumsum(area)[:-1]]).astype(np.int64)
    total = int(area.sum())
    roi = np.empty(total, dtype=np.uint8)
    if uniform_pixels:
        roi[:] = rng.integers(0, 256, total, dtype=np.uint8)
    else:
        for i in range(n_rows):
            if area[i]:
                roi[start[i] : start[i] + area[i]] = synth_roi_pixels(rng, int(w[i]), int(h[i])).ravel()
    return {"adc_text": adc_text(w, h, start, rng), "roi_bytes": roi, "w": w, "h": h, "start": start}


def bin_name(index, instrument=114):
    """D<YYYYMMDD>T<HHMMSS>_IFCB<nnn>, one bin every 20 minutes from 2021-05-23."""
    import datetime

    t = datetime.datetime(2021, 5, 23) + datetime.timedelta(minutes=20 * index)
    return t.strftime("D%Y%m%dT%H%M%S") + f"_IFCB{instrument}"


def write_bin(raw_dir, name, b):
    raw_dir = Path(raw_dir)
    raw_dir.mkdir(parents=True, exist_ok=True)
    with open(raw_dir / f"{name}.adc", "w", newline="") as fh:
        fh.write(b["adc_text"])
    np.asarray(b["roi_bytes"], dtype=np.uint8).tofile(raw_dir / f"{name}.roi")
    return raw_dir / name


# --------------------------------------------------------------------------- checkpoints
def _conv(rng, cout, cin, k, gain=1.0):
    std = gain * np.sqrt(2.0 / (cin * k * k))
    return (std * rng.standard_normal((cout, cin, k, k))).astype(np.float32)


def _bn(rng, sd, prefix, c, randomize):
    if randomize:
        sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    else:
        sd[prefix + ".weight"] = np.ones(c, np.float32)
        sd[prefix + ".bias"] = np.zeros(c, np.float32)
        sd[prefix + ".running_mean"] = np.zeros(c, np.float32)
        sd[prefix + ".running_var"] = np.ones(c, np.float32)
    sd[prefix + ".num_batches_tracked"] = np.asarray(0, dtype=np.int64)


def _head(rng, sd, dims, logit_gain):
    for i in range(len(dims) - 1):
        bound = 1.0 / np.sqrt(dims[i])
        g = logit_gain if i == len(dims) - 2 else 1.0
        sd[f"head.{i}.weight"] = (g * rng.uniform(-bound, bound, (dims[i + 1], dims[i]))).astype(np.float32)
        sd[f"head.{i}.bias"] = (g * rng.uniform(-bound, bound, dims[i + 1])).astype(np.float32)


def synth_state_dict(arch, n_classes=50, head=(256, 128), seed=0, randomize_bn=True, logit_gain=8.0, bn_stats=None,
                     res_gamma=1.0):
    """Seeded numpy state_dict with the key layout/shapes of `TorchVisionNet(arch, ...)`.

    `logit_gain` scales the last Linear so the softmax is peaky enough for the
    1e-4 / 2e-2 probability gates to mean something (default-init logits are
    ~1e-1 and every probability would be ~1/K).  `bn_stats` (a mapping, e.g. an
    npz under tests/golden/) overrides `*.running_mean` / `*.running_var` with
    statistics calibrated on synthetic ROIs, which makes the random network as
    input-sensitive as a trained one (tests/golden/make_golden.py).
    `res_gamma` scales the affine parameters of the LAST BatchNorm of every
    residual branch (ResNets): 1.0 leaves a branch as strong as the identity
    path, which makes a deep random network amplify any perturbation (a 0.2 %
    rounding becomes 13 % in ResNet-50's pooled features, tools/bf16_sweep.py);
    trained ResNets' branches are small corrections of the identity path
    (torchvision's zero_init_residual starts them at 0), which < 1 mimics.
    """
    rng = np.random.default_rng(10_000 + seed)
    sd = OrderedDict()
    if arch in RESNET_SPECS:
        bottleneck, layers = RESNET_SPECS[arch]
        exp = 4 if bottleneck else 1
        sd["base.0.weight"] = _conv(rng, 64, 3, 7)
        _bn(rng, sd, "base.1", 64, randomize_bn)
        inplanes = 64
        for si, (planes, nblk) in enumerate(zip((64, 128, 256, 512), layers)):
            for b in range(nblk):
                p = f"base.{4 + si}.{b}"
                stride = 2 if (si > 0 and b == 0) else 1
                # residual branches are damped so 8-50 blocks of random weights stay O(1)
                if bottleneck:
                    sd[p + ".conv1.weight"] = _conv(rng, planes, inplanes, 1)
                    _bn(rng, sd, p + ".bn1", planes, randomize_bn)
                    sd[p + ".conv2.weight"] = _conv(rng, planes, planes, 3)
                    _bn(rng, sd, p + ".bn2", planes, randomize_bn)
                    sd[p + ".conv3.weight"] = _conv(rng, planes * 4, planes, 1, gain=0.5)
                    _bn(rng, sd, p + ".bn3", planes * 4, randomize_bn)
                else:
                    sd[p + ".conv1.weight"] = _conv(rng, planes, inplanes, 3)
                    _bn(rng, sd, p + ".bn1", planes, randomize_bn)
                    sd[p + ".conv2.weight"] = _conv(rng, planes, planes, 3, gain=0.5)
                    _bn(rng, sd, p + ".bn2", planes, randomize_bn)
                if stride != 1 or inplanes != planes * exp:
                    sd[p + ".downsample.0.weight"] = _conv(rng, planes * exp, inplanes, 1, gain=0.7)
                    _bn(rng, sd, p + ".downsample.1", planes * exp, randomize_bn)
                inplanes = planes * exp
        feat = 512 * exp
    elif arch in DENSENET_SPECS:
        growth, cfg, c, bn_size = DENSENET_SPECS[arch]
        p = "base.0."
        sd[p + "conv0.weight"] = _conv(rng, c, 3, 7)
        _bn(rng, sd, p + "norm0", c, randomize_bn)
        for bi, nl in enumerate(cfg, start=1):
            for li in range(1, nl + 1):
                q = p + f"denseblock{bi}.denselayer{li}"
                _bn(rng, sd, q + ".norm1", c, randomize_bn)
                sd[q + ".conv1.weight"] = _conv(rng, bn_size * growth, c, 1)
                _bn(rng, sd, q + ".norm2", bn_size * growth, randomize_bn)
                sd[q + ".conv2.weight"] = _conv(rng, growth, bn_size * growth, 3)
                c += growth
            if bi != len(cfg):
                t = p + f"transition{bi}"
                _bn(rng, sd, t + ".norm", c, randomize_bn)
                sd[t + ".conv.weight"] = _conv(rng, c // 2, c, 1)
                c //= 2
        _bn(rng, sd, p + "norm5", c, randomize_bn)
        feat = c
    else:
        raise ValueError(f"unsupported network {arch!r}")
    _head(rng, sd, [feat, *head, n_classes], logit_gain)
    if res_gamma != 1.0 and arch in RESNET_SPECS:
        last = ".bn3." if RESNET_SPECS[arch][0] else ".bn2."
        for k in sd:
            if last in k and (k.endswith(".weight") or k.endswith(".bias")):
                sd[k] = (sd[k] * np.float32(res_gamma)).astype(np.float32)
    if bn_stats is not None:
        for k in bn_stats:
            if k in sd:
                assert sd[k].shape == bn_stats[k].shape, k
                sd[k] = np.asarray(bn_stats[k], dtype=np.float32)
    return sd


def class_names(n):
    return [f"Class_{i:02d}" for i in range(n)]


CONFIG_TEMPLATE = """[model]
network = {network}
weights =
head = {head}
dropout =

[image]
shape = 3, {t}, {t}
augmentations = flip, translate, zoom, brightness
imagenet_normalization = {norm}
border = {border}
zoom_range = 0.6, 1.4
brightness_range = 0.95, 1.1
max_rotation = 10
"""


def write_model_dir(model_dir, arch="resnet18", t=224, n_classes=50, head=(256, 128), seed=0, border="mode",
                    imagenet_normalization=False, randomize_bn=True, classes=None, logit_gain=8.0, bn_stats=None,
                    res_gamma=1.0):
    """Writes config.ini + class_names.txt + best_state.pth (torch.save of a state_dict)."""
    import torch

    model_dir = Path(model_dir)
    model_dir.mkdir(parents=True, exist_ok=True)
    classes = classes or class_names(n_classes)
    (model_dir / "class_names.txt").write_text("\n".join(classes) + "\n")
    (model_dir / "config.ini").write_text(
        CONFIG_TEMPLATE.format(network=arch, head=", ".join(str(h) for h in head), t=t,
                               norm="yes" if imagenet_normalization else "no", border=border)
    )
    sd = synth_state_dict(arch, len(classes), head, seed, randomize_bn, logit_gain, bn_stats, res_gamma)
    torch.save(OrderedDict((k, torch.from_numpy(np.asarray(v))) for k, v in sd.items()), model_dir / "best_state.pth")
    return model_dir
