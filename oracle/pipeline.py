"""Oracle: one IFCB bin -> .prob.csv text, end to end on the CPU.

Test infrastructure only (see oracle/__init__.py).

Follows sykepic/compute/probability.py: `prepare_model` :118-130,
`process_sample` :133-162 (without the PNG round trip, which is lossless),
`net_pass` :180-197, `probabilities_to_csv` :200-206.
"""

from configparser import ConfigParser
from pathlib import Path

import numpy as np
import torch

from . import ifcb, network, prediction, preprocess


class Model:
    def __init__(self, classes, img_shape, border, imagenet_normalization, arch, state_dict):
        self.classes = classes
        self.img_shape = img_shape
        self.border = border
        self.imagenet_normalization = imagenet_normalization
        self.arch = arch
        self.state_dict = state_dict


def prepare_model(model_dir):
    """probability.py:118-130 + config.py:20-29,55,63-77 (eval subset)."""
    model_dir = Path(model_dir)
    with open(model_dir / "class_names.txt") as fh:
        classes = fh.read().splitlines()
    config = ConfigParser()
    config.read(model_dir / "config.ini")
    img_shape = tuple(int(i) for i in config.get("image", "shape").split(","))
    border = config.get("image", "border")
    # sykepic/train/config.py:55-56 appends Normalize to the TRAIN transform only:
    # the eval transform the prob path uses never normalises, whatever the flag says.
    config.getboolean("image", "imagenet_normalization")
    norm = False
    arch = config.get("model", "network")
    sd = torch.load(model_dir / "best_state.pth", map_location="cpu")
    return Model(classes, img_shape, border, norm, arch, sd)


def preprocess_rois(model, images):
    c, th, tw = model.img_shape
    return np.stack(
        [preprocess.eval_transform(img, th, tw, model.border, model.imagenet_normalization, c) for img in images]
    )


def net_pass(model, rows, roi_bytes, batch_size=64):
    """-> sorted list of (roi_id, [float probs])."""
    results = []
    decoded = list(ifcb.decode_rois(rows, roi_bytes))  # ValueError here == "Faulty raw data"
    for i in range(0, len(decoded), batch_size):
        chunk = decoded[i : i + batch_size]
        x = torch.from_numpy(preprocess_rois(model, [img for _, img in chunk]))
        probs = network.probabilities(network.forward_logits(model.state_dict, x))
        results.extend(zip((rid for rid, _ in chunk), probs.tolist()))
    return sorted(results)


def process_bin(model, adc_path, roi_path, batch_size=64):
    """-> .prob.csv text for one bin."""
    rows = ifcb.parse_adc(adc_path)
    roi_bytes = np.fromfile(roi_path, dtype=np.uint8)
    return prediction.probabilities_to_csv_text(net_pass(model, rows, roi_bytes, batch_size), model.classes)
