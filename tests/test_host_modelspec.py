"""`ModelSpec.from_dir`: what the reference's `prepare_model` reads from a model directory (sykepic/compute/probability.py:
118-130, sykepic/train/config.py:25-77): class names, `[image] shape / border`, `[model] network`, the state dict -- from
the reference's own test configuration (tests/model/resnet18_20201022/config.ini, committed as a fixture)."""
import shutil

import pytest
import torch

from sykepic_b200 import engine
from tests.cases import FIXTURE, fixture_classes


def _model_dir(tmp_path, edit=None):
    d = tmp_path / "model"
    d.mkdir()
    text = (FIXTURE / "config.ini").read_text()
    if edit:
        text = edit(text)
    (d / "config.ini").write_text(text)
    shutil.copy(FIXTURE / "class_names.txt", d / "class_names.txt")
    torch.save({"head.0.weight": torch.zeros(2, 3), "base.1.num_batches_tracked": torch.tensor(7)}, d / "best_state.pth")
    return d


def test_reference_configuration(tmp_path):
    spec = engine.ModelSpec.from_dir(_model_dir(tmp_path))
    assert spec.classes == fixture_classes() and len(spec.classes) == 50
    assert spec.img_shape == (3, 180, 180) and spec.border == "mode" and spec.arch == "resnet18"
    assert spec.imagenet_normalization is False
    assert set(spec.state_dict) == {"head.0.weight", "base.1.num_batches_tracked"}


def test_variants_and_errors(tmp_path):
    d = _model_dir(tmp_path, lambda t: t.replace("border = mode", "border = white").replace("imagenet_normalization = no", "imagenet_normalization = yes")
                   .replace("shape = 3, 180, 180", "shape = 3,224,224").replace("network = resnet18", "network = resnet50"))
    spec = engine.ModelSpec.from_dir(d)
    assert (spec.border, spec.imagenet_normalization, spec.img_shape, spec.arch) == ("white", True, (3, 224, 224), "resnet50")
    # a `weights` key (which makes the reference download ImageNet weights first, config.py:65-70) changes nothing here
    (d / "config.ini").write_text((d / "config.ini").read_text().replace("[model]", "[model]\nweights = DEFAULT"))
    assert engine.ModelSpec.from_dir(d).arch == "resnet50"
    # unknown border: the reference leaves the attribute unset and fails at the first image (image.py:20-28)
    (d / "config.ini").write_text((d / "config.ini").read_text().replace("border = white", "border = pink"))
    with pytest.raises(ValueError, match="border"):
        engine.ModelSpec.from_dir(d)
    # a pickled module instead of a state dict is refused (weights_only load)
    shutil.rmtree(d)
    d = _model_dir(tmp_path)
    (d / "best_state.pth").write_bytes(b"not a checkpoint")
    with pytest.raises(Exception):
        engine.ModelSpec.from_dir(d)
