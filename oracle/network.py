"""Oracle: the CNN forward of the reference, restated with torch-CPU fp32 functional ops.

Test infrastructure only (see oracle/__init__.py).  This is the "torch fp32
reference of the same op" for the floating-point kernels; it never runs on the
product path.

Reference: `TorchVisionNet` sykepic/train/network.py:11-72 builds
`base = Sequential(*torchvision_model.children()[:-1])` and an all-Linear
`head` (no activations; Dropout is identity in eval), `forward` :66-72 is
`head(base(x).view(N,-1))`.  The layer graph of `base` is torchvision's
(third-party; reference pins torchvision==0.12.0, requirements/cpu.txt:344):
ResNet = conv7x7/2 -> BN -> ReLU -> maxpool3x3/2 -> 4 stages of Basic/Bottleneck
blocks (stride on the 3x3 of the first block of stages 2-4, "v1.5") -> global
average pool.  It is restated here directly from the state_dict key layout
(`base.0.weight`, `base.1.*`, `base.4.0.conv1.weight`, ...,
`base.5.0.downsample.{0,1}.*`, `head.{i}.{weight,bias}`; SURVEY.md section 8a
A5) so the oracle does not depend on torchvision at run time, and pinned
against the reference's own module by tests/golden/make_golden.py.

DenseNet-121 (BASELINE config 4): the reference raises at 224x224 (SURVEY.md
8a A7); the behaviour restated here is torchvision's own forward
(features -> ReLU -> global average pool) followed by the syke-pic head,
state_dict keys `base.0.<features keys>` + `head.*` ("reference-undefined;
parity vs torchvision").

`net_pass` tail: sykepic/compute/probability.py:189-197 (logits * ln(1.3) in
fp32, softmax over classes).
"""

import re

import numpy as np
import torch
import torch.nn.functional as F

SOFTMAX_EXP = 1.3  # sykepic/compute/probability.py:18
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by torchvision resnet/densenet


def _bn(x, sd, prefix):
    return F.batch_norm(
        x,
        sd[prefix + ".running_mean"],
        sd[prefix + ".running_var"],
        sd[prefix + ".weight"],
        sd[prefix + ".bias"],
        training=False,
        eps=BN_EPS,
    )


def _resnet_block(x, sd, p, stride):
    bottleneck = (p + ".conv3.weight") in sd
    identity = x
    if bottleneck:
        out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
        out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
        out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    else:
        out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn1"))
        out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        identity = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + identity)


def resnet_features(sd, x):
    """torchvision ResNet children()[:-1] on NCHW fp32 -> [N, C_last]."""
    x = F.conv2d(x, sd["base.0.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, "base.1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for stage in (4, 5, 6, 7):
        blocks = sorted({int(m.group(1)) for k in sd if (m := re.match(rf"base\.{stage}\.(\d+)\.", k))})
        for b in blocks:
            stride = 2 if (stage > 4 and b == 0) else 1
            x = _resnet_block(x, sd, f"base.{stage}.{b}", stride)
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def densenet_features(sd, x):
    """torchvision DenseNet forward up to (not including) `classifier`."""
    p = "base.0."
    x = F.conv2d(x, sd[p + "conv0.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, p + "norm0"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    blk = 1
    while (p + f"denseblock{blk}.denselayer1.norm1.weight") in sd:
        layer = 1
        while (q := p + f"denseblock{blk}.denselayer{layer}") + ".norm1.weight" in sd:
            y = F.conv2d(F.relu(_bn(x, sd, q + ".norm1")), sd[q + ".conv1.weight"])
            y = F.conv2d(F.relu(_bn(y, sd, q + ".norm2")), sd[q + ".conv2.weight"], padding=1)
            x = torch.cat([x, y], 1)
            layer += 1
        t = p + f"transition{blk}"
        if (t + ".norm.weight") in sd:
            x = F.conv2d(F.relu(_bn(x, sd, t + ".norm")), sd[t + ".conv.weight"])
            x = F.avg_pool2d(x, kernel_size=2, stride=2)
        blk += 1
    x = F.relu(_bn(x, sd, p + "norm5"))
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def head_logits(sd, feat):
    """sykepic/train/network.py:56-63,68-69: chain of Linear layers, no activation."""
    idx = sorted({int(m.group(1)) for k in sd if (m := re.match(r"head\.(\d+)\.weight", k))})
    for i in idx:
        feat = F.linear(feat, sd[f"head.{i}.weight"], sd[f"head.{i}.bias"])
    return feat


def forward_logits(sd, x):
    """x: [N,3,T,T] fp32 torch CPU tensor -> logits [N,K] fp32."""
    with torch.no_grad():
        sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
        feats = densenet_features(sd, x) if "base.0.conv0.weight" in sd else resnet_features(sd, x)
        return head_logits(sd, feats)


def probabilities(logits):
    """sykepic/compute/probability.py:192-194; np.log(1.3) lands as an fp32 scalar."""
    with torch.no_grad():
        out = logits * np.log(SOFTMAX_EXP)
        return F.softmax(out, dim=1)
