"""CPU emulation of the bf16x3 split convolution: does it meet |dp| <= 1e-4 vs the fp32 forward?"""
import re, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import network as onet
from sykepic_b200 import synth

def split(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo

def q16(x):  # what a hi+lo pair stores
    hi, lo = split(x)
    return hi + lo

def fold(sd, w, bn):
    s = sd[bn + '.weight'].double() / torch.sqrt(sd[bn + '.running_var'].double() + 1e-5)
    return (w.double() * s[:, None, None, None]).float(), (sd[bn + '.bias'].double() - sd[bn + '.running_mean'].double() * s).float()

def conv3(x, w, b, **kw):
    xh, xl = split(x); wh, wl = split(w)
    y = F.conv2d(xh.double(), wh.double(), **kw) + F.conv2d(xh.double(), wl.double(), **kw) + F.conv2d(xl.double(), wh.double(), **kw)
    return y.float() + b[None, :, None, None]

def fwd(sd, x, mode):
    sd = {k: torch.as_tensor(v).float() for k, v in sd.items() if np.issubdtype(np.asarray(v).dtype, np.floating)}
    C = (lambda x, w, b, **kw: conv3(x, w, b, **kw)) if mode == 'split' else (lambda x, w, b, **kw: F.conv2d(x, w, b, **kw))
    st = (lambda t: q16(t)) if mode == 'split' else (lambda t: t)
    w, b = fold(sd, sd['base.0.weight'], 'base.1')
    x = st(F.max_pool2d(F.relu(C(x, w, b, stride=2, padding=3)), 3, 2, 1))
    for stage in (4, 5, 6, 7):
        blocks = sorted({int(m.group(1)) for k in sd if (m := re.match(rf"base\.{stage}\.(\d+)\.", k))})
        for bi in blocks:
            p = f"base.{stage}.{bi}"; stride = 2 if (stage > 4 and bi == 0) else 1
            idn = x
            w1, b1 = fold(sd, sd[p + '.conv1.weight'], p + '.bn1')
            t = st(F.relu(C(x, w1, b1, stride=stride, padding=1)))
            if (p + '.downsample.0.weight') in sd:
                wd, bd = fold(sd, sd[p + '.downsample.0.weight'], p + '.downsample.1')
                idn = st(C(x, wd, bd, stride=stride))
            w2, b2 = fold(sd, sd[p + '.conv2.weight'], p + '.bn2')
            x = st(F.relu(C(t, w2, b2, padding=1) + idn))
    feat = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return onet.probabilities(onet.head_logits(sd, feat))

rng = np.random.default_rng(0)
for gain in (8.0, 30.0):
    sd = synth.synth_state_dict('resnet18', 50, (256, 128), seed=0, randomize_bn=True, logit_gain=gain)
    b = synth.synth_bin(1234, 40)
    from oracle import preprocess as opre
    keep = np.flatnonzero(b['w'] > 0)[:24]
    imgs = [b['roi_bytes'][b['start'][k]:b['start'][k] + int(b['w'][k]) * int(b['h'][k])].reshape(int(b['h'][k]), int(b['w'][k])) for k in keep]
    u8 = np.stack([opre.resize_with_border_u8(im, 224, 224, 'mode') for im in imgs])
    x = torch.from_numpy(u8.astype(np.float32) / np.float32(255.0))[:, None].repeat(1, 3, 1, 1)
    with torch.no_grad():
        p32 = fwd(sd, x, 'fp32'); ps = fwd(sd, x, 'split')
        pref = onet.probabilities(onet.forward_logits({k: torch.as_tensor(v) for k, v in sd.items()}, x))
    print('gain', gain, 'max p', float(pref.max()), 'fold-vs-ref', float((p32 - pref).abs().max()), 'split-vs-ref', float((ps - pref).abs().max()))
