"""Accuracy of one FP32-mode convolution: tensor-core split path and CUDA-core path against float64 on the same inputs."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F

from sykepic_b200 import _lib
from tests.gpu_util import RawCtx

ctx = RawCtx()
lib = ctx.lib
BN_EPS = 1e-5


def run(img, w1, w2, impl, n, t, relu=False):
    ctx.ck(lib.spk_net_begin(ctx.ctx, t, t, 1, _lib.PRECISION_FP32_TC, n))
    c1, c2, k = w1.shape[0], w2.shape[0], w2.shape[2]
    ctx.ck(lib.spk_net_conv(ctx.ctx, 0, 0, 1, 0, -1, w1.ctypes.data, c1, 1, 3, 3, 1, 1, None, None, None, None, BN_EPS, None, 1, _lib.CONV_SIMT))
    bias = np.zeros(c2, np.float32)
    ctx.ck(lib.spk_net_conv(ctx.ctx, 1, 0, 2, 0, -1, w2.ctypes.data, c2, c1, k, k, 1, k // 2, None, None, None, None, BN_EPS, bias.ctypes.data, int(relu), impl))
    hw = np.zeros((4, c2), np.float32); hb = np.zeros(4, np.float32)
    ctx.ck(lib.spk_net_head(ctx.ctx, 2, 1, (C.c_void_p * 1)(hw.ctypes.data), (C.c_void_p * 1)(hb.ctypes.data), (C.c_int * 2)(c2, 4)))
    ctx.ck(lib.spk_net_end(ctx.ctx))
    with torch.cuda.device(ctx.device), torch.cuda.stream(ctx.stream):
        x = torch.from_numpy(img).to(ctx.device)
        probs = torch.empty((n, 4), dtype=torch.float32, device=ctx.device)
        ctx.ck(lib.spk_forward(ctx.ctx, x.data_ptr(), n, 0.0, None, probs.data_ptr(), None, None))
        ctx.sync()
    outs = []
    for buf in (1, 2):
        h, ww, c = C.c_int(), C.c_int(), C.c_int()
        ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, None, 0, C.byref(h), C.byref(ww), C.byref(c)))
        out = np.empty((n, h.value, ww.value, c.value), np.float32)
        ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, out.ctypes.data, out.size, C.byref(h), C.byref(ww), C.byref(c)))
        outs.append(out)
    return outs


rng = np.random.default_rng(0)
for c1, c2, k, t in ((64, 64, 3, 28), (256, 256, 3, 14), (512, 512, 3, 7), (512, 128, 1, 14)):
    n = 4
    img = rng.integers(0, 256, (n, t, t), dtype=np.uint8)
    w1 = (rng.standard_normal((c1, 1, 3, 3)) * 0.6).astype(np.float32)
    w2 = (rng.standard_normal((c2, c1, k, k)) * np.sqrt(2.0 / (c1 * k * k))).astype(np.float32)
    a_tc, y_tc = run(img, w1, w2, _lib.CONV_AUTO, n, t)
    a_si, y_si = run(img, w1, w2, _lib.CONV_SIMT, n, t)
    ref = F.conv2d(torch.from_numpy(a_tc).double().permute(0, 3, 1, 2), torch.from_numpy(w2).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    rms = np.sqrt((ref ** 2).mean())
    print(f"{c1}->{c2} k{k} t{t}: same input {np.array_equal(a_tc, a_si)}; rel err (max|d| / rms) tc {np.abs(y_tc - ref).max() / rms:.2e} (mean signed {(y_tc - ref).mean() / rms:+.2e}), "
          f"simt {np.abs(y_si - ref).max() / rms:.2e}; rms err tc {np.sqrt(((y_tc - ref) ** 2).mean()) / rms:.2e}")
