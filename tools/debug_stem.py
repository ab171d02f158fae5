"""Debug helper: where does the fused stem differ from the CUDA-core path? (prints the error map)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sykepic_b200 import _lib
from tests.gpu_util import RawCtx

BN_EPS = 1e-5
t, n = int(sys.argv[1]), int(sys.argv[2])
import torch

ctx = RawCtx()
rng = np.random.default_rng(t)
img = rng.integers(0, 256, (n, t, t), dtype=np.uint8)
w = (rng.standard_normal((64, 3, 7, 7)) * 0.08).astype(np.float32)
gamma = rng.uniform(0.5, 1.5, 64).astype(np.float32)
beta = (0.3 * rng.standard_normal(64)).astype(np.float32)
mean = (0.2 * rng.standard_normal(64)).astype(np.float32)
var = rng.uniform(0.5, 1.5, 64).astype(np.float32)
lib = ctx.lib


def run(impl):
    ctx.ck(lib.spk_net_begin(ctx.ctx, t, t, 1, _lib.PRECISION_BF16, n))
    ctx.ck(lib.spk_net_conv(ctx.ctx, 0, 0, 1, 0, -1, w.ctypes.data, 64, 3, 7, 7, 2, 3, gamma.ctypes.data, beta.ctypes.data,
                            mean.ctypes.data, var.ctypes.data, BN_EPS, None, 1, impl))
    ctx.ck(lib.spk_net_maxpool(ctx.ctx, 1, 2, 3, 2, 1))
    hw = np.zeros((4, 64), np.float32)
    hb = np.zeros(4, np.float32)
    ctx.ck(lib.spk_net_head(ctx.ctx, 2, 1, (C.c_void_p * 1)(hw.ctypes.data), (C.c_void_p * 1)(hb.ctypes.data), (C.c_int * 2)(64, 4)))
    ctx.ck(lib.spk_net_end(ctx.ctx))
    with torch.cuda.device(ctx.device), torch.cuda.stream(ctx.stream):
        x = torch.from_numpy(img).to(ctx.device)
        probs = torch.empty((n, 4), dtype=torch.float32, device=ctx.device)
        ctx.ck(lib.spk_forward(ctx.ctx, x.data_ptr(), n, 0.0, None, probs.data_ptr(), None, None))
        ctx.sync()
    h, ww, c = C.c_int(), C.c_int(), C.c_int()
    ctx.ck(lib.spk_net_read_buffer(ctx.ctx, 2, n, None, 0, C.byref(h), C.byref(ww), C.byref(c)))
    out = np.empty((n, h.value, ww.value, c.value), np.float32)
    ctx.ck(lib.spk_net_read_buffer(ctx.ctx, 2, n, out.ctypes.data, out.size, C.byref(h), C.byref(ww), C.byref(c)))
    return out


for rep in range(3):
    fused = run(_lib.CONV_AUTO)
    ref = run(_lib.CONV_SIMT)
    d = np.abs(fused - ref)
    print("  differing elements %.4f%%, mean |d| %.3e, max |d| %.3e (ref mean |y| %.3f)" % (100.0 * (d > 0).mean(), d.mean(), d.max(), np.abs(ref).mean()))
    bad = d > 0.05
    print("rep", rep, "bad elements", int(bad.sum()), "of", bad.size)
    if bad.any():
        im, r, c, ch = np.nonzero(bad)
        print(" images", np.unique(im), "rows", np.unique(r), "cols", np.unique(c)[:40], "channels", np.unique(ch)[:70])
        for i in np.unique(im)[:2]:
            m = bad[i].any(axis=2)
            print(" image", i, "bad rows -> cols:", {int(rr): np.flatnonzero(m[rr]).tolist()[:12] for rr in np.flatnonzero(m.any(axis=1))[:8]})
