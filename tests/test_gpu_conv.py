"""K2 unit test: the tcgen05 implicit-GEMM convolution against the CUDA-core convolution and a torch
fp32 reference of the same op, layer geometry by layer geometry (3x3 / 1x1, stride 1 / 2, every Cout
tile width, odd image sizes that leave partial tiles, residual add, ReLU, batches that do not fill a
tile).  Both paths read the same bf16 activations, so they differ only by accumulation order."""

import ctypes as C

import numpy as np
import pytest

from sykepic_b200 import _lib
from tests.gpu_util import RawCtx

pytestmark = pytest.mark.gpu

BN_EPS = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = RawCtx()
    yield c
    c.close()


def _bf16(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).float().numpy()


def run_net(ctx, img, w1, w2, stride, pad, relu, residual, impl, n, t):
    """u8 image [n,t,t] -> conv1 (3x3/1, 1->c1, ReLU, CUDA cores) -> conv2 (the layer under test)
    [-> + residual branch conv3 1x1 of the same input].  Returns (conv1 out, conv2 out) as fp32 NHWC."""
    import torch

    lib = ctx.lib
    ctx.ck(lib.spk_net_begin(ctx.ctx, t, t, 1, _lib.PRECISION_BF16, n))
    c1 = w1.shape[0]
    c2 = w2.shape[0]
    ctx.ck(lib.spk_net_conv(ctx.ctx, 0, 0, 1, 0, -1, w1.ctypes.data, c1, 1, 3, 3, 1, 1, None, None, None, None, BN_EPS, None, 1, _lib.CONV_SIMT))
    res_buf = -1
    if residual is not None:
        # a second branch producing a tensor of conv2's output shape (1x1, same stride), added in conv2's epilogue
        ctx.ck(lib.spk_net_conv(ctx.ctx, 1, 0, 3, 0, -1, residual.ctypes.data, c2, c1, 1, 1, stride, 0, None, None, None, None, BN_EPS,
                                None, 0, _lib.CONV_SIMT))
        res_buf = 3
    k = w2.shape[2]
    bias = np.linspace(-0.5, 0.5, c2).astype(np.float32)
    ctx.ck(lib.spk_net_conv(ctx.ctx, 1, 0, 2, 0, res_buf, w2.ctypes.data, c2, c1, k, k, stride, pad, None, None, None, None, BN_EPS,
                            bias.ctypes.data, int(relu), impl))
    hw = np.zeros((4, c2), np.float32)
    hb = np.zeros(4, np.float32)
    wp = (C.c_void_p * 1)(hw.ctypes.data)
    bp = (C.c_void_p * 1)(hb.ctypes.data)
    dims = (C.c_int * 2)(c2, 4)
    ctx.ck(lib.spk_net_head(ctx.ctx, 2, 1, wp, bp, dims))
    ctx.ck(lib.spk_net_end(ctx.ctx))
    with torch.cuda.device(ctx.device), torch.cuda.stream(ctx.stream):
        x = torch.from_numpy(img).to(ctx.device)
        probs = torch.empty((n, 4), dtype=torch.float32, device=ctx.device)
        ctx.ck(lib.spk_forward(ctx.ctx, x.data_ptr(), n, 0.0, None, probs.data_ptr(), None, None))
        ctx.sync()

    def read(buf):
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, None, 0, C.byref(h), C.byref(w), C.byref(c)))
        out = np.empty((n, h.value, w.value, c.value), np.float32)
        ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, out.ctypes.data, out.size, C.byref(h), C.byref(w), C.byref(c)))
        return out

    return read(1), read(2), (read(3) if residual is not None else None), bias


GEOMS = [
    # (t, n, c1, c2, k, stride, pad, relu, residual)
    (56, 4, 64, 64, 3, 1, 1, True, False),      # ResNet-18 layer1 workhorse
    (56, 3, 64, 64, 3, 1, 1, True, True),       # + residual, batch that leaves a partial tile
    (28, 9, 64, 128, 3, 2, 1, True, False),     # stride-2 3x3 (four parity maps)
    (28, 5, 64, 128, 1, 2, 0, False, False),    # stride-2 1x1 downsample
    (14, 20, 128, 256, 3, 1, 1, True, True),    # two k chunks per tap, BN = 256
    (7, 70, 256, 512, 3, 1, 1, True, False),    # 7x7 maps, two N tiles
    (45, 3, 64, 64, 3, 1, 1, True, False),      # odd size (T = 180 path): partial tiles in both directions
    (23, 5, 64, 128, 3, 2, 1, True, True),      # odd size with stride 2
    (30, 2, 96, 32, 3, 1, 1, False, False),     # DenseNet-like: Cin not a multiple of 64, Cout = 32
    (16, 6, 160, 128, 1, 1, 0, True, False),    # DenseNet 1x1 with a 32-channel tail chunk
    # halo-resident 3x3 kernel (conv_halo.cu): streamed weights shared by two M tiles, odd tile counts, clipped last rows
    (28, 5, 128, 128, 3, 1, 1, True, True),     # ResNet-18 layer2: BN = 128, MT = 2, two k chunks, residual by TMA
    (28, 7, 128, 64, 3, 1, 1, False, False),    # BN = 64 streamed, 49 M tiles (odd: the last pair is half empty)
    (57, 3, 64, 64, 3, 1, 1, True, True),       # resident weights, H not a multiple of the tile rows (TMA store clips)
    (30, 3, 64, 128, 3, 1, 1, True, False),     # Cin 64 -> Cout 128, 4-row tiles with a 2-row remainder
    (28, 2, 64, 192, 3, 1, 1, True, True),      # three N tiles of 64
    (28, 3, 128, 32, 3, 1, 1, False, False),    # DenseNet growth convolution: halo-pair kernel with N = 32
    (56, 3, 128, 32, 3, 1, 1, True, False),     # ... on 56x56 maps (2-row tiles, odd tile count)
    (14, 5, 128, 32, 3, 1, 1, False, False),    # ... on 14x14 maps (8-row tiles, the second one clipped)
    (14, 6, 128, 128, 3, 1, 1, True, True),     # 14x14 with residual through the halo-pair kernel
]
HALO_GEOMS = [0, 1, 10, 12, 2, 4, 5, 9]  # halo / pair geometries also run through the tap-per-TMA kernel (SPK_CONV_TCGEN05_TAPS)


@pytest.mark.parametrize("gi", HALO_GEOMS, ids=[f"g{i}" for i in HALO_GEOMS])
def test_tap_kernel_on_halo_geometries(ctx, gi):
    test_tcgen05_conv_matches_cuda_cores_and_torch(ctx, *GEOMS[gi], impl_tc=_lib.CONV_TCGEN05_TAPS)


@pytest.mark.parametrize("t,n,c1,c2,k,stride,pad,relu,residual", GEOMS, ids=[f"g{i}" for i in range(len(GEOMS))])
def test_tcgen05_conv_matches_cuda_cores_and_torch(ctx, t, n, c1, c2, k, stride, pad, relu, residual, impl_tc=_lib.CONV_TCGEN05):
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(t * 1000 + c2 + k)
    img = rng.integers(0, 256, (n, t, t), dtype=np.uint8)
    w1 = (rng.standard_normal((c1, 1, 3, 3)) * 0.6).astype(np.float32)
    w2 = (rng.standard_normal((c2, c1, k, k)) * np.sqrt(2.0 / (c1 * k * k))).astype(np.float32)
    wr = (rng.standard_normal((c2, c1, 1, 1)) * np.sqrt(1.0 / c1)).astype(np.float32) if residual else None
    a_simt, y_simt, r_simt, bias = run_net(ctx, img, w1, w2, stride, pad, relu, wr, _lib.CONV_SIMT, n, t)
    a_tc, y_tc, r_tc, _ = run_net(ctx, img, w1, w2, stride, pad, relu, wr, impl_tc, n, t)
    assert np.array_equal(a_simt, a_tc)  # same conv1 on both paths
    # torch fp32 reference of conv2 on the SAME bf16 input and bf16-rounded weights
    x = torch.from_numpy(a_tc).permute(0, 3, 1, 2)
    ref = F.conv2d(x, torch.from_numpy(_bf16(w2)), torch.from_numpy(bias), stride=stride, padding=pad)
    if residual:
        ref = ref + torch.from_numpy(r_tc).permute(0, 3, 1, 2)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1).numpy()
    assert y_tc.shape == ref.shape == y_simt.shape
    scale = max(1.0, float(np.abs(ref).max()))
    # bf16 output rounding: half an ulp = 2^-9 relative
    err_tc = float(np.abs(y_tc - ref).max())
    assert err_tc <= 6e-3 * scale, (err_tc, scale)
    # the CUDA-core path keeps fp32 weights (not bf16-rounded): looser
    assert float(np.abs(y_simt - y_tc).max()) <= 3e-2 * scale


@pytest.mark.parametrize("t,n", [(224, 5), (180, 3), (64, 9), (90, 2), (256, 2), (30, 4), (224, 300), (96, 200), (176, 160)], ids=lambda v: str(v))
def test_fused_stem_matches_unfused_and_torch(ctx, t, n):
    """conv 7x7/2 (3 equal planes folded to 1) + BN + ReLU + max-pool 3x3/2: the fused tcgen05 kernel
    against the unfused CUDA-core path and a torch fp32 reference of the same ops."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(t)
    img = rng.integers(0, 256, (n, t, t), dtype=np.uint8)
    img[0] = 255
    img[-1, ::2] = 0
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

    fused = run(_lib.CONV_AUTO)
    unfused = run(_lib.CONV_SIMT)
    # torch reference with the same folding: w' = sum_c w * gamma / sqrt(var + eps) / 255 (the kernel carries it as
    # per-channel scaled fp16, i.e. to 2^-12; SPK_STEM=hilo / fp32_tc: bf16 hi + bf16 lo, 2^-17), integer pixels
    scale = gamma.astype(np.float64) / np.sqrt(var.astype(np.float64) + BN_EPS)
    wf = (w.astype(np.float64).sum(axis=1) * scale[:, None, None])
    bias = (beta - mean * scale).astype(np.float32)
    w_bf = (wf / 255.0).astype(np.float32)[:, None]
    x = torch.from_numpy(img.astype(np.float32))[:, None]
    ref = F.max_pool2d(F.relu(F.conv2d(x, torch.from_numpy(w_bf), torch.from_numpy(bias), stride=2, padding=3)), 3, 2, 1)
    ref = _bf16(ref.permute(0, 2, 3, 1).numpy())
    assert fused.shape == ref.shape == unfused.shape
    mag = max(1.0, float(np.abs(ref).max()))
    # bf16 output of fp16 weights (2^-12 relative): element-wise within ONE bf16 step of the rounded reference
    # (a value that sits next to a rounding boundary may land on the other side), and close on average
    d = np.abs(fused - ref)
    assert (d <= 2.0 ** -7 * np.abs(ref) + 1e-3 * mag).all(), float(d.max())
    assert float(d.mean()) <= 2e-4 * mag, float(d.mean())
    assert float((d > 0).mean()) <= 0.10, float((d > 0).mean())
    assert float(np.abs(fused - unfused).max()) <= 2e-2 * mag


@pytest.mark.parametrize("t,n,c1,c2", [(56, 5, 64, 128), (28, 9, 128, 256), (14, 33, 256, 512), (45, 3, 64, 128)], ids=lambda v: str(v))
def test_fused_downsample_shortcut(ctx, t, n, c1, c2):
    """A ResNet stage entry: the 1x1 / stride-2 shortcut fused into the 3x3 / stride-2 convolution's kernel (second
    accumulator on the centre tap) against the two separate launches and a torch fp32 reference."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(t + c2)
    img = rng.integers(0, 256, (n, t, t), dtype=np.uint8)
    w1 = (rng.standard_normal((c1, 1, 3, 3)) * 0.6).astype(np.float32)
    w3 = (rng.standard_normal((c2, c1, 3, 3)) * np.sqrt(2.0 / (c1 * 9))).astype(np.float32)
    wd = (rng.standard_normal((c2, c1, 1, 1)) * np.sqrt(1.0 / c1)).astype(np.float32)
    b3 = np.linspace(-0.5, 0.5, c2).astype(np.float32)
    bd = np.linspace(0.3, -0.3, c2).astype(np.float32)
    lib = ctx.lib

    def run(fused):
        # the fusion pass matches "shortcut declared directly before its 3x3" (the order engine.py emits); declaring the
        # 3x3 first keeps the two convolutions separate launches
        ctx.ck(lib.spk_net_begin(ctx.ctx, t, t, 1, _lib.PRECISION_BF16, n))
        ctx.ck(lib.spk_net_conv(ctx.ctx, 0, 0, 1, 0, -1, w1.ctypes.data, c1, 1, 3, 3, 1, 1, None, None, None, None, BN_EPS, None, 1, _lib.CONV_SIMT))

        def shortcut():
            ctx.ck(lib.spk_net_conv(ctx.ctx, 1, 0, 3, 0, -1, wd.ctypes.data, c2, c1, 1, 1, 2, 0, None, None, None, None, BN_EPS, bd.ctypes.data, 0, _lib.CONV_TCGEN05))

        def conv3():
            ctx.ck(lib.spk_net_conv(ctx.ctx, 1, 0, 2, 0, -1, w3.ctypes.data, c2, c1, 3, 3, 2, 1, None, None, None, None, BN_EPS, b3.ctypes.data, 1, _lib.CONV_TCGEN05))

        for op in ((shortcut, conv3) if fused else (conv3, shortcut)):
            op()
        hw = np.zeros((4, c2), np.float32)
        hb = np.zeros(4, np.float32)
        ctx.ck(lib.spk_net_head(ctx.ctx, 2, 1, (C.c_void_p * 1)(hw.ctypes.data), (C.c_void_p * 1)(hb.ctypes.data), (C.c_int * 2)(c2, 4)))
        ctx.ck(lib.spk_net_end(ctx.ctx))
        launches0 = lib.spk_launch_count(ctx.ctx)
        with torch.cuda.device(ctx.device), torch.cuda.stream(ctx.stream):
            x = torch.from_numpy(img).to(ctx.device)
            probs = torch.empty((n, 4), dtype=torch.float32, device=ctx.device)
            ctx.ck(lib.spk_forward(ctx.ctx, x.data_ptr(), n, 0.0, None, probs.data_ptr(), None, None))
            ctx.sync()
        outs = []
        for buf in (1, 2, 3):
            h, w, c = C.c_int(), C.c_int(), C.c_int()
            ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, None, 0, C.byref(h), C.byref(w), C.byref(c)))
            out = np.empty((n, h.value, w.value, c.value), np.float32)
            ctx.ck(lib.spk_net_read_buffer(ctx.ctx, buf, n, out.ctypes.data, out.size, C.byref(h), C.byref(w), C.byref(c)))
            outs.append(out)
        return outs, lib.spk_launch_count(ctx.ctx) - launches0

    (a_f, y_f, d_f), launches_f = run(True)
    (a_u, y_u, d_u), launches_u = run(False)
    assert launches_f == launches_u - 1  # one launch fewer
    assert np.array_equal(a_f, a_u)
    x = torch.from_numpy(a_f).permute(0, 3, 1, 2)
    ref3 = F.relu(F.conv2d(x, torch.from_numpy(_bf16(w3)), torch.from_numpy(b3), stride=2, padding=1)).permute(0, 2, 3, 1).numpy()
    refd = F.conv2d(x, torch.from_numpy(_bf16(wd)), torch.from_numpy(bd), stride=2).permute(0, 2, 3, 1).numpy()
    for got, sep, ref in ((y_f, y_u, ref3), (d_f, d_u, refd)):
        assert got.shape == ref.shape == sep.shape
        scale = max(1.0, float(np.abs(ref).max()))
        assert float(np.abs(got - ref).max()) <= 6e-3 * scale
        assert float(np.abs(got - sep).max()) <= 6e-3 * scale
