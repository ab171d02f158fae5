"""K1 (ROI decode + mode border + cv2-exact resize + pad + ToTensor) on the GPU, through the C ABI,
against the reference's goldens and the oracle.  Bit-exact: integer work, and the fp32 tensor is a LUT."""

import hashlib

import numpy as np
import pytest

from oracle import ifcb as o_ifcb
from oracle import preprocess as o_pre
from sykepic_b200 import _lib, synth
from tests.cases import CASES, GOLDEN, case_bins
from tests.gpu_util import RawCtx, bin_arrays

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    c = RawCtx()
    yield c
    c.close()


ALL_BINS = [(c, i) for c in CASES for i in range(len(CASES[c]["bins"]))]


@pytest.mark.parametrize("case,idx", ALL_BINS)
def test_matches_reference_tensors(ctx, case, idx):
    """u8 taps and sha256 of the fp32 [3,T,T] tensor the REFERENCE fed its network, ROI by ROI."""
    c = CASES[case]
    bname, b = case_bins(case)[idx]
    g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
    rid, w, h, start, roi = bin_arrays(b)
    assert rid.tolist() == g["roi_id"].tolist()
    u8 = ctx.preprocess(roi, w, h, start, c["t"], c["border"])
    for k in range(len(g["u8_taps"])):
        assert np.array_equal(u8[k], g["u8_taps"][k]), (int(rid[k]), int(w[k]), int(h[k]))
    f32 = ctx.preprocess(roi, w, h, start, c["t"], c["border"], _lib.DTYPE_F32, 3, _lib.LAYOUT_NCHW)
    for k in range(len(rid)):
        assert sha(f32[k]) == str(g["f32_sha"][k]), (int(rid[k]), int(w[k]), int(h[k]))
    assert ctx.fault_count() == 0


@pytest.mark.parametrize("case,idx", ALL_BINS)
def test_matches_oracle_every_roi_and_layout(ctx, case, idx):
    c = CASES[case]
    bname, b = case_bins(case)[idx]
    rid, w, h, start, roi = bin_arrays(b)
    rows = o_ifcb.parse_adc_text(b["adc_text"])
    want = np.stack([o_pre.resize_with_border_u8(img, c["t"], c["t"], c["border"]) for _, img in o_ifcb.decode_rois(rows, roi)])
    got = ctx.preprocess(roi, w, h, start, c["t"], c["border"])
    assert np.array_equal(got, want)
    # NHWC fp32 with a per-channel LUT (ImageNet normalisation), and bf16
    lut = o_pre.to_tensor_lut(True)
    nhwc = ctx.preprocess(roi, w, h, start, c["t"], c["border"], _lib.DTYPE_F32, 3, _lib.LAYOUT_NHWC, lut)
    assert np.array_equal(nhwc, np.stack([lut[ch][want] for ch in range(3)], axis=-1))
    import torch

    bf = ctx.preprocess(roi, w, h, start, c["t"], c["border"], _lib.DTYPE_BF16, 3, _lib.LAYOUT_NCHW)
    lut0 = o_pre.to_tensor_lut(False)[0]
    want_bf = torch.from_numpy(lut0[want]).to(torch.bfloat16).float().numpy()
    assert np.array_equal(bf, np.repeat(want_bf[:, None], 3, axis=1))
    one = ctx.preprocess(roi, w, h, start, c["t"], c["border"], _lib.DTYPE_F32, 1, _lib.LAYOUT_NHWC)
    assert np.array_equal(one[..., 0], lut0[want])


@pytest.mark.parametrize("t,border", [(224, "mode"), (180, "white"), (299, "black")])
def test_full_size_bin(ctx, t, border):
    """A full synthetic bin (~5000 ROIs of IFCB geometry): every ROI against the oracle's border value
    and padding geometry, a 400-ROI sample (plus the largest ROIs) against the full oracle transform."""
    b = synth.synth_bin(1000)
    rid, w, h, start, roi = bin_arrays(b)
    n = len(rid)
    assert n > 4000
    got = ctx.preprocess(roi, w, h, start, t, border)
    rng = np.random.default_rng(5)
    pick = set(rng.choice(n, 400, replace=False).tolist()) | set(np.argsort(-(w.astype(np.int64) * h))[:12].tolist())
    for k in range(n):
        img = roi[start[k]:start[k] + int(w[k]) * int(h[k])].reshape(int(h[k]), int(w[k]))
        fill = o_pre.border_value(img, border)
        nh, nw = o_pre.get_new_dims(int(h[k]), int(w[k]), t, t)
        top, left = (t - nh) // 2, (t - nw) // 2
        out = got[k]
        assert (out[:top] == fill).all() and (out[top + nh:] == fill).all(), k
        assert (out[:, :left] == fill).all() and (out[:, left + nw:] == fill).all(), k
        if k in pick:
            assert np.array_equal(out, o_pre.resize_with_border_u8(img, t, t, border)), (k, int(w[k]), int(h[k]))
    assert ctx.fault_count() == 0


def test_identity_decode_round_trip(ctx):
    """ROIs that already have the target size come back byte for byte (decode + offsets, at scale)."""
    t = 224
    rng = np.random.default_rng(11)
    n = 600
    roi = rng.integers(0, 256, n * t * t, dtype=np.uint8)
    w = np.full(n, t, np.int32)
    h = np.full(n, t, np.int32)
    start = (np.arange(n, dtype=np.int64) * t * t)
    perm = rng.permutation(n)  # descriptors need not be in stream order
    got = ctx.preprocess(roi, w, h, start[perm], t, "mode")
    assert np.array_equal(got, roi.reshape(n, t, t)[perm])


def test_invalid_geometry_is_counted_not_read(ctx):
    before = ctx.fault_count()
    roi = np.arange(100, dtype=np.uint8)
    w = np.array([10, 10, 1000, 0], np.int32)
    h = np.array([10, 11, 2, 5], np.int32)
    start = np.array([0, 0, 0, 0], np.int64)
    got = ctx.preprocess(roi, w, h, start, 64, "white")
    assert ctx.fault_count() - before == 3
    assert np.array_equal(got[0], o_pre.resize_with_border_u8(roi.reshape(10, 10), 64, 64, "white"))
    assert (got[1] == 255).all() and (got[3] == 255).all()
