"""Oracle: eval-time image transform of the reference, restated in numpy.

Test infrastructure only (see oracle/__init__.py).

Reference call chain: `ImageDataset.__getitem__` sykepic/train/data.py:210-231
-> `Compose.__call__` sykepic/train/image.py:25-56 -> `mode_pixel_value`
:229-237, `get_new_dims` :183-198, `resize_with_border` :201-226 ->
torchvision `ToTensor` / `Normalize` (wired in sykepic/train/config.py:52-58).

The arithmetic of `cv2.resize(..., INTER_LINEAR)` on uint8 lives in OpenCV
(third-party; reference pins opencv-python-headless==4.5.5.64,
requirements/cpu.txt:146).  It is restated here from OpenCV's published
algorithm (modules/imgproc/src/resize.cpp: `resizeGeneric_` with
`HResizeLinear` / `VResizeLinear`, INTER_RESIZE_COEF_BITS = 11, and the
INTER_LINEAR -> INTER_AREA switch for exact 2x decimation) and pinned
bit-for-bit against the cv2 in the build container by
tests/test_oracle_preprocess.py (when cv2 is importable) and by the golden
tensors the reference produced (tests/golden/).
"""

import numpy as np

INTER_BITS = 11
INTER_SCALE = 1 << INTER_BITS  # 2048

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def mode_pixel_value(img):
    """sykepic/train/image.py:229-237: 256-bin histogram, argmax (lowest wins ties)."""
    hist = np.bincount(np.asarray(img, dtype=np.uint8).ravel(), minlength=256)
    return int(np.argmax(hist))


def get_new_dims(h, w, target_h, target_w):
    """sykepic/train/image.py:183-198 (Python float64 arithmetic, int() truncation)."""
    if h > w:
        r = target_h / float(h)
        new_h = target_h
        new_w = int(w * r)
    else:
        r = target_w / float(w)
        new_h = int(h * r)
        new_w = target_w
    return new_h, new_w


def _rint_f32_to_i16(x):
    # cv::saturate_cast<short>(float) == cvRound == round-half-to-even
    return np.clip(np.rint(x), -32768, 32767).astype(np.int32)


def linear_coeffs(src, dst, horizontal):
    """Per-destination-index source tap and 11-bit fixed-point weights.

    OpenCV resize.cpp (cv::resize, INTER_LINEAR branch): scale = 1/(dst/src) in
    double; f = float((d+0.5)*scale-0.5); s = floor(f); f -= s (float32).
    Horizontal taps are clamped with the weight zeroed; vertical taps keep the
    weights and clamp the ROW INDICES instead (done by the caller).
    """
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        f = np.where(lo, np.float32(0), f)
        s = np.where(lo, 0, s)
        hi = s >= src - 1
        f = np.where(hi, np.float32(0), f)
        s = np.where(hi, src - 1, s)
    a0 = _rint_f32_to_i16((np.float32(1.0) - f) * np.float32(INTER_SCALE))
    a1 = _rint_f32_to_i16(f * np.float32(INTER_SCALE))
    return s, a0, a1


def resize_linear_u8(img, new_w, new_h):
    """`cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)` for 2-D uint8."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape
    if new_w <= 0 or new_h <= 0:
        # cv2.resize asserts !dsize.empty(); the reference lets cv2.error escape
        # and the bin is skipped (sykepic/compute/probability.py:113-114).
        raise ValueError(f"resize to empty size {new_w}x{new_h}")
    if (new_w, new_h) == (w, h):
        return img.copy()
    if w == 2 * new_w and h == 2 * new_h:
        # INTER_LINEAR silently becomes INTER_AREA (fast 2x2 box, rounded).
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = linear_coeffs(w, new_w, horizontal=True)
    sy, b0, b1 = linear_coeffs(h, new_h, horizontal=False)
    sx1 = np.minimum(sx + 1, w - 1)
    src = img.astype(np.int32)
    # horizontal pass for every source row (int32, scale 2^11)
    hbuf = src[:, sx] * a0[None, :] + src[:, sx1] * a1[None, :]
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    h0 = hbuf[y0] >> 4
    h1 = hbuf[y1] >> 4
    out = (((b0[:, None] * h0) >> 16) + ((b1[:, None] * h1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def border_value(img, border):
    """sykepic/train/image.py:20-28."""
    if border == "mode":
        return mode_pixel_value(img)
    if border == "white":
        return 255
    if border == "black":
        return 0
    raise ValueError(f"unknown border {border!r}")


def resize_with_border_u8(img, target_h, target_w, border="mode"):
    """(h,w) uint8 -> (target_h,target_w) uint8; image.py:25-34 + :201-226."""
    img = np.asarray(img, dtype=np.uint8)
    h, w = img.shape
    fill = border_value(img, border)
    new_h, new_w = get_new_dims(h, w, target_h, target_w)
    small = resize_linear_u8(img, new_w, new_h)
    pad_h = max(target_h - new_h, 0)
    pad_w = max(target_w - new_w, 0)
    top = pad_h // 2
    left = pad_w // 2
    out = np.full((new_h + pad_h, new_w + pad_w), fill, dtype=np.uint8)
    out[top : top + new_h, left : left + new_w] = small
    return out


def to_tensor_lut(imagenet_normalization):
    """3x256 fp32 table: ToTensor's `x.div(255)` (true fp32 division), then
    torchvision Normalize `(t - mean) / std` in fp32 (config.py:55-56)."""
    v = np.arange(256, dtype=np.float32) / np.float32(255)
    lut = np.stack([v, v, v]).astype(np.float32)
    if imagenet_normalization:
        mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)[:, None]
        std = np.asarray(IMAGENET_STD, dtype=np.float32)[:, None]
        lut = ((lut - mean) / std).astype(np.float32)
    return lut


def eval_transform(img, target_h, target_w, border="mode", imagenet_normalization=False, channels=3):
    """(h,w) uint8 ROI -> [channels,T_h,T_w] fp32, the tensor the reference feeds the net.

    channels == 3: cv2.imread of the gray PNG gives three identical planes
    (data.py:217-219).  channels == 1 (data.py:220-223) gives one plane.
    """
    u8 = resize_with_border_u8(img, target_h, target_w, border)
    lut = to_tensor_lut(imagenet_normalization)
    return np.stack([lut[c][u8] for c in range(channels)]).astype(np.float32)
