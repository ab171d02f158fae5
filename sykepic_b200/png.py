"""Minimal PNG reader for `sykepic prob --image-dir/--images` (probability.py:28-36,165-177).

The reference reads ROI images with `cv2.imread` (train/data.py:217-219); IFCB ROI
PNGs are 8-bit grayscale, so the three planes it feeds the network are identical.
Here the file is decoded to ONE gray plane (zlib inflate, then the five PNG filters in the C-ABI
host library) and takes the same device path as a raw bin.  8-bit gray / gray+alpha / RGB(A) with equal
channels, non-interlaced.
"""

import struct
import zlib

import numpy as np

from . import _lib


def read_gray(path):
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(data):
        ln, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
    if hdr is None:
        raise ValueError(f"{path}: no IHDR")
    w, h, depth, ctype, _, _, interlace = hdr
    chans = {0: 1, 2: 3, 4: 2, 6: 4}.get(ctype)
    if depth != 8 or chans is None or interlace:
        raise ValueError(f"{path}: unsupported PNG (depth {depth}, colour type {ctype}, interlace {interlace})")
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8)
    stride = w * chans
    if raw.size != h * (stride + 1):
        raise ValueError(f"{path}: {raw.size} bytes of image data, expected {h * (stride + 1)}")
    # the five scanline filters: sequential per byte, done by the C-ABI host library (spk_png_unfilter)
    out = np.empty((h, stride), np.uint8)
    rc = _lib.load().spk_png_unfilter(raw.ctypes.data, h, stride, chans, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"{path}: {_lib.last_error()}")
    img = out.reshape(h, w, chans)
    if chans >= 3 and not (np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 0], img[..., 2])):
        raise ValueError(f"{path}: colour PNG; IFCB ROI images are grayscale")
    return np.ascontiguousarray(img[..., 0])


def read_gray_many(paths, threads=0):
    """A sample's ROI images -> (w int32[n], h int32[n], start int64[n], data uint8[sum w*h]): the byte-stream layout of a
    `.roi` file, decoded by the C-ABI host library on `threads` threads (0 = all cores, at most 16).  Python threads do not
    help here (measured: 14.7 k ROI/s on one thread, 10 k on four -- the interpreter lock between the small numpy steps)."""
    import ctypes as C

    paths = [str(p) for p in paths]
    n = len(paths)
    w = np.empty(n, np.int32)
    h = np.empty(n, np.int32)
    if n == 0:
        return w, h, np.empty(0, np.int64), np.empty(0, np.uint8)
    lib = _lib.load()
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    bad = C.c_int64(-1)
    if lib.spk_png_probe(arr, n, w.ctypes.data, h.ctypes.data, threads, C.byref(bad)) != 0:
        raise ValueError(_lib.last_error())
    area = w.astype(np.int64) * h
    start = np.concatenate([[0], np.cumsum(area)[:-1]]).astype(np.int64)
    data = np.empty(int(area.sum()), np.uint8)
    if lib.spk_png_decode_batch(arr, n, w.ctypes.data, h.ctypes.data, start.ctypes.data, data.ctypes.data, data.size, threads, C.byref(bad)) != 0:
        raise ValueError(_lib.last_error())
    return w, h, start, data
