"""Minimal PNG reader for `sykepic prob --image-dir/--images` (probability.py:28-36,165-177).

The reference reads ROI images with `cv2.imread` (train/data.py:217-219); IFCB ROI
PNGs are 8-bit grayscale, so the three planes it feeds the network are identical.
Here the file is decoded to ONE gray plane (zlib + the five PNG filters, numpy) and
takes the same device path as a raw bin.  8-bit gray / gray+alpha / RGB(A) with equal
channels, non-interlaced.
"""

import struct
import zlib

import numpy as np


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
    raw = raw.reshape(h, stride + 1)
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        f = int(raw[y, 0])
        line = raw[y, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        elif f == 1 and chans == 1:
            cur = np.cumsum(line) & 255
        else:  # Sub (multi-channel), Average, Paeth: sequential in x
            cur = np.zeros(stride, np.int32)
            for x in range(stride):
                a = cur[x - chans] if x >= chans else 0
                b = prev[x]
                c = prev[x - chans] if x >= chans else 0
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b) >> 1
                elif f == 4:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad filter {f}")
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
        prev = cur
    img = out.reshape(h, w, chans)
    if chans >= 3 and not (np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 0], img[..., 2])):
        raise ValueError(f"{path}: colour PNG; IFCB ROI images are grayscale")
    return np.ascontiguousarray(img[..., 0])
