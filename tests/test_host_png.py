"""PNG decode of the image mode (`sykepic prob --image-dir/--images`): sykepic_b200/png.py + spk_png_unfilter (C ABI, host)
against cv2.imread, which is what the reference uses (sykepic/train/data.py:217-219), and against hand-filtered files
that force each of the five scanline filters."""
import struct
from pathlib import Path
import zlib

import numpy as np
import pytest

from sykepic_b200 import _lib, png, synth


def _chunk(typ, body):
    return struct.pack(">I", len(body)) + typ + body + struct.pack(">I", zlib.crc32(typ + body) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, img, filters, ctype=0):
    """Encode `img` ([h, w] or [h, w, chans] uint8) with the given filter type per row (cycled)."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape[:2]
    chans = 1 if img.ndim == 2 else img.shape[2]
    rows = img.reshape(h, w * chans).astype(np.int32)
    stride = w * chans
    raw = bytearray()
    for y in range(h):
        f = filters[y % len(filters)]
        cur, prev = rows[y], rows[y - 1] if y else np.zeros(stride, np.int32)
        line = np.empty(stride, np.int32)
        for x in range(stride):
            a = cur[x - chans] if x >= chans else 0
            b = prev[x]
            c = prev[x - chans] if x >= chans else 0
            pred = [0, a, b, (a + b) >> 1, _paeth(a, b, c)][f]
            line[x] = (cur[x] - pred) & 255
        raw.append(f)
        raw += line.astype(np.uint8).tobytes()
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    comp = zlib.compress(bytes(raw), 6)
    half = len(comp) // 2  # two IDAT chunks: the stream may be split anywhere
    data += _chunk(b"IDAT", comp[:half]) + _chunk(b"IDAT", comp[half:]) + _chunk(b"IEND", b"")
    path.write_bytes(data)


def write_png_up(path, img):
    """Fast encoder for whole bins of ROI images (numpy only): every row with the Up filter."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    diff = img.astype(np.int16)
    diff[1:] -= img[:-1].astype(np.int16)
    raw = np.empty((h, w + 1), np.uint8)
    raw[:, 0] = 2
    raw[:, 1:] = (diff & 255).astype(np.uint8)
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0))
    path.write_bytes(data + _chunk(b"IDAT", zlib.compress(raw.tobytes(), 1)) + _chunk(b"IEND", b""))


def test_fast_encoder_round_trip(tmp_path):
    rng = np.random.default_rng(9)
    for w, h in [(1, 1), (5, 1), (1, 7), (1380, 1034), (88, 50)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        write_png_up(tmp_path / "u.png", img)
        assert np.array_equal(png.read_gray(tmp_path / "u.png"), img)


@pytest.mark.parametrize("filters", [[0], [1], [2], [3], [4], [4, 3, 2, 1, 0], [1, 4]], ids=str)
def test_every_filter_type(tmp_path, filters):
    rng = np.random.default_rng(sum(filters) + len(filters))
    for w, h in [(1, 1), (2, 3), (17, 9), (88, 50)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        write_png(tmp_path / "g.png", img, filters)
        assert np.array_equal(png.read_gray(tmp_path / "g.png"), img)
        # gray replicated into RGB / RGBA / gray+alpha (what some tools write for the same picture)
        for ctype, chans in ((2, 3), (6, 4), (4, 2)):
            multi = np.repeat(img[..., None], chans, axis=2)
            if chans in (2, 4):
                multi[..., -1] = 255
            write_png(tmp_path / "m.png", multi, filters, ctype)
            assert np.array_equal(png.read_gray(tmp_path / "m.png"), img)


def test_matches_cv2_on_roi_images(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for k in range(40):
        w, h = int(rng.integers(8, 400)), int(rng.integers(4, 200))
        img = synth.synth_roi_pixels(rng, w, h) if k % 2 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        path = tmp_path / f"D20210523T000000_IFCB114_{k:05d}.png"
        assert cv2.imwrite(str(path), img)  # the reference writes ROI images with cv2 (utils/ifcb.py:118)
        ref = cv2.imread(str(path))  # colour read of a gray file: three identical planes (train/data.py:217-219)
        got = png.read_gray(path)
        assert np.array_equal(got, img) and np.array_equal(ref[..., 0], got) and np.array_equal(ref[..., 2], got)


def test_rejects_what_the_path_cannot_take(tmp_path):
    rng = np.random.default_rng(1)
    colour = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    write_png(tmp_path / "c.png", colour, [0], ctype=2)
    with pytest.raises(ValueError, match="colour"):
        png.read_gray(tmp_path / "c.png")
    (tmp_path / "n.png").write_bytes(b"not a png at all")
    with pytest.raises(ValueError, match="not a PNG"):
        png.read_gray(tmp_path / "n.png")
    # an unknown filter type comes back from the C ABI as SPK_ERR_PARSE
    raw = np.array([7, 1, 2, 3], np.uint8)
    out = np.empty(3, np.uint8)
    assert _lib.load().spk_png_unfilter(raw.ctypes.data, 1, 3, 1, out.ctypes.data) == _lib.SPK_ERR_PARSE
    # truncated image data
    w, h = 7, 5
    stream = b"".join(b"\x00" + bytes(range(w)) for _ in range(h))[:-3]
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0))
    (tmp_path / "s.png").write_bytes(data + _chunk(b"IDAT", zlib.compress(stream)) + _chunk(b"IEND", b""))
    with pytest.raises(ValueError, match="bytes of image data"):
        png.read_gray(tmp_path / "s.png")


def test_process_images_decodes_in_order(tmp_path):
    """`process_images` (probability.py:165-177 of the reference) with a stand-in engine: the PNGs of one sample are
    decoded on the thread pool, packed in the order given and the CSV comes out sorted by ROI id (:197)."""
    from collections import namedtuple

    from sykepic_b200.compute import probability

    rng = np.random.default_rng(3)
    sample = "D20210523T000000_IFCB114"
    ids = [7, 2, 11, 3, 5, 40, 1]
    imgs = {}
    for k, i in enumerate(ids):
        img = rng.integers(0, 256, (4 + k, 9 + 2 * k), dtype=np.uint8)
        imgs[i] = img
        write_png(tmp_path / f"{sample}_{i:05d}.png", img, [4, 1, 3])

    class Net:
        def run_rois(self, roi_id, w, h, start, data, batch_size=None):
            # "probabilities" that identify the picture behind each descriptor
            out = np.zeros((len(w), 2), np.float32)
            for k in range(len(w)):
                px = data[int(start[k]): int(start[k]) + int(w[k]) * int(h[k])].reshape(int(h[k]), int(w[k]))
                assert np.array_equal(px, imgs[int(roi_id[k])])
                out[k] = (float(px[0, 0]) / 255.0, float(px[-1, -1]) / 255.0)
            return out

    params = namedtuple("P", "classes batch_size")(["a", "b"], 4)
    csv_path = tmp_path / "out" / f"{sample}.prob.csv"
    probability.process_images([tmp_path / f"{sample}_{i:05d}.png" for i in ids], Net(), params, csv_path)
    lines = csv_path.read_text().splitlines()
    assert lines[0] == "roi,a,b" and [int(l.split(",")[0]) for l in lines[1:]] == sorted(ids)
    for l in lines[1:]:
        i, a, b = l.split(",")
        img = imgs[int(i)]
        assert a == f"{np.float32(img[0, 0] / np.float32(255.0)):.5f}" or abs(float(a) - img[0, 0] / 255.0) < 1e-5
        assert abs(float(b) - img[-1, -1] / 255.0) < 1e-5
    # existing file: skipped unless forced (probability.py:136-141)
    csv_path.write_text("sentinel")
    probability.process_images([tmp_path / f"{sample}_{ids[0]:05d}.png"], Net(), params, csv_path)
    assert csv_path.read_text() == "sentinel"
    probability.process_images([tmp_path / f"{sample}_{ids[0]:05d}.png"], Net(), params, csv_path, force=True)
    assert csv_path.read_text().startswith("roi,a,b\n7,")


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_batch_decoder_matches_per_file_decoder(tmp_path, threads):
    rng = np.random.default_rng(11)
    paths, imgs = [], []
    for k in range(60):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 120))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        p = tmp_path / f"S_{k:05d}.png"
        if k % 3 == 0:
            write_png_up(p, img)
        elif k % 3 == 1:
            write_png(p, img[: min(h, 12), : min(w, 40)], [4, 3, 1, 2, 0])
            img = img[: min(h, 12), : min(w, 40)]
        else:  # gray stored as RGB / RGBA
            c = 3 if k % 2 else 4
            small = img[: min(h, 10), : min(w, 30)]
            multi = np.repeat(small[..., None], c, axis=2)
            write_png(p, multi, [1, 4], 2 if c == 3 else 6)
            img = small
        paths.append(p)
        imgs.append(np.ascontiguousarray(img))
    w, h, start, data = png.read_gray_many(paths, threads)
    assert w.tolist() == [i.shape[1] for i in imgs] and h.tolist() == [i.shape[0] for i in imgs]
    assert start.tolist() == np.concatenate([[0], np.cumsum([i.size for i in imgs])[:-1]]).tolist()
    for k, img in enumerate(imgs):
        got = data[start[k]: start[k] + img.size].reshape(img.shape)
        assert np.array_equal(got, img) and np.array_equal(png.read_gray(paths[k]), img), k
    assert png.read_gray_many([])[3].size == 0


def test_batch_decoder_matches_cv2(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(12)
    paths = []
    for k in range(50):
        w, h = int(rng.integers(8, 500)), int(rng.integers(4, 260))
        img = synth.synth_roi_pixels(rng, w, h) if k % 2 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        p = tmp_path / f"D20210523T000000_IFCB114_{k:05d}.png"
        assert cv2.imwrite(str(p), img, [cv2.IMWRITE_PNG_COMPRESSION, int(k % 10)])
        paths.append(p)
    w, h, start, data = png.read_gray_many(paths)
    for k, p in enumerate(paths):
        ref = cv2.imread(str(p))
        assert ref.shape[:2] == (h[k], w[k])
        assert np.array_equal(data[start[k]: start[k] + int(w[k]) * int(h[k])].reshape(h[k], w[k]), ref[..., 1])


def test_batch_decoder_errors_name_the_file(tmp_path):
    rng = np.random.default_rng(13)
    good = []
    for k in range(4):
        p = tmp_path / f"S_{k:05d}.png"
        write_png_up(p, rng.integers(0, 256, (9, 14), dtype=np.uint8))
        good.append(p)
    # missing file
    with pytest.raises(ValueError, match="S_00009.png.*cannot open"):
        png.read_gray_many(good + [tmp_path / "S_00009.png"])
    # flipped payload byte: CRC error, as libpng / cv2.imread would refuse it
    bad = tmp_path / "S_00005.png"
    data = bytearray(good[0].read_bytes())
    data[data.index(b"IDAT") + 9] ^= 0x55
    bad.write_bytes(bytes(data))
    with pytest.raises(ValueError, match="S_00005.png.*CRC"):
        png.read_gray_many(good[:2] + [bad] + good[2:])
    # colour image
    col = tmp_path / "S_00006.png"
    write_png(col, rng.integers(0, 256, (5, 7, 3), dtype=np.uint8), [0], ctype=2)
    with pytest.raises(ValueError, match="S_00006.png.*colour"):
        png.read_gray_many([col])
    # not a PNG, 16-bit PNG
    (tmp_path / "S_00007.png").write_bytes(b"GIF89a" + bytes(40))
    with pytest.raises(ValueError, match="not a PNG"):
        png.read_gray_many([tmp_path / "S_00007.png"])
    hdr16 = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", 4, 4, 16, 0, 0, 0, 0)) + _chunk(b"IEND", b"")
    (tmp_path / "S_00008.png").write_bytes(hdr16)
    with pytest.raises(ValueError, match="unsupported PNG"):
        png.read_gray_many([tmp_path / "S_00008.png"])
    # truncated image data
    w, h = 7, 5
    stream = b"".join(b"\x00" + bytes(range(w)) for _ in range(h))[:-3]
    short = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(stream)) + _chunk(b"IEND", b"")
    (tmp_path / "S_00010.png").write_bytes(short)
    with pytest.raises(ValueError, match="bytes of image data"):
        png.read_gray_many([tmp_path / "S_00010.png"])


def test_main_image_mode_with_stand_in_engine(tmp_path, monkeypatch):
    """`probability.main(..., samples_as_images=True)`: one CSV per sample directly under OUT (probability.py:96), the next
    sample decoded while the current one runs, existing CSVs skipped (no decode either), engine closed at the end."""
    from sykepic_b200.compute import probability

    class Spec:
        classes = ["a", "b"]
        img_shape = (3, 64, 64)

    class Eng:
        instances = []

        def __init__(self, spec, device=None, precision=None, max_batch=None):
            self.device, self.calls, self.closed = "cpu", [], False
            Eng.instances.append(self)

        def run_rois(self, ids, w, h, start, data, batch_size=None):
            self.calls.append(ids.tolist())
            first = np.array([data[int(s)] for s in start], np.float32) / 255.0
            return np.stack([first, 1.0 - first], axis=1).astype(np.float32)

        def close(self):
            self.closed = True

    monkeypatch.setattr(probability._engine, "Engine", Eng)
    monkeypatch.setattr(probability._engine.ModelSpec, "from_dir", classmethod(lambda cls, d: Spec()))
    monkeypatch.setattr(probability, "_devices", lambda d: [0])
    decoded = []
    real_decode = probability.decode_images
    monkeypatch.setattr(probability, "decode_images", lambda paths: decoded.append(Path(list(paths)[0]).name.rpartition("_")[0]) or real_decode(paths))

    rng = np.random.default_rng(21)
    samples = {}
    for s in range(3):
        name = f"D2021052{s}T000000_IFCB114"
        paths = []
        for i in (3, 1, 2):
            p = tmp_path / f"{name}_{i:05d}.png"
            write_png_up(p, rng.integers(0, 256, (6 + i, 10 + s), dtype=np.uint8))
            paths.append(p)
        samples[name] = sorted(paths)
    out = tmp_path / "out"
    out.mkdir()
    names = list(samples)
    (out / f"{names[1]}.prob.csv").write_text("sentinel")
    assert probability.main(samples, tmp_path / "model", out, progress_bar=False, samples_as_images=True) is None
    eng = Eng.instances[-1]
    assert eng.closed and eng.calls == [[1, 2, 3], [1, 2, 3]]
    assert decoded == [names[0], names[2]]  # the skipped sample is not even decoded
    assert (out / f"{names[1]}.prob.csv").read_text() == "sentinel"
    for nm in (names[0], names[2]):
        lines = (out / f"{nm}.prob.csv").read_text().splitlines()
        assert lines[0] == "roi,a,b" and [l.split(",")[0] for l in lines[1:]] == ["1", "2", "3"]
        first = png.read_gray(samples[nm][0])[0, 0] / 255.0
        assert abs(float(lines[1].split(",")[1]) - first) < 1e-5
    # --force: all three
    decoded.clear()
    probability.main(samples, tmp_path / "model", out, force=True, progress_bar=False, samples_as_images=True)
    assert decoded == names and (out / f"{names[1]}.prob.csv").read_text().startswith("roi,a,b")


def test_forged_header_is_refused_not_allocated(tmp_path):
    for w, h in [(60000, 60000), (2 ** 31, 1), (0, 5)]:
        p = tmp_path / "S_00001.png"
        p.write_bytes(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(b"\x00")) + _chunk(b"IEND", b""))
        with pytest.raises(ValueError, match="bad image size"):
            png.read_gray_many([p])
