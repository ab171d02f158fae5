"""Oracle: IFCB raw decode (.adc text + .roi byte stream -> per-ROI uint8 images).

Test infrastructure only (see oracle/__init__.py).

Follows sykepic/utils/ifcb.py:76-118 (`raw_to_png`) and :121-145
(`raw_to_numpy` / `next_roi`): the ROI id is the 1-based LINE NUMBER of the
.adc file (not the trigger column), width/height/start byte are comma fields
15/16/17, rows with width < 1 or height < 1 are skipped, and a slice that is
shorter than width*height makes numpy's reshape raise ValueError, which
sykepic/compute/probability.py:111-112 turns into "Faulty raw data" (whole bin
skipped).
"""

from pathlib import Path

import numpy as np


def parse_adc_text(text):
    """.adc text -> list of (roi_id, width, height, start) for non-empty ROIs.

    sykepic/utils/ifcb.py:102-110.  `int()` semantics of the reference are kept:
    surrounding whitespace and a sign are accepted; anything else raises
    ValueError, a short line raises IndexError.
    """
    import io

    rows = []
    # `for line in fh` with universal newlines: \n, \r\n and \r end a line
    # (and only those; str.splitlines would also split on \x0b, \x0c, ...).
    for i, line in enumerate(io.StringIO(text, newline=None), start=1):
        f = line.split(",")
        w = int(f[15])
        h = int(f[16])
        start = int(f[17])
        if w < 1 or h < 1:
            continue
        rows.append((i, w, h, start))
    return rows


def parse_adc(adc_path):
    # newline=None (the default of open()) == universal newlines.
    with open(adc_path) as fh:
        return parse_adc_text(fh.read())


def decode_rois(rows, roi_bytes):
    """Yield (roi_id, (h,w) uint8 array); sykepic/utils/ifcb.py:111-116."""
    roi_bytes = np.asarray(roi_bytes, dtype=np.uint8)
    for roi_id, w, h, start in rows:
        end = start + w * h
        # numpy slicing clamps; reshape raises ValueError when truncated.
        yield roi_id, roi_bytes[start:end].reshape((h, w))


def raw_to_numpy(adc_path, roi_path):
    """sykepic/utils/ifcb.py:121-130."""
    roi_bytes = np.fromfile(roi_path, dtype=np.uint8)
    return decode_rois(parse_adc(adc_path), roi_bytes)


def sample_to_datetime(sample, isoformat=False):
    """sykepic/utils/ifcb.py:16-43: name[1:16] parsed as %Y%m%dT%H%M%S, UTC."""
    import datetime

    ts = datetime.datetime.strptime(sample[1:16], "%Y%m%dT%H%M%S")
    ts = ts.replace(tzinfo=datetime.timezone.utc)
    return ts.isoformat() if isoformat else ts


def sample_csv_path(sample_path, out_dir, suffix=None):
    """sykepic/utils/files.py:27-37."""
    sample = Path(sample_path).name
    name = sample + (suffix or "") + ".csv"
    return Path(out_dir) / sample_to_datetime(sample).strftime("%Y/%m/%d") / name
