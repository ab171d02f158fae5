"""IFCB helpers of the hot path (host side).

Mirrors the parts of sykepic/utils/ifcb.py the prob / class path uses:
`sample_to_datetime` :16-43, `raw_to_numpy` / `next_roi` :121-145 (here backed by the
C-ABI .adc parser; no PNG files are ever written) and
`filter_out_quality_flagged_samples` :149-157.
"""

import datetime
from pathlib import Path

import numpy as np

from .. import engine


def sample_to_datetime(sample, isoformat=False):
    """D20180703T093453_IFCB114 -> aware UTC datetime (or its ISO-8601 string)."""
    ts = datetime.datetime.strptime(sample[1:16], "%Y%m%dT%H%M%S").replace(tzinfo=datetime.timezone.utc)
    return ts.isoformat() if isoformat else ts


def read_raw(adc, roi):
    """-> (roi_id, width, height, start, roi_bytes uint8[...]) of one bin; geometry is NOT validated."""
    with open(adc, "rb") as fh:
        roi_id, w, h, start = engine.parse_adc(fh.read())
    return roi_id, w, h, start, np.fromfile(roi, dtype=np.uint8)


def raw_to_numpy(adc, roi):
    """Generator of (roi_id, (h,w) uint8 array), the reference's `raw_to_numpy` (ifcb.py:121-130).

    A slice shorter than width*height raises ValueError like numpy's reshape does there."""
    roi_id, w, h, start, data = read_raw(adc, roi)
    for i in range(len(roi_id)):
        s, e = int(start[i]), int(start[i]) + int(w[i]) * int(h[i])
        yield int(roi_id[i]), data[s:e].reshape((int(h[i]), int(w[i])))


def filter_out_quality_flagged_samples(sample_paths, exclusion_list):
    """Drop every path that contains one of the listed sample names (substring match, ifcb.py:149-157)."""
    with open(exclusion_list) as fh:
        exclude = [line.strip() for line in fh]
    return [Path(p) for p in sample_paths if not any(s in str(p) for s in exclude)]
