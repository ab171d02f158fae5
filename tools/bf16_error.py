"""Probability error of a precision mode against the committed reference goldens, per test case
(max and mean over all ROIs and classes).  Usage: python tools/bf16_error.py [precision=bf16]
Used for A/B runs of kernel variants (e.g. SPK_STEM=hilo)."""
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sykepic_b200 import engine
from tests.cases import CASES, GOLDEN, case_bins, case_model_dir

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
root = Path(tempfile.mkdtemp())
for case in CASES:
    eng = engine.Engine(case_model_dir(case, root), precision=precision, max_batch=64)
    errs = []
    for bname, b in case_bins(case):
        g = np.load(GOLDEN / f"case_{case}__{bname}.npz")
        _, probs = eng.run_bin(b["adc_text"], b["roi_bytes"])
        errs.append(np.abs(probs - g["probs"]).ravel())
    e = np.concatenate(errs)
    print(f"{case:12s} {precision}: max {e.max():.3e}  mean {e.mean():.3e}  rois*classes {e.size}")
    eng.close()
