"""Oracle: .prob.csv text, thresholds and the per-ROI label rule.

Test infrastructure only (see oracle/__init__.py).

* `probabilities_to_csv_text`: sykepic/compute/probability.py:200-206.
* `threshold_dictionary`:      sykepic/compute/prediction.py:31-46.
* `row_prediction` / `predict`: sykepic/compute/prediction.py:49-79.  The rule
  works on the 5-decimal values re-read from the CSV.  Ties between equal
  decimals: `idxmax` takes the first column; the descending `sort_values` of
  the reference is numpy's non-stable sort, so the order of exactly-equal
  above-threshold candidates is formally undefined there -- the restatement
  (and the CUDA kernel) pick the lowest class index.
* `class_counts_probs_only`:   sykepic/compute/classification.py:109-135.
"""

import numpy as np


def probabilities_to_csv_text(rows, classes):
    """rows: iterable of (roi_id, [p...]) with p Python floats (fp32 widened)."""
    out = ["roi," + ",".join(classes) + "\n"]
    for roi, probs in rows:
        out.append(f"{roi}," + ",".join(f"{p:.5f}" for p in probs) + "\n")
    return "".join(out)


def parse_prob_csv_text(text):
    """-> (classes, roi_ids int64[N], values float64[N,K]) like pd.read_csv(index_col=0)."""
    lines = text.splitlines()
    classes = lines[0].split(",")[1:]
    ids, vals = [], []
    for line in lines[1:]:
        if not line:
            continue
        f = line.split(",")
        ids.append(int(f[0]))
        vals.append([float(v) for v in f[1:]])
    return classes, np.asarray(ids, dtype=np.int64), np.asarray(vals, dtype=np.float64).reshape(len(ids), len(classes))


def threshold_dictionary(path, default=None):
    thres = {}
    with open(path) as fh:
        for line in fh:
            f = line.strip().split()
            key = f[0]
            if len(f) > 1:
                value = float(f[1])
            elif default:
                value = float(default)
            else:
                raise ValueError(f"Missing threshold for {key}, and no default value specified.")
            thres[key] = value
    return thres


def row_prediction(values, classes, thresholds):
    """One row of decimals -> (class name, classified)."""
    values = list(values)
    best = max(range(len(values)), key=lambda k: (values[k], -k))  # idxmax: first maximum
    if isinstance(thresholds, (int, float)):
        return classes[best], bool(values[best] > thresholds)
    order = sorted(range(len(values)), key=lambda k: (-values[k], k))
    for k in order:
        name = classes[k]
        if name in thresholds and values[k] >= thresholds[name]:
            return name, True
    return classes[best], False


def predict(values, classes, thresholds):
    """[N,K] decimals -> (list of names, bool array)."""
    names, flags = [], []
    for row in np.asarray(values):
        n, f = row_prediction(row, classes, thresholds)
        names.append(n)
        flags.append(f)
    return names, np.asarray(flags, dtype=bool)


def class_counts_probs_only(values, classes, thresholds):
    """One bin -> {class: count of classified ROIs predicted as class, ..., 'Total': N}.

    Columns are the threshold-file order + 'Total' (classification.py:111,132);
    predicted classes that have no threshold entry are dropped by the
    DataFrame(columns=...) reindex.
    """
    names, flags = predict(values, classes, thresholds)
    counts = {name: 0 for name in thresholds}
    for n, f in zip(names, flags):
        if f and n in counts:
            counts[n] += 1
    counts["Total"] = len(names)
    return counts
