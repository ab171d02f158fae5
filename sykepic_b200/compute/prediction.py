"""Predictions from class probabilities and thresholds.

Drop-in for the reference's `sykepic.compute.prediction` (prediction.py:8-79): same
functions, same DataFrame layout (`prediction` category column at 0, `classified` at 1).
The per-row rule (`row_prediction` :49-71: highest-probability class that is at or above
its own threshold, else `(idxmax, False)`; scalar threshold: `(idxmax, p > thr)`) is
evaluated for all rows at once with numpy instead of `df.apply`; the same rule runs fused
in the CUDA head kernel (csrc/head.cu) when labels are requested from the engine.
Ties between equal decimals go to the lowest column index (the reference's `idxmax`;
its descending `sort_values` is formally unstable).
"""

from pathlib import Path

import numpy as np
import pandas as pd


def read_prob_csv(csv, index_col=None):
    """`pd.read_csv(csv)` / `pd.read_csv(csv, index_col=0)` for a `.prob.csv`, through the C-ABI host parser
    (`spk_prob_csv_parse`: same float64 values as pandas' default parser -- correctly rounded -- at a tenth of the time;
    pandas.read_csv is 3/4 of `sykepic class`).  Anything that is not the plain `roi,<class>,...` layout, an empty file, or a
    missing library goes to pandas itself."""
    try:
        import ctypes as C

        from .. import _lib

        lib = _lib.load()
        data = Path(csv).read_bytes()
        n_rows, n_cols = C.c_int64(), C.c_int()
        if lib.spk_prob_csv_shape(data, len(data), C.byref(n_rows), C.byref(n_cols)) != 0 or n_rows.value == 0 or n_cols.value < 1:
            raise ValueError
        header = data[: data.index(b"\n") if b"\n" in data else len(data)].rstrip(b"\r").decode("utf-8")
        names = header.split(",")
        if len(names) != n_cols.value + 1 or len(set(names)) != len(names) or any(n != n.strip() or not n for n in names):
            raise ValueError  # duplicate / padded / empty names: pandas renames or strips them its own way
        roi = np.empty(n_rows.value, np.int64)
        values = np.empty((n_rows.value, n_cols.value), np.float64)
        if lib.spk_prob_csv_parse(data, len(data), n_rows.value, n_cols.value, roi.ctypes.data, values.ctypes.data) != 0:
            raise ValueError
    except Exception:
        return pd.read_csv(csv, index_col=index_col)
    if index_col == 0:
        return pd.DataFrame(values, index=pd.Index(roi, name=names[0]), columns=names[1:])
    df = pd.DataFrame(values, columns=names[1:])
    df.insert(0, names[0], roi)
    return df


def _frame_of_many(csv_files):
    """Several bins in one frame, indexed by (sample, roi); the sample is the file name without `.prob.csv`."""
    parts = []
    for csv in csv_files:
        part = read_prob_csv(csv)
        part.insert(0, "sample", Path(csv).with_suffix("").stem)
        parts.append(part.set_index(["sample", "roi"]))
    return pd.concat(parts)


def prediction_dataframe(probabilities, thresholds=0.0):
    """One `.prob.csv` (path) or several (list of paths) -> probabilities with `prediction` / `classified` in front.
    `thresholds`: a number, a {class: value} dict, or the path of a thresholds file."""
    if isinstance(probabilities, (str, Path)):
        frame = read_prob_csv(probabilities, index_col=0)
    elif isinstance(probabilities, list):
        frame = _frame_of_many(probabilities)
    else:
        raise ValueError(f"Type {type(probabilities)} not allowed for probabilities")
    if isinstance(thresholds, (str, Path)):
        thresholds = threshold_dictionary(thresholds)
    if len(frame.index) and len(frame.columns):
        insert_prediction(frame, thresholds)
    return frame


def threshold_dictionary(thresholds, default=None):
    """Thresholds file -> {class: value}.  One `<class> <value>` per line, any whitespace between them; a line with the
    class alone takes `default` (ValueError without one); a blank line is an IndexError, as in the reference."""
    with open(thresholds) as fh:
        entries = [line.split() for line in fh]
    table = {}
    for fields in entries:
        name = fields[0]
        if len(fields) >= 2:
            table[name] = float(fields[1])
        elif default:
            table[name] = float(default)
        else:
            raise ValueError(f"Missing threshold for {name}, and no default value specified.")
    return table


def predict_array(values, classes, thresholds):
    """[N,K] probabilities (as read from the CSV) -> (class index int64[N], classified bool[N])."""
    values = np.asarray(values, dtype=np.float64)
    if values.ndim != 2:
        values = values.reshape(len(values), -1)
    best = values.argmax(axis=1)  # first maximum == Series.idxmax
    rows = np.arange(len(values))
    if isinstance(thresholds, (int, float)):
        return best, values[rows, best] > thresholds
    thr = np.array([thresholds.get(name, np.inf) for name in classes], dtype=np.float64)
    has = np.array([name in thresholds for name in classes], dtype=bool)
    ok = (values >= thr[None, :]) & has[None, :]
    masked = np.where(ok, values, -np.inf)
    cand = masked.argmax(axis=1)
    classified = ok.any(axis=1)
    return np.where(classified, cand, best), classified


def row_prediction(row, thresholds):
    """(name, classified) for one row (a Series of probabilities indexed by class name)."""
    idx, flag = predict_array(np.asarray(row.values, dtype=np.float64)[None, :], list(row.index), thresholds)
    return (row.index[int(idx[0])], bool(flag[0]))


def insert_prediction(df, thresholds):
    """Modifies `df` in place: inserts `prediction` (category) and `classified` (bool)."""
    classes = list(df.columns)
    idx, flags = predict_array(df.to_numpy(dtype=np.float64), classes, thresholds)
    df.insert(0, "prediction", [classes[i] for i in idx])
    df["prediction"] = df["prediction"].astype("category")
    df.insert(1, "classified", [bool(f) for f in flags])
