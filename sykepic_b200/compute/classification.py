"""`sykepic class`: per-bin class counts from probabilities and thresholds.

Drop-in for the reference's `sykepic.compute.classification` (file:line are the reference's):
`main` :21-48, `class_df_probs_only` :109-135, `swell_df` :138-155, `df_to_csv` :158-161 for the
probabilities-only branch, and the feature-joined branch (`--feat`):
`class_df` :51-106, `process_sample` :164-238 (sample volume from the feature file's comment
header, Nodularia / Dolichospermum coiled-colony corrections :12-16,:188-189,:229-237),
`read_divisions` / `divide_row` / `names_of_divisions` :241-292.  The feature CSVs themselves are
produced by `sykepic feat` (the un-vendored `ifcb_features` dependency), outside this build; this
module only joins them with the labels derived from the GPU path's `.prob.csv` files.
"""

from pathlib import Path

import pandas as pd

from ..utils import logger
from ..utils.ifcb import filter_out_quality_flagged_samples, sample_to_datetime
from .prediction import prediction_dataframe, threshold_dictionary

log = logger.get_logger("class")

# colony corrections of the reference (classification.py:12-16)
DOLI_COILED_FACTOR_V2 = 7.056
NODU_COILED_FACTOR = 2.15
NODU_COILED_BIG_BV = 36431
NODU_COILED_BV_THRESHOLD = 200000


def _output_target(args):
    """The summary file and whether it is appended to.  The reference's rules (classification.py:28-33): the name must end
    in `.csv`; an existing file needs --append or --force."""
    target = Path(args.out)
    if target.suffix != ".csv":
        raise ValueError("Make sure output file ends with .csv")
    exists = target.is_file()
    if exists and not (args.append or args.force):
        raise FileExistsError(f"{args.out} exists, --append or --force not used")
    return target, bool(args.append) and exists


def main(args):
    """`sykepic class` (argparse namespace): every `*.csv` under `probabilities`, minus quality-flagged samples, ->
    one summary row per bin, written to `--out`."""
    target, appending = _output_target(args)
    prob_files = sorted(Path(args.probabilities).glob("**/*.csv"))
    flagged = getattr(args, "exclusion_list", None)
    if flagged:
        prob_files = filter_out_quality_flagged_samples(prob_files, Path(flagged))
    feat_root = getattr(args, "feat", None)
    if feat_root:
        table = class_df(prob_files, sorted(Path(feat_root).glob("**/*.csv")), thresholds_file=args.thresholds,
                         divisions_file=getattr(args, "divisions", None),
                         summary_feature=getattr(args, "value_column", None) or "biomass_ugl", progress_bar=True)
    else:
        table = class_df_probs_only(prob_files, args.thresholds, progress_bar=True)
    table = swell_df(table)
    table.to_csv(target, mode="a" if appending else "w", header=not appending)


def class_df_probs_only(probs, thresholds_file, progress_bar=False):
    """One row per bin: number of classified ROIs per predicted class (threshold-file order) + `Total` = all ROIs."""
    thresholds = threshold_dictionary(thresholds_file)
    names = list(thresholds)
    table = {}
    for csv in _wrap(probs, progress_bar, f"Processing {len(probs)} samples"):
        frame = prediction_dataframe(Path(csv), thresholds)
        if "prediction" not in frame.columns:
            continue  # an empty bin has no prediction column: the reference's groupby raises KeyError and the bin is dropped (:122-123)
        kept = frame.loc[frame["classified"], "prediction"].astype(str).value_counts()
        row = {name: int(kept.get(name, 0)) for name in names}
        row["Total"] = len(frame)
        table[_stem(csv)] = row
    out = pd.DataFrame.from_dict(table, orient="index", columns=names + ["Total"], dtype=int)
    out.index.name = "sample"
    return out


# summed into one extra column by swell_df (classification.py:143-149): the first class plus the two genus subtotals, in
# this order -- with the biomass columns of the --feat branch the association of the floating-point sum shows in the
# last printed digit
CYANOBACTERIA = (
    ("Aphanizomenon_flosaquae",),
    ("Dolichospermum-Anabaenopsis", "Dolichospermum-Anabaenopsis_coiled"),
    ("Nodularia_spumigena", "Nodularia_spumigena-coiled"),
)


def swell_df(df):
    """Bin names -> ISO-8601 UTC timestamps (index `Time`); a `Filamentous cyanobacteria` column (the classes above)
    just before `Total`; underscores in the column names become spaces.  KeyError when one of those classes is not a column,
    like the reference."""
    df.index = pd.Index([sample_to_datetime(name, isoformat=True) for name in df.index], name="Time")
    subtotals = [df[list(group)].sum(axis=1) if len(group) > 1 else df[group[0]] for group in CYANOBACTERIA]
    filamentous = subtotals[0]
    for part in subtotals[1:]:
        filamentous = filamentous + part
    df.insert(df.shape[1] - 1, "Filamentous cyanobacteria", filamentous)
    df.columns = [str(c).replace("_", " ") for c in df.columns]
    return df


def df_to_csv(df, out_file, append=False):
    """Writes (or, with `append` and an existing file, appends without a header) the summary table."""
    extend = bool(append) and Path(out_file).is_file()
    df.to_csv(out_file, header=not extend, mode="a" if extend else "w")


# ---------------------------------------------------------------------- feature-joined branch (--feat)
def _stem(path):
    return Path(path).with_suffix("").stem  # <bin>.prob.csv / <bin>.feat.csv -> <bin>


def _pairs(probs, feats):
    """(prob csv, feat csv) per bin.  Equal counts are zipped in sorted order; otherwise every feature file picks
    the probability file of the same bin (classification.py:64-73)."""
    probs, feats = sorted(probs), sorted(feats)
    if len(probs) == len(feats):
        return list(zip(probs, feats))
    by_bin = {}
    for p in probs:
        by_bin.setdefault(_stem(p), []).append(p)
    return [(p, f) for f in feats for p in by_bin.get(_stem(f), [])]


def _wrap(items, progress_bar, desc):
    if progress_bar:
        try:
            from tqdm import tqdm

            return tqdm(items, desc=desc)
        except ImportError:
            pass
    return items


def class_df(probs, feats, thresholds_file, divisions_file=None, summary_feature="biomass_ugl", progress_bar=False):
    """One row per bin: `summary_feature` summed per predicted class (+ size divisions) and `Total`."""
    thresholds = threshold_dictionary(thresholds_file)
    divisions = read_divisions(divisions_file) if divisions_file else None
    rows = []
    for prob_csv, feat_csv in _wrap(_pairs(probs, feats), progress_bar, f"Processing {len(feats)} samples"):
        if _stem(prob_csv) != _stem(feat_csv):
            raise ValueError(f"CSV mismatch: {Path(prob_csv).name} & {Path(feat_csv).name}")
        try:
            per_class = process_sample(prob_csv, feat_csv, thresholds, divisions)
        except KeyError:
            log.exception(_stem(prob_csv))
            continue
        column = per_class[summary_feature]
        column.name = _stem(prob_csv)
        rows.append(column)
    names = set(thresholds)
    if divisions:
        names = (names | set(names_of_divisions(divisions))) - set(divisions)
    df = pd.DataFrame(rows, columns=sorted(names) + ["Total"])
    df.index.name = "sample"
    return df.fillna(0)


def sample_volume(feat_csv):
    """The value of the LAST `# key=value` comment line that precedes the table (`# volume_ml=...`), as text."""
    last = None
    with open(feat_csv) as fh:
        for line in fh:
            if not line.startswith("#"):
                break
            last = line
    return last[1:].strip().split("=")[1]


def process_sample(prob_csv, feat_csv, thresholds, divisions=None, division_column="biovolume_px"):
    """Labels of one bin joined with its per-ROI features -> per-class frequency, biovolume_um3 and biomass_ugl
    (classified ROIs only), sorted by biomass, plus a `Total` row over ALL ROIs."""
    volume_ml = float(sample_volume(feat_csv))
    df = pd.concat([prediction_dataframe(prob_csv, thresholds), pd.read_csv(feat_csv, index_col=0, comment="#")], axis=1)
    df.index.name = "roi"
    # coiled Nodularia colonies: small ones are over-estimated by a constant factor, big ones get a fixed biovolume
    coiled = df["prediction"] == "Nodularia_spumigena-coiled"
    small = coiled & (df["biovolume_um3"] < NODU_COILED_BV_THRESHOLD)
    big = coiled & (df["biovolume_um3"] >= NODU_COILED_BV_THRESHOLD)
    df.loc[small, "biomass_ugl"] /= NODU_COILED_FACTOR
    df.loc[big, "biomass_ugl"] = NODU_COILED_BIG_BV / volume_ml / 1000
    totals = [len(df), df["biovolume_um3"].sum(), df["biomass_ugl"].sum()]  # before the unclassified rows go
    df = df[df["classified"]]
    if df.isna().any(axis=1).any():
        log.warning(f"{Path(feat_csv).name}: classified ROIs without feature values")
    if divisions:
        df = df.apply(divide_row, axis=1, args=(divisions, division_column))
    out = df.groupby("prediction", observed=False).sum()[["classified", "biovolume_um3", "biomass_ugl"]]
    out = out.rename(columns={"classified": "frequency"})
    out.index.name = "class"
    out = out.sort_values("biomass_ugl", ascending=False)
    out = out[out["frequency"] > 0]
    out.loc["Total"] = totals
    if "Dolichospermum-Anabaenopsis_coiled" in out.index:
        out.loc["Dolichospermum-Anabaenopsis_coiled", ["biovolume_um3", "biomass_ugl"]] /= DOLI_COILED_FACTOR_V2
    return out


def read_divisions(division_file):
    """`<class> <int> [<int> ...]` per line -> {class: [size limits]}."""
    divisions = {}
    with open(division_file) as fh:
        for line in fh:
            fields = line.split()
            if fields:
                divisions[fields[0]] = [int(v) for v in fields[1:]]
    return divisions


def divide_row(row, divisions, column):
    """Renames a row's prediction to its size class.  Mirrors the reference's loop literally (classification.py:
    251-272), including that the LAST limit decides: every limit overwrites the name chosen by the previous one."""
    name = row["prediction"]
    limits = divisions.get(name)
    if limits:
        value = row[column]
        new_name = name
        for i, limit in enumerate(limits):
            if value < limit:
                new_name = f"{name}_under_{limit}" if i == 0 else f"{name}_{limits[i - 1]}_{limit}"
            else:
                new_name = f"{name}_over_{limit}"  # (the reference's `i == len(limits)` branch is unreachable)
        row["prediction"] = new_name
    return row


def names_of_divisions(divisions):
    """Column names of the size classes: per class, below its smallest limit, above its largest, then every interval."""
    out = []
    for cls, limits in divisions.items():
        cuts = sorted(limits)
        out += [f"{cls}_under_{cuts[0]}", f"{cls}_over_{cuts[-1]}"]
        out += [f"{cls}_{lo}_{hi}" for lo, hi in zip(cuts[:-1], cuts[1:])]
    return out
