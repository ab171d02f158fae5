"""`sykepic class`: per-bin class counts from probabilities and thresholds.

Drop-in for the probabilities-only branch of the reference's
`sykepic.compute.classification` (classification.py:21-48 `main`, :109-135
`class_df_probs_only`, :138-155 `swell_df`, :158-161 `df_to_csv`).  The feature-joined
branch (`--feat`: `class_df`, biomass / coiled-colony corrections, size divisions,
classification.py:51-106,164-284) consumes feature CSVs that are produced outside the
hot path; it is not part of this build and raises.
"""

from pathlib import Path

import pandas as pd

from ..utils import logger
from ..utils.ifcb import filter_out_quality_flagged_samples, sample_to_datetime
from .prediction import prediction_dataframe, threshold_dictionary

log = logger.get_logger("class")


def main(args):
    all_probs = sorted(Path(args.probabilities).glob("**/*.csv"))
    if getattr(args, "exclusion_list", None):
        probs = filter_out_quality_flagged_samples(all_probs, Path(args.exclusion_list))
    else:
        probs = all_probs
    out_file = Path(args.out)
    if out_file.suffix != ".csv":
        raise ValueError("Make sure output file ends with .csv")
    if out_file.is_file():
        if not (args.append or args.force):
            raise FileExistsError(f"{args.out} exists, --append or --force not used")
    if getattr(args, "feat", None):
        raise NotImplementedError(
            "sykepic class --feat joins feature CSVs (sykepic feat output), which is outside the B200 prob/class "
            "hot path; run without --feat for per-class counts")
    df = class_df_probs_only(probs, args.thresholds, progress_bar=True)
    df = swell_df(df)
    df_to_csv(df, out_file, args.append)


def class_df_probs_only(probs, thresholds_file, progress_bar=False):
    """One row per bin: number of classified ROIs per predicted class (threshold-file order) + Total."""
    thresholds = threshold_dictionary(thresholds_file)
    classes = list(thresholds.keys()) + ["Total"]
    rows = []
    iterator = probs
    if progress_bar:
        try:
            from tqdm import tqdm

            iterator = tqdm(probs, desc=f"Processing {len(probs)} samples")
        except ImportError:
            pass
    for prob in iterator:
        prob = Path(prob)
        sample = prob.with_suffix("").stem
        try:
            pdf = prediction_dataframe(prob, thresholds)
            counts = pdf.groupby("prediction", observed=False)["classified"].sum()
        except KeyError:
            continue  # e.g. an empty bin (no 'prediction' column): silently skipped like the reference
        counts.index.name = "class"
        counts.loc["Total"] = len(pdf)
        counts.name = sample
        rows.append(counts)
    df = pd.DataFrame(rows, columns=classes)
    df.index.name = "sample"
    df = df.fillna(0)
    return df.astype(int)


def swell_df(df):
    """Index -> ISO-8601 UTC timestamps named `Time`; adds `Filamentous cyanobacteria` before `Total`;
    underscores -> spaces in the column names.  KeyError when the cyanobacteria classes are not among
    the columns, like the reference (classification.py:143-149)."""
    df.index = df.index.map(lambda x: sample_to_datetime(x, isoformat=True))
    df.index.name = "Time"
    doli_sum = df[["Dolichospermum-Anabaenopsis", "Dolichospermum-Anabaenopsis_coiled"]].sum(axis=1)
    nodu_sum = df[["Nodularia_spumigena", "Nodularia_spumigena-coiled"]].sum(axis=1)
    cyano_sum = df["Aphanizomenon_flosaquae"] + doli_sum + nodu_sum
    df.insert(len(df.columns) - 1, "Filamentous cyanobacteria", cyano_sum)
    df.columns = df.columns.str.replace("_", " ")
    return df


def df_to_csv(df, out_file, append=False):
    append = append and Path(out_file).is_file()
    df.to_csv(out_file, mode="a" if append else "w", header=not append)
