"""Output paths and bin listing (sykepic/utils/files.py:27-44)."""

from pathlib import Path

from . import ifcb


def sample_csv_path(sample_path, out_dir, suffix=None):
    """OUT/YYYY/MM/DD/<sample><suffix>.csv, the date parsed from the sample name."""
    sample = Path(sample_path).name
    name = sample + (suffix or "") + ".csv"
    return Path(out_dir) / ifcb.sample_to_datetime(sample).strftime("%Y/%m/%d") / name


def list_sample_paths(root_dir, filter=None):
    """Every `**/*.roi` under root_dir without its suffix, in glob order (unsorted, like the reference)."""
    paths = (roi.with_suffix("") for roi in Path(root_dir).glob("**/*.roi"))
    if filter is not None:
        paths = (p for p in paths if p.name in filter)
    return list(paths)


def list_sample_csvs(root_dir, filter=None):
    """Every `**/*.csv` under root_dir; with `filter`, only the files of those samples (`<sample>.<kind>.csv`)."""
    found = Path(root_dir).glob("**/*.csv")
    if not filter:
        return list(found)
    return [csv for csv in found if csv.name.split(".")[0] in filter]
