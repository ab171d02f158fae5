"""Goldens for `sykepic class --feat`, produced by the REFERENCE itself.

Run in the build container (needs /root/reference):  python tests/golden/make_class_feat_golden.py
Writes tests/golden/class_feat/: a synthetic bin's .prob.csv / .feat.csv, thresholds with the colony classes the
reference hard-codes (Nodularia coiled, Dolichospermum coiled, the cyanobacteria sum of swell_df), a divisions file,
and what the reference's class_df / main return for them.  tests/test_host_classification.py compares ours."""
import sys
import types
from collections import namedtuple
from pathlib import Path

import numpy as np
import pandas  # noqa: F401  (before the pytz stub)

sys.modules.setdefault("pytz", types.ModuleType("pytz")).timezone = lambda name: None
sys.path.insert(0, "/root/reference")
from sykepic.compute import classification as ref  # noqa: E402

OUT = Path(__file__).resolve().parent / "class_feat"
OUT.mkdir(exist_ok=True)
CLASSES = ["Aphanizomenon_flosaquae", "Chaetoceros_sp", "Dolichospermum-Anabaenopsis", "Dolichospermum-Anabaenopsis_coiled",
           "Heterocapsa_triquetra", "Nodularia_spumigena", "Nodularia_spumigena-coiled", "Skeletonema_marinoi"]
THR = [0.30, 0.55, 0.25, 0.35, 0.40, 0.30, 0.20, 0.50]
NAME = "D20210523T000000_IFCB114"
Args = namedtuple("Args", "probabilities feat thresholds divisions out value_column append force exclusion_list")


def main():
    rng = np.random.default_rng(11)
    n = 400
    logits = rng.normal(0, 2.5, (n, len(CLASSES)))
    probs = np.exp(logits)
    probs /= probs.sum(1, keepdims=True)
    (OUT / "prob").mkdir(exist_ok=True)
    (OUT / "feat").mkdir(exist_ok=True)
    with open(OUT / "prob" / f"{NAME}.prob.csv", "w") as fh:
        fh.write("roi," + ",".join(CLASSES) + "\n")
        for i in range(n):
            fh.write(f"{i + 2}," + ",".join(f"{p:.5f}" for p in probs[i]) + "\n")
    bv = rng.lognormal(11, 1.6, n)  # both sides of the 200000 um3 coiled-Nodularia limit
    with open(OUT / "feat" / f"{NAME}.feat.csv", "w") as fh:
        fh.write("# version=py-v4\n# volume_ml=4.25\n")
        fh.write("roi,biovolume_px,biovolume_um3,biomass_ugl,area\n")
        for i in range(n):
            fh.write(f"{i + 2},{bv[i] * 40:.6f},{bv[i]:.6f},{bv[i] / 4.25 / 1000:.9f},{int(bv[i] ** 0.5)}\n")
    thr = OUT / "thresholds.txt"
    thr.write_text("".join(f"{c} {t}\n" for c, t in zip(CLASSES, THR)))
    div = OUT / "divisions.txt"
    div.write_text("Nodularia_spumigena 1000000 9000000\nChaetoceros_sp 2500000\n")
    probs_l, feats_l = sorted((OUT / "prob").glob("*.csv")), sorted((OUT / "feat").glob("*.csv"))
    for d, tag in ((None, "nodiv"), (div, "div")):
        for vc in ("biomass_ugl", "biovolume_um3", "frequency"):
            (OUT / f"ref_class_df_{tag}_{vc}.csv").write_text(ref.class_df(probs_l, feats_l, thr, d, vc).to_csv())
    o = OUT / "ref_main_biomass.csv"
    o.unlink(missing_ok=True)
    ref.main(Args(str(OUT / "prob"), str(OUT / "feat"), str(thr), None, o, "biomass_ugl", False, False, None))
    o = OUT / "ref_main_probs_only.csv"
    o.unlink(missing_ok=True)
    ref.main(Args(str(OUT / "prob"), None, str(thr), None, o, None, False, False, None))
    print("written", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
