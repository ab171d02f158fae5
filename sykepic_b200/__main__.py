"""`sykepic` command line: the `prob` and `class` sub-commands of the reference
(sykepic/__main__.py:63-99 and :134-190) with identical flags, served by the B200 path.

Extra flags of `prob` (not in the reference): `--precision {fp32,fp32_tc,bf16}` and `--gpus N`.
The reference's other sub-commands (train, feat, size, abundance, class_stats,
features_per_prediction) are outside this build.
"""

import argparse
import sys

from .compute import classification, probability


def build_parser():
    parser = argparse.ArgumentParser(prog="sykepic", description="syke-pic prob/class hot path on B200 GPUs")
    subparsers = parser.add_subparsers(dest="command")

    prob_parser = subparsers.add_parser("prob", description="Calculate class probabilities")
    prob_parser.set_defaults(func=probability.call)
    prob_raw = prob_parser.add_mutually_exclusive_group(required=True)
    prob_raw.add_argument("-r", "--raw", metavar="DIR", help="Root directory of raw IFCB data")
    prob_raw.add_argument("-s", "--samples", nargs="+", metavar="SAMPLE PATH",
                          help="One or more sample paths (raw file without suffix)")
    prob_raw.add_argument("--image-dir", metavar="DIR", help="Root directory of images")
    prob_raw.add_argument("--images", nargs="+", metavar="FILE", help="One or more image paths")
    prob_parser.add_argument("-m", "--model", required=True, help="Model directory")
    prob_parser.add_argument("-o", "--out", required=True, help="Root output directory")
    prob_parser.add_argument("-b", "--batch-size", type=int, default=64, metavar="INT", help="Default is 64")
    prob_parser.add_argument("-w", "--num-workers", type=int, default=2, metavar="INT",
                             help="Accepted for compatibility (no loader processes on this path)")
    prob_parser.add_argument("-f", "--force", action="store_true", help="Force overwrite of previous probabilities")
    prob_parser.add_argument("--precision", choices=("fp32", "fp32_tc", "bf16"), default=None,
                             help="fp32_tc (default: fp32-level accuracy on the tensor cores, within 1e-4 of the reference), "
                                  "fp32 (exact CUDA-core path) or bf16 tensor cores (within 2e-2)")
    prob_parser.add_argument("--gpus", dest="devices", type=int, default=None, metavar="N",
                             help="Shard bins over the first N GPUs of this box (default 1)")

    class_parser = subparsers.add_parser(
        "class", description="Use thresholds together with probabilities for classification")
    class_parser.set_defaults(func=classification.main)
    class_parser.add_argument("probabilities", help="Root directory of probabilities")
    class_parser.add_argument("--feat", metavar="DIR", help="Root directory of features (and use them in results)")
    class_parser.add_argument("-t", "--thresholds", metavar="FILE", required=True,
                              help="Probability thresholds file (required)")
    class_parser.add_argument("-d", "--divisions", metavar="FILE", help="Feature divisions file (optional)")
    class_parser.add_argument("-o", "--out", metavar="FILE", required=True, help="Output CSV-file path (required)")
    class_parser.add_argument("-v", "--value-column", metavar="FEATURE", default="biomass_ugl",
                              help="Feature used to aggregate results, default is biomass_ugl")
    class_parser.add_argument("-a", "--append", action="store_true", help="Append to output file if it exists")
    class_parser.add_argument("-f", "--force", action="store_true", help="Overwrite output file if it exists")
    class_parser.add_argument("-exc", "--exclusion_list", metavar="FILE",
                              help="Text file containing a list of sample names to exclude e.g. D20180703T181501")
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if not getattr(args, "func", None):
        parser.print_help()
        return 2
    args.func(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
