"""Work-unit / wave arithmetic of the ResNet convolutions on 148 SMs (no GPU needed): how many M x N tiles a layer has for
a kernel's tile shape, how many waves of 74 CTA pairs (or 148 CTAs) that is, and the fraction of the last wave that is
filled.  This is the quantisation part of the gap to the tensor peak; the rest is per-unit
pipeline efficiency, which the clock traces in the kernels show.

    python tools/wave_model.py [--batch 256] [--arch resnet18]
"""
import argparse
import math

SMS = 148

# (name, cin, cout, in h=w, stride, kernel)
RESNET18 = [
    ("layer1 3x3 64->64", 64, 64, 56, 1, 3), ("layer2.0 3x3/2 64->128", 64, 128, 56, 2, 3), ("layer2 3x3 128->128", 128, 128, 28, 1, 3),
    ("layer3.0 3x3/2 128->256", 128, 256, 28, 2, 3), ("layer3 3x3 256->256", 256, 256, 14, 1, 3),
    ("layer4.0 3x3/2 256->512", 256, 512, 14, 2, 3), ("layer4 3x3 512->512", 512, 512, 7, 1, 3),
]
RESNET50 = [
    ("layer1 1x1 64->256", 64, 256, 56, 1, 1), ("layer1 1x1 256->64", 256, 64, 56, 1, 1), ("layer1 3x3 64->64", 64, 64, 56, 1, 3),
    ("layer2 1x1 128->512", 128, 512, 28, 1, 1), ("layer2 3x3 128->128", 128, 128, 28, 1, 3), ("layer2 1x1 512->128", 512, 128, 28, 1, 1),
    ("layer3 1x1 256->1024", 256, 1024, 14, 1, 1), ("layer3 3x3 256->256", 256, 256, 14, 1, 3), ("layer3 1x1 1024->256", 1024, 256, 14, 1, 1),
    ("layer4 1x1 512->2048", 512, 2048, 7, 1, 1), ("layer4 3x3 512->512", 512, 512, 7, 1, 3), ("layer4 1x1 2048->512", 2048, 512, 7, 1, 1),
]


def halo_pair_units(n, hw, cout, bn):
    """conv3x3_hp_kernel: a CTA takes floor(128 / (w + 2)) output rows of one image; two CTAs form a unit; all of Cout per unit."""
    pitch = hw + 2
    hb = max(1, 128 // pitch)
    tiles = n * math.ceil(hw / hb)
    useful = hb * hw / 128.0  # MMA rows that are real outputs
    return math.ceil(tiles / 2) * math.ceil(cout / bn), SMS // 2, useful


def pair_units(n, ho, cout, bn):
    """conv_pair_kernel: M = 256 output pixels per CTA pair (linear over the batch), N = bn channels."""
    return math.ceil(n * ho * ho / 256) * math.ceil(cout / bn), SMS // 2, 1.0


def report(name, units, slots, useful):
    waves = units / slots
    full = math.ceil(waves)
    return f"{name:28s} units {units:6d}  waves {waves:6.2f} -> {full:3d}  wave efficiency {waves / full:5.1%}  useful MMA rows {useful:5.1%}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--arch", default="resnet18", choices=("resnet18", "resnet50"))
    args = ap.parse_args()
    layers = RESNET18 if args.arch == "resnet18" else RESNET50
    print(f"{args.arch}, batch {args.batch}, {SMS} SMs = {SMS // 2} CTA pairs")
    for name, cin, cout, hw, stride, k in layers:
        ho = hw // stride
        if k == 3 and stride == 1 and cin in (64, 128) and cout <= 128:
            u, slots, useful = halo_pair_units(args.batch, hw, cout, cout)
            print(report(name + " [halo pair]", u, slots, useful))
        else:
            for bn in (128, 256):
                if cout % bn:
                    continue
                u, slots, useful = pair_units(args.batch, ho, cout, bn)
                print(report(f"{name} [pair N={bn}]", u, slots, useful))


if __name__ == "__main__":
    main()
