#!/usr/bin/env python
"""Turn the round's ncu captures (tools/gpu/ncu_r2.sh -> gpurun_out/r2_*) into the committed evidence under profiles/:

    python tools/ncu_to_profiles.py            # launches list, one summary per capture, profiles/r2_traffic.json

`r2_traffic.json` is what bench.py reads for `roofline.traffic` / `roofline_preprocess.traffic`: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of the dominant convolution kernel and of K1, next to their algorithmic bytes.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "profiles"))
import summarize_ncu  # noqa: E402

OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units = r[0], r[1]
    return [({k: v for k, v in zip(hdr, row)}, dict(zip(hdr, units))) for row in r[2:]]


def dram(d, u):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k].replace(",", "")) * BYTES[u[k]]
    return tot


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    if (OUT / f"{tag}_launches.csv").exists():
        summarize_ncu.launches(str(OUT / f"{tag}_launches.csv"), str(PROF / f"{tag}_launches_resnet18_bf16.txt"))
    traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` at HEAD "
                           f"(tools/gpu/ncu_r2.sh; summaries in profiles/{tag}_ncu_*.txt), ResNet-18 224 bf16 batch 256, K1 chunk 4096"}
    for rep in sorted(OUT.glob(f"{tag}_*.ncu-rep")):
        name = rep.stem[len(tag) + 1:]
        dst = PROF / f"{tag}_ncu_{name}.txt"
        summarize_ncu.full(str(rep), str(dst))
        rows = raw_rows(rep)
        if name == "conv3x3_hp64":
            # launch 0 = layer1.0 conv1 (no residual), launch 1 = conv2 (+residual); 103 MB per 256 x 56 x 56 x 64 bf16 tensor
            t = 256 * 56 * 56 * 64 * 2
            per = [dram(d, u) for d, u in rows[:2]]
            traffic["conv3x3_hp_kernel<64>"] = {
                "dram_bytes_per_launch": sum(per) / len(per), "algorithmic_bytes_per_launch": (2 * t + 3 * t) / 2,
                "detail": f"no residual {per[0] / 1e6:.1f} MB (algorithmic {2 * t / 1e6:.1f}: input + output); "
                          f"+residual {per[1] / 1e6:.1f} MB (algorithmic {3 * t / 1e6:.1f}: input + residual + output)",
                "source": f"profiles/{dst.name}"}
        elif name == "preprocess_u8":
            d, u = rows[0]
            traffic["preprocess_u8_kernel"] = {"dram_bytes_per_launch": dram(d, u), "source": f"profiles/{dst.name}",
                                               "grid": d.get("Grid Size"), "duration_us": d.get("gpu__time_duration.sum")}
        elif name in ("stem_pool_t", "head", "conv_pair"):
            per = [dram(d, u) for d, u in rows]
            traffic[f"{name}_kernel"] = {"dram_bytes_per_launch": per, "source": f"profiles/{dst.name}"}
    (PROF / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
