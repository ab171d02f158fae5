#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_resnet18.txt
    python profiles/summarize_ncu.py full gpurun_out/prof_conv_tc.ncu-rep profiles/r1_ncu_conv_tc.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst):
    rows = []
    with open(src) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        short = name.split("(")[0].replace("spk::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        out.write(f"# {len(rows)} launches, {total / 1e3:.1f} us total\n")
        out.write("kernel\tlaunches\ttotal_us\tshare\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"{k}\t{n}\t{ns / 1e3:.1f}\t{ns / total:.3f}\n")
        out.write("\n# every launch in order\nidx\tkernel\tgrid\tblock\tus\n")
        for i, (name, ns, grid, block) in enumerate(rows):
            short = name.split("(")[0].replace("spk::<unnamed>::", "").replace("void ", "")
            out.write(f"{i}\t{short}\t{grid}\t{block}\t{ns / 1e3:.2f}\n")


def full(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units, body = r[0], r[1], r[2:]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full --clock-control none --import-source on; source: {src}\n")
        for row in body:
            d = dict(zip(hdr, row))
            out.write(f"\n== {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for k in RAW_KEYS:
                if k in d:
                    out.write(f"{k}\t{d[k]}\t{units[hdr.index(k)]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
