"""Builds libsykepic_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m sykepic_b200._build [--force] [--verbose] [--debug]

`--debug` adds -DSPK_DEBUG_SWITCHES: the A/B and trace switches of tools/README.md are then read from the environment
(csrc/spk_debug.h); the product build has them compiled out.

Every translation unit is compiled with
`-gencode arch=compute_100a,code=sm_100a -lineinfo` and linked (static cudart)
into `sykepic_b200/libsykepic_b200.so`.  The library is built here in the
build container (nvcc cross-compiles without a GPU) and travels to the GPU box
with the repo snapshot; it is git-ignored.
"""

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"
BUILD = ROOT / "build" / "spk"
LIB = PKG / "libsykepic_b200.so"

SOURCES = ["host.cpp", "png.cpp", "net.cu", "preprocess.cu", "conv_simt.cu", "conv_tc.cu", "conv_pair.cu", "conv_hp.cu", "stem.cu", "stem_t.cu", "head.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", f"-I{INCLUDE}", f"-I{CSRC}"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(str(p).encode())
        h.update(Path(p).read_bytes())
    return h.hexdigest()


def _compile(src, obj, verbose, extra=()):
    cmd = [nvcc(), *ARCH, *FLAGS, *extra, "-x", "cu", "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 or verbose:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}")


def build(force=False, verbose=False, debug=False):
    """Compile (if stale) and return the path of the shared library."""
    extra = ("-DSPK_DEBUG_SWITCHES",) if debug else ()
    BUILD.mkdir(parents=True, exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    flags_key = " ".join(ARCH + FLAGS + list(extra))
    hdr_digest = _digest(headers, flags_key)
    jobs, objs = [], []
    for src in srcs:
        obj = BUILD / (src.stem + ".o")
        stamp = BUILD / (src.stem + ".sha")
        want = _digest([src], hdr_digest)
        objs.append(obj)
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != want:
            jobs.append((src, obj, stamp, want))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda j: _compile(j[0], j[1], verbose, extra), jobs))
        for _, _, stamp, want in jobs:
            stamp.write_text(want)
    if jobs or not LIB.exists():
        cmd = [nvcc(), *ARCH, "-shared", "-cudart", "static", "-o", str(LIB), *map(str, objs), "-lz"]  # zlib: png.cpp
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libsykepic_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug="--debug" in sys.argv))
