#!/bin/bash
# Installs the UNMODIFIED reference (sykefi/syke-pic, pure Python) into baseline/_ref so that
# `bench.py --impl reference` can time its own `sykepic.compute.probability.main` on the GPU box's host cores.
# Build container only (/root/reference does not exist on the GPU box; baseline/_ref is git-ignored but travels
# with the gpurun snapshot).  /root/reference is read-only and setuptools writes an egg-info into the source tree,
# hence the scratch copy; --no-deps because the pinned torch==1.11 / opencv==4.5.5.64 / pytz / matplotlib are not in
# the offline wheelhouse (the image's torch 2.11, torchvision 0.26 and cv2 4.13 are used; pytz is stubbed by bench.py).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/sykepic" ] || { echo "install_ref.sh: no reference at $REF"; exit 3; }
SCRATCH="$(mktemp -d /tmp/spk_ref_XXXXXX)"
cp -r "$REF" "$SCRATCH/src"
chmod -R u+w "$SCRATCH/src"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$SCRATCH/src"
rm -rf "$SCRATCH"
test -f "$HERE/_ref/sykepic/compute/probability.py" && echo "reference installed in $HERE/_ref"
