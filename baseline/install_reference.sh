#!/bin/bash
# Installs the UNMODIFIED reference (tools21cm v2.4.0) into baseline/_ref (git-ignored, shipped to the GPU box by
# gpurun) so that `bench.py --impl reference` and the cpu_baseline leg can time the reference's own CPU code.
# /root/reference is read-only and the build writes into the source tree, hence the copy under /tmp.
# --no-deps: only numpy / scipy are needed by power_spectrum.py; astropy, scikit-image ... are not in this image.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC" ] || { echo "no reference at $SRC"; exit 0; }
TMP="$(mktemp -d /tmp/t2c_ref.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
rm -rf "$ROOT/baseline/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
# bench inputs only need the python sources: drop the bundled data tables (SKA layouts, 20 MB)
rm -rf "$ROOT/baseline/_ref/tools21cm/input_data"
echo "reference installed in baseline/_ref: $(ls "$ROOT/baseline/_ref")"
