#!/usr/bin/env bash
# ORACLE tooling (test / baseline infrastructure only, never the product path).
#
# The reference (xl-tang3/RCOT) is pure Python: "building" it means placing its UNMODIFIED hot-path
# files where the GPU box can import them.  /root/reference does not exist on that box, so this
# recipe copies the four files + util/ the reference's trainer.py imports into oracle/_ref/
# (git-ignored -- reference sources never enter the history -- but shipped by gpurun like a built .so).
# Consumers: oracle/ref_shim.py (bench.py --impl reference, bench.py's gpu_eager_baseline,
# scripts/loss_curve.py, tests that run the verbatim reference).
set -euo pipefail
SRC="${RCOT_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/Net_Restormer.py" ]; then
  echo "build_ref: $SRC not present; keeping existing $DST (if any)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/util"
cp "$SRC/Net_Restormer.py" "$SRC/trainer.py" "$SRC/utils.py" "$DST/"
cp "$SRC"/util/*.py "$DST/util/"
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/COMMIT"
echo "build_ref: copied reference hot-path files to $DST"
