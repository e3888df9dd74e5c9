#!/bin/bash
# Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied from the read-only
# reference checkout so that they travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored;
# /root/reference does not exist there).  Nothing under oracle/_ref/ is product code: only tests/, smoke()
# and bench.py's cpu_baseline / --impl reference legs import it, as the checker or the timed CPU baseline.
#   unet_model.py  -- UNet / DoubleConv  (the forward this repository replaces)
#   inference.py   -- load_model / preprocess / run_unet
# usage: oracle/make_ref.sh [reference checkout, default /root/reference]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${1:-/root/reference}"
if [ ! -f "$ref/unet_model.py" ] || [ ! -f "$ref/inference.py" ]; then
  echo "make_ref: no reference checkout at $ref (expected on the GPU box: the prebuilt oracle/_ref/ is used)" >&2
  exit 0
fi
mkdir -p "$here/_ref"
for f in unet_model.py inference.py; do
  cp -f "$ref/$f" "$here/_ref/$f"
done
( cd "$ref" && sha256sum unet_model.py inference.py ) > "$here/_ref/SHA256SUMS"
echo "make_ref: $(wc -l < "$here/_ref/SHA256SUMS") reference modules -> $here/_ref"
