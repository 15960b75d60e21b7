#!/usr/bin/env bash
# Stage the UNMODIFIED reference (pure Python) and the two MIT-BIH records the parity configs use under oracle/_ref/,
# so that it travels to the GPU box with the gpurun snapshot (oracle/_ref/ is git-ignored, NOT gpurun-ignored; nothing
# from /root/reference ever enters the repository history).  Used there ONLY by the checker side: the reference arm of
# bench.py (--impl reference, cpu_baseline) and tests/test_reference_fit_gpu.py, which drives the reference's own
# include_batch / include_sample with hdpgpc_b200.integration enabled.  The product package never imports it.
#   usage: tools/make_ref.sh [reference root, default /root/reference]
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/../oracle/_ref"
[ -d "$SRC/hdpgpc/hdpgpc" ] || { echo "make_ref: $SRC/hdpgpc/hdpgpc not found (nothing staged)"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/hdpgpc/hdpgpc" "$DST/hdpgpc/data/mitbih" "$DST/hdpgpc/tests"
cp "$SRC"/hdpgpc/hdpgpc/*.py "$DST/hdpgpc/hdpgpc/"
cp "$SRC"/hdpgpc/tests/test_offline.py "$SRC"/hdpgpc/tests/test_online.py "$DST/hdpgpc/tests/"
for rec in 100 102; do
    cp "$SRC/hdpgpc/data/mitbih/$rec.npy" "$SRC/hdpgpc/data/mitbih/${rec}_labels.npy" "$DST/hdpgpc/data/mitbih/"
done
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
chmod -R u+w "$DST"
( cd "$DST" && find . -type f | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "make_ref: staged $(find "$DST" -type f | wc -l) files under oracle/_ref ($(du -sh "$DST" | cut -f1))"
