#!/usr/bin/env bash
# Builds htslib 1.10 — the copy the reference vendors (R:vendor/samtools-1.10.tar.bz2, R:cmake/BuildSamtools.cmake:3) — as a static
# library for the C++ host's CRAM input: bam_readcount_b200/third_party/htslib/{libhts.a, htslib/*.h} (git-ignored build
# output; it travels to the GPU box with the prebuilt brc-readcount).  Third-party dependency, like zlib; no reference code.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${BRC_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/../bam_readcount_b200/third_party/htslib"
if [ -f "$OUT/libhts.a" ] && [ "${1:-}" != "--force" ]; then echo "build_htslib.sh: $OUT/libhts.a present"; exit 0; fi
if [ ! -f "$R/vendor/samtools-1.10.tar.bz2" ]; then echo "build_htslib.sh: $R/vendor/samtools-1.10.tar.bz2 not present; CRAM input stays disabled" >&2; exit 0; fi
W="$(mktemp -d)"
trap 'rm -rf "$W"' EXIT
cd "$W" && tar xjf "$R/vendor/samtools-1.10.tar.bz2" samtools-1.10/htslib-1.10
cd samtools-1.10/htslib-1.10
./configure --disable-bz2 --disable-lzma --disable-libcurl --disable-gcs --disable-s3 >/dev/null 2>&1
make -j8 libhts.a >/dev/null 2>&1
mkdir -p "$OUT/htslib"
cp libhts.a "$OUT/" && cp htslib/*.h "$OUT/htslib/"
echo "build_htslib.sh: built $OUT/libhts.a"
