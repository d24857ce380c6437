#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference bam-readcount binary (and the
# samtools 1.10 CLI it vendors) from the sources where they lie under /root/reference,
# following SURVEY.md Appendix B.  Outputs go only into oracle/_ref/ (git-ignored; it
# travels to the GPU box with gpurun).  Nothing here is linked into the product.
#
# Not the reference's own CMake superbuild: that one also wants curl/mbedtls/bz2/xz.
# We compile the four first-party .cpp files directly against htslib-1.10/libbam from
# the vendored tarball (configured without bz2/lzma/curl, which the fixtures never need).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${BRC_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
W="$OUT/work"
if [ ! -d "$R/src/exe/bam-readcount" ]; then
  echo "build_ref.sh: $R not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
if [ -x "$OUT/bam-readcount" ] && [ -x "$OUT/samtools" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref.sh: oracle/_ref already built"; exit 0
fi
mkdir -p "$W" && cd "$W"
[ -d samtools-1.10 ] || tar xjf "$R/vendor/samtools-1.10.tar.bz2"
[ -d boost-1.55-bamrc ] || tar xzf "$R/vendor/boost-1.55-bamrc.tar.gz"
S="$W/samtools-1.10"; H="$S/htslib-1.10"; B="$W/boost-1.55-bamrc"
( cd "$H" && [ -f libhts.a ] || { ./configure --disable-bz2 --disable-lzma --disable-libcurl --disable-gcs --disable-s3 >/dev/null \
    && make -j8 libhts.a >/dev/null 2>&1; } )
( cd "$S" && [ -x samtools ] || { ./configure --without-curses --disable-bz2 --disable-lzma --disable-libcurl >/dev/null \
    && make -j8 samtools libbam.a >/dev/null 2>&1; } )
mkdir -p "$W/brc/version" && cd "$W/brc"
printf '#pragma once\nconst static char* __g_prog_version = "oracle";\nconst static char* __g_commit_hash = "c7c76e6";\n' > version/version.h
if [ ! -f libboost_po.a ]; then
  g++ -O2 -std=c++0x -w -I"$B" -c "$B"/libs/program_options/src/{cmdline,config_file,convert,options_description,parsers,positional_options,split,utf8_codecvt_facet,value_semantic,variables_map}.cpp
  ar rc libboost_po.a *.o && rm -f *.o
fi
g++ -O2 -std=c++0x -w -Iversion -I"$B" -I"$S" -I"$H" -I"$R/src/lib" \
    "$R/src/exe/bam-readcount/bamreadcount.cpp" "$R"/src/lib/bamrc/{BasicStat,IndelQueue,IndelQueueEntry}.cpp \
    libboost_po.a "$S/libbam.a" "$H/libhts.a" -lz -lpthread -lm -o "$OUT/bam-readcount"
cp "$S/samtools" "$OUT/samtools"
# fixtures the reference's own integration tests use (data, not source)
mkdir -p "$OUT/test-data" && cp "$R"/test-data/* "$OUT/test-data/" && chmod -R u+w "$OUT/test-data"
# the extracted/compiled vendor trees are only needed during the build: drop them so oracle/_ref stays small (it travels to the GPU box)
rm -rf "$W"
echo "build_ref.sh: built $OUT/bam-readcount and $OUT/samtools"
