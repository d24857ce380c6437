#!/usr/bin/env bash
# TEST INFRASTRUCTURE — compiles the reference's own main() with the INTEGRATION.md binding applied (oracle/patch_reference.py),
# linked against libbrc_engine.so: the reference's option parsing, htslib file/index/FASTA handling and region loops, with
# fetch_func / pileup_func / the pileup buffer replaced by the C ABI.  Output: oracle/_ref/bam-readcount-brc (git-ignored).
# The patched copy of the source exists only under oracle/_ref/work during the build.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${BRC_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
W="$OUT/work"
PKG="$HERE/../bam_readcount_b200"
if [ ! -d "$R/src/exe/bam-readcount" ]; then
  echo "build_patched_ref.sh: $R not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
if [ -x "$OUT/bam-readcount-brc" ] && [ "$OUT/bam-readcount-brc" -nt "$HERE/patch_reference.py" ] && [ "${1:-}" != "--force" ]; then
  echo "build_patched_ref.sh: oracle/_ref/bam-readcount-brc already built"; exit 0
fi
[ -f "$PKG/libbrc_engine.so" ] || { echo "build_patched_ref.sh: build libbrc_engine.so first" >&2; exit 1; }
mkdir -p "$W" && cd "$W"
[ -d samtools-1.10 ] || tar xjf "$R/vendor/samtools-1.10.tar.bz2"
[ -d boost-1.55-bamrc ] || tar xzf "$R/vendor/boost-1.55-bamrc.tar.gz"
S="$W/samtools-1.10"; H="$S/htslib-1.10"; B="$W/boost-1.55-bamrc"
( cd "$H" && [ -f libhts.a ] || { ./configure --disable-bz2 --disable-lzma --disable-libcurl --disable-gcs --disable-s3 >/dev/null \
    && make -j8 libhts.a >/dev/null 2>&1; } )
( cd "$S" && [ -f libbam.a ] || { ./configure --without-curses --disable-bz2 --disable-lzma --disable-libcurl >/dev/null \
    && make -j8 libbam.a >/dev/null 2>&1; } )
mkdir -p "$W/brc/version" && cd "$W/brc"
printf '#pragma once\nconst static char* __g_prog_version = "oracle+brc";\nconst static char* __g_commit_hash = "c7c76e6";\n' > version/version.h
if [ ! -f libboost_po.a ]; then
  g++ -O2 -std=c++0x -w -I"$B" -c "$B"/libs/program_options/src/{cmdline,config_file,convert,options_description,parsers,positional_options,split,utf8_codecvt_facet,value_semantic,variables_map}.cpp
  ar rc libboost_po.a *.o && rm -f *.o
fi
python "$HERE/patch_reference.py" "$R" "$W/brc/bamreadcount_brc.cpp"
g++ -O2 -std=c++0x -w -Iversion -I"$B" -I"$S" -I"$H" -I"$R/src/lib" -I"$R/src/exe/bam-readcount" -I"$HERE/../include" \
    "$W/brc/bamreadcount_brc.cpp" "$R"/src/lib/bamrc/{BasicStat,IndelQueue,IndelQueueEntry}.cpp \
    libboost_po.a "$S/libbam.a" "$H/libhts.a" -L"$PKG" -lbrc_engine -Wl,-rpath,'$ORIGIN/../../bam_readcount_b200' -lz -lpthread -lm -o "$OUT/bam-readcount-brc"
rm -rf "$W"
echo "build_patched_ref.sh: built $OUT/bam-readcount-brc"
