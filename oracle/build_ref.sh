#!/usr/bin/env bash
# Build the UNMODIFIED reference gaf2paf / gaf2unstable / gaffilter binaries into oracle/_ref/
# straight from the sources where they lie under /root/reference (nothing is copied
# into this repository).  Test infrastructure only: the product never calls these.
#
# The reference's own Makefile is not run; the two tools compile from a handful of
# files (Makefile:46-47 gaf2paf, Makefile:103-107 gaf2unstable) with its own flags
# (-O3 -std=c++14, asserts live: no -DNDEBUG).  -fopenmp is dropped: no parallel
# region is ever executed on this path and libgomp.spec is missing
# from the default toolchain here; outputs were verified md5-identical either way.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
CXX="${CXX:-/usr/bin/g++}"
if [ ! -d "$REF" ]; then
    echo "[build_ref] $REF not present (GPU box): keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
FLAGS="-O3 -std=c++14 -pthread -w -I$REF"
$CXX $FLAGS "$REF/gaf2paf_main.cpp" -o "$OUT/gaf2paf" &
$CXX $FLAGS "$REF/gaf2unstable_main.cpp" "$REF/rgfa-split.cpp" -o "$OUT/gaf2unstable" &
$CXX $FLAGS "$REF/gaffilter_main.cpp" -o "$OUT/gaffilter" &   # SURVEY.md §8f N1 (Makefile: gaffilter_main.o alone)
wait
strip "$OUT/gaf2paf" "$OUT/gaf2unstable" "$OUT/gaffilter"
ls -la "$OUT" >&2
