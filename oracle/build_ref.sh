#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Builds the reference's own ContigsMerger (the CPU path this repo accelerates) from the
# sources where they lie under /root/reference, into oracle/_ref/ (git-ignored):
#   oracle/_ref/ContigsMerger   the reference binary, reference flags (-O3 -mcmodel=medium)
#   oracle/_ref/libcm_ref.so    the same objects + oracle/ref_harness.cpp (C-ABI around Evaluate)
#   oracle/_ref/libla_ref.so    TERefiner's affine local aligner (TERefiner/algorithms/local_alignment.cpp, unpatched,
#                               compiled where it lies) + oracle/la_harness.cpp (C-ABI around optAlign / aln_stdaln)
#
# The sources are copied to a scratch directory under $TMPDIR, patched there, compiled and the
# scratch directory is removed: no reference source is ever written into this repository.
# Patches: two non-void functions fall off their end, which g++ >= 8
# turns into a crash; nothing else about the algorithm is touched.
#   1. ContigsCompactor.h:57    SetContainedFlag        -> add `return b;`
#   2. ContigsCompactor.cpp:769 ContigsCompactor::addEdges -> add `return 0;`
# For libcm_ref.so only, one observation hook is inserted before ContigsCompactor.cpp:1712
# (`int res = 2;`) so the best-cell scan result can be read for rejected pairs as well.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GAPPADDER_REFERENCE:-/root/reference}"
SRC="$REF/ContigsCompactor-v0.2.0/ContigsMerger"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
    echo "build_ref.sh: reference sources not found at $SRC (expected on the GPU box); keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d "${TMPDIR:-/tmp}/cmref.XXXXXX")"
trap 'rm -rf "$TMP"' EXIT
UNITS="abstractcharsequence fastareader fastaMultiSeqs GenSeqsUtils MurmurHash3 ContigsCompactor KmerUtils Utils-basic GraphUtils ScaffoldUtils"
for u in $UNITS main; do cp "$SRC/$u.cpp" "$TMP/"; done
cp "$SRC"/*.h "$TMP/"
rm -f "$TMP"/._*
# patch 1 + 2 (byte-wise; the sources are Latin-1)
LC_ALL=C sed -i 's/bool SetContainedFlag(bool b) { bcontained = b; }/bool SetContainedFlag(bool b) { bcontained = b; return b; }/' "$TMP/ContigsCompactor.h"
grep -q 'bcontained = b; return b;' "$TMP/ContigsCompactor.h"
LC_ALL=C awk 'BEGIN{inadd=0} /^int ContigsCompactor::addEdges\(\)/{inadd=1} { if (inadd && $0 ~ /^}/) { print "\treturn 0;"; inadd=0 } print }' \
    "$TMP/ContigsCompactor.cpp" > "$TMP/cc.tmp" && mv "$TMP/cc.tmp" "$TMP/ContigsCompactor.cpp"
[ "$(grep -c $'^\treturn 0;$' "$TMP/ContigsCompactor.cpp")" -ge 1 ]
CXX="${CXX:-g++}"
CXXFLAGS="-O3 -w -mcmodel=medium"
cd "$TMP"
# (a) the reference binary
OBJS=""
for u in $UNITS main; do $CXX $CXXFLAGS -c "$u.cpp" -o "$u.o" & OBJS="$OBJS $u.o"; done; wait
$CXX $CXXFLAGS -o "$OUT/ContigsMerger" $OBJS -lz -lm -lpthread
# (b) harness .so: same units (PIC), Evaluate hook, no main.cpp
LC_ALL=C sed -i 's/^\tint res = 2;$/\tcmref_hook(scoreMax, posRowEnd, posColEnd, nclip);\n\tint res = 2;/' ContigsCompactor.cpp
grep -q 'cmref_hook(scoreMax' ContigsCompactor.cpp
LC_ALL=C sed -i '1i extern "C" void cmref_hook(int,int,int,int);' ContigsCompactor.cpp
POBJS=""
for u in $UNITS; do $CXX -O3 -w -fPIC -c "$u.cpp" -o "$u.pic.o" & POBJS="$POBJS $u.pic.o"; done
$CXX -O3 -w -fPIC -I"$TMP" -c "$HERE/ref_harness.cpp" -o ref_harness.pic.o &
wait
$CXX -shared -o "$OUT/libcm_ref.so" $POBJS ref_harness.pic.o -lz -lm -lpthread
# (c) TERefiner's affine local aligner: one self-contained source file, no patch
LA="$REF/TERefiner/algorithms"
$CXX -O3 -w -fPIC -shared -I"$LA" -o "$OUT/libla_ref.so" "$LA/local_alignment.cpp" "$HERE/la_harness.cpp"
echo "built $OUT/ContigsMerger, $OUT/libcm_ref.so and $OUT/libla_ref.so"
