#!/bin/sh
# Builds oracle/_ref/libref_{B,blocks}.so: the reference's OWN element code
# (MFEM/mechanic2d/asym_elasto_damage_model.cc: the prologue with the macros and asym_stress,
# and class damIntegrator) compiled from where it lies under the reference tree against
# the minimal MFEM stand-in of this directory.  Nothing of the reference is copied
# into the repository: its lines are streamed into the compiler.
# The two pieces are cut out by MARKERS, not by line numbers (file start .. the comment that opens the
# output-projection classes; the comment that opens damIntegrator .. the line before main's doc comment), and
# the file's SHA-256 is checked against the one the goldens were generated from: a reference tree that moved
# produces a warning and NO library (the committed goldens in tests/golden/ stay the pin).
#   variant B      : as shipped (#define USE_B, B.D.B^t product, M.cc:699-704,886-887)
#   variant blocks : USE_B commented out (tensor-product blocks, M.cc:705-717,893-911)
set -e
REF=${1:-/root/reference}
SRC=$REF/MFEM/mechanic2d/asym_elasto_damage_model.cc
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../_ref
CXX=${CXX:-g++}
[ -x /usr/bin/g++ ] && CXX=/usr/bin/g++
[ -f "$SRC" ] || { echo "oracle/_ref: $SRC not found, skipped"; exit 0; }
mkdir -p "$OUT"
WANT=a9f33df4401b2e09c0ab1b02118943d09c543474972d3d56af75c61e54efb6cc
HAVE=$(sha256sum "$SRC" | cut -d' ' -f1)
if [ "$HAVE" != "$WANT" ]; then
   echo "oracle/_ref: $SRC has SHA-256 $HAVE, the goldens were generated from $WANT: not built"; exit 0
fi
extract() {
   awk '/^\/\/ to project strain on vectorial DGspace for output/ {part = 1}
        /^\/\/ Non linear integrator to compute an asymmetric traction\/compression damaged elasticity law:/ {part = 2}
        /^\/\/\/ \\brief test program/ {part = 3}
        part != 1 && part != 3 {print}' "$SRC"
}
N=$(extract | wc -l)
[ "$N" -gt 700 ] || { echo "oracle/_ref: markers not found in $SRC ($N lines extracted): not built"; exit 0; }
FLAGS="-x c++ -std=c++17 -O3 -DNDEBUG -fPIC -shared -I$HERE -w"   # the reference flags (MFEM/setting.mk.in:4)
( extract; cat "$HERE/ref_driver.cc" ) | $CXX $FLAGS -o "$OUT/libref_B.so" -
( extract | sed 's|^#define USE_B$|//#define USE_B|'; cat "$HERE/ref_driver.cc" ) | $CXX $FLAGS -o "$OUT/libref_blocks.so" -
echo "oracle/_ref: built libref_B.so libref_blocks.so from $SRC"
