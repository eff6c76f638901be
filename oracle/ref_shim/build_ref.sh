#!/bin/sh
# Builds oracle/_ref/libref_{B,blocks}.so: the reference's OWN element code
# (MFEM/mechanic2d/asym_elasto_damage_model.cc lines 1-330: macros + asym_stress, and
# 487-953: damIntegrator) compiled from where it lies under the reference tree against
# the minimal MFEM stand-in of this directory.  Nothing of the reference is copied
# into the repository: its lines are streamed into the compiler.
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
FLAGS="-x c++ -std=c++17 -O3 -DNDEBUG -fPIC -shared -I$HERE -w"   # the reference flags (MFEM/setting.mk.in:4)
( sed -n '1,330p;487,953p' "$SRC"; cat "$HERE/ref_driver.cc" ) | $CXX $FLAGS -o "$OUT/libref_B.so" -
( sed -n '1,330p;487,953p' "$SRC" | sed 's|^#define USE_B$|//#define USE_B|'; cat "$HERE/ref_driver.cc" ) | $CXX $FLAGS -o "$OUT/libref_blocks.so" -
echo "oracle/_ref: built libref_B.so libref_blocks.so from $SRC"
