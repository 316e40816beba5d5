#!/bin/sh
# Builds an A/B variant of liboxg.so that differs only in one translation unit's -D flags.
#   usage: tools/build_variant.sh <name> <file.cu> [-DFLAG=V ...]   ->  tools/bin/liboxg_<name>.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; SRC="$2"; shift 2
mkdir -p "$ROOT/tools/bin"
cd "$ROOT/oxmpl_b200/csrc"
OBJ="$ROOT/tools/bin/${SRC%.cu}_$NAME.o"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-Wall -ccbin /usr/bin/g++ "$@" -c -o "$OBJ" "$SRC"
OBJS=""
for f in oxg_api oxg_nn oxg_validity oxg_prm oxg_planner oxg_peak oxg_so3 oxg_tree oxg_dist oxg_query; do
  [ -f "$f.cu" ] || continue
  if [ "$f.cu" = "$SRC" ]; then OBJS="$OBJS $OBJ"; else OBJS="$OBJS $f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/tools/bin/liboxg_$NAME.so" $OBJS -lcudart -ccbin /usr/bin/g++
echo "built tools/bin/liboxg_$NAME.so"
