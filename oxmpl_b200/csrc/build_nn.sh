#!/bin/sh
# Compiles oxg_nn.cu.  nvcc's --split-compile (which also cuts this file's compile time from
# ~160 s to ~80 s) partitions the translation unit differently from run to run, and ptxas's
# register allocation of knn_rv_dot_kernel<7,2> is bimodal under it: 135-144 registers
# (measured 342 ms for the 1M x 1M 7-D scan) or 160-168 registers (390-400 ms), from the same
# source.  Every attempt to pin the better allocation from the source side (register caps,
# separate relocatable translation unit for the exact path, indirect calls) produced the worse
# one plus spills, so this script simply keeps the better of a few compiles.  Results are
# bit-identical either way; only the speed differs.
#   usage: build_nn.sh <nvcc> <out.o> <flags...>
NVCC="$1"; OUT="$2"; shift 2
CUOBJDUMP="$(dirname "$NVCC")/cuobjdump"
KERNEL="knn_rv_dot_kernelILi7ELi2"
best=999
tries="${OXG_NN_BUILD_TRIES:-4}"
i=1
while [ "$i" -le "$tries" ]; do
  "$NVCC" "$@" --split-compile=0 -c -o "$OUT.try" oxg_nn.cu || exit 1
  regs=$("$CUOBJDUMP" -res-usage "$OUT.try" 2>/dev/null | grep -A1 "$KERNEL" | sed -n 's/.*REG:\([0-9]*\).*/\1/p' | head -1)
  [ -z "$regs" ] && regs=998
  echo "build_nn.sh: try $i: $KERNEL uses $regs registers"
  if [ "$regs" -lt "$best" ]; then best="$regs"; mv -f "$OUT.try" "$OUT"; fi
  [ "$best" -le 150 ] && break
  i=$((i + 1))
done
rm -f "$OUT.try"
echo "build_nn.sh: kept the $best-register build"
