#!/usr/bin/env python
"""Writes profiles/r02_sass_hot_kernels.txt: instruction count, opcode histogram and identifying
mnemonics of the hot kernels of oxmpl_b200/liboxg.so (cuobjdump -sass / -res-usage, sm_100a)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oxmpl_b200", "liboxg.so")
KERNELS = [
    ("tree_knn_kernel<7,false>", "_ZN3oxg15tree_knn_kernelILi7ELb0EEEvNS_13TreeKnnParamsE"),
    ("prm_row_cull_kernel<7>", "_ZN3oxg19prm_row_cull_kernelILi7EEEvNS_13RowCullParamsE"),
    ("motion_cull_kernel<7,4>", "_ZN3oxg18motion_cull_kernelILi7ELi4EEEvNS_12MotionParamsE"),
    ("knn_rv_dot_kernel<7,2>", "_ZN3oxg17knn_rv_dot_kernelILi7ELi2EEEvNS_10ScanParamsE"),
]
MARKS = ["CREDUX", "REDUX", "VOTE", "DADD", "DMUL", "DFMA", "FADD.RM", "FMUL.RM", "F2F", "FMNMX3", "FFMA2",
         "FADD2", "FMUL2", "LDG.E.64.CONSTANT", "LDS.128", "SHFL", "UBLKCP", "SYNCS", "MUFU", "LDL", "STL"]
HEADER = """SASS of the hot kernels of liboxg.so (cuobjdump -sass, sm_100a), round 2, final build (tools/sass_summary.py).
Per kernel: instruction count, opcode histogram, and the first occurrence of the mnemonics that
identify the design (CREDUX = __reduce_min_sync child selection, VOTE = ballots, DADD/DMUL without DFMA =
separately rounded reference-order sums, F*.RM/.RP = directed-rounding certified box bounds, FFMA2/FADD2 =
packed FP32 pre-filters, UBLKCP = bulk (TMA-engine) tile copies, SYNCS = mbarrier pipeline).
DFMA never appears in a result-defining sum: the occurrences are the Newton steps of the correctly rounded
sqrt / division expansions (DFMA R, R, -R, 1 ...) the compiler emits for sqrt() and '/'.
"""


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    out = [HEADER]
    for label, sym in KERNELS:
        m = re.search(r"Function : " + re.escape(sym) + r"\n(.*?)(?=\n\s*Function : |\Z)", sass, re.S)
        if not m:
            out.append(f"== {label}: not found\n")
            continue
        usage = ""
        for i, line in enumerate(res):
            if sym in line and i + 1 < len(res):
                usage = res[i + 1].strip()
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(.*?);", m.group(1))
        ops = collections.Counter()
        first = {}
        for text in ins:
            parts = text.split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            ops[op.split(".")[0]] += 1
            for mk in MARKS:
                if op == mk or op.startswith(mk + ".") or (mk.count(".") and op.startswith(mk)):
                    first.setdefault(mk, [0, text])
                    first[mk][0] += 1
        out.append(f"== {label}  ({len(ins)} instructions)  {usage}")
        out.append("opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
        for mk in MARKS:
            if mk in first:
                out.append(f"  {mk:<18} x {first[mk][0]:4d}   e.g. {first[mk][1]} ;")
        out.append("")
    path = os.path.join(ROOT, "profiles", "r02_sass_hot_kernels.txt")
    open(path, "w").write("\n".join(out))
    print(path)


if __name__ == "__main__":
    sys.exit(main())
