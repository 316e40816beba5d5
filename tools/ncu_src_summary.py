"""Summarises `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp-instructions by
opcode, stall-sample totals by reason, and the hottest instructions.  usage: ncu_src_summary.py file.csv [top]"""
import csv
import collections
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    ops = collections.Counter()
    total = 0
    stalls = collections.Counter()
    samples = 0
    for r in body:
        n = int(r[col["Instructions Executed"]] or 0)
        src = r[col["Source"]].strip()
        parts = src.split()
        op = parts[0] if parts else "?"
        if op.startswith("@") and len(parts) > 1:
            op = parts[1]
        ops[op.split(".")[0]] += n
        total += n
        samples += int(r[col["# Samples"]] or 0)
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stalls[h] += int(r[col[h]] or 0)
    print(f"total warp-instructions executed: {total:.4g}; samples {samples}")
    print("by opcode:", ", ".join(f"{o} {100.0 * n / total:.1f}%" for o, n in ops.most_common(24)))
    print("stall samples:", ", ".join(f"{k[6:]} {100.0 * v / max(samples, 1):.1f}%" for k, v in stalls.most_common(10)))
    print("hottest instructions (samples, executed, sass):")
    for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
        print(f"  {r[col['# Samples']]:>7} {r[col['Instructions Executed']]:>12}  {r[col['Source']].strip()[:110]}")


if __name__ == "__main__":
    main()
