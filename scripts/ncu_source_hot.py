#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` dump (gzip'd, as scripts/ncu_box.sh writes them): executed-instruction
mix by opcode and the SASS lines with the most stall samples.   python scripts/ncu_source_hot.py <file.csv.gz> [top]"""
import csv
import gzip
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(gzip.open(path, "rt")))
    hdr = rows[1]
    isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    mix, total_exec, total_samp = Counter(), 0, 0
    lines = []
    for r in rows[2:]:
        if len(r) <= iexec:
            continue
        ex, sa = int(r[iexec] or 0), int(r[isamp] or 0)
        toks = r[isrc].split()
        op = toks[0] if toks else "?"
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        mix[op.split(".")[0]] += ex
        total_exec += ex
        total_samp += sa
        st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        lines.append((sa, ex, r[isrc].strip(), st))
    print(f"{rows[0][1][:100]}\nwarp instructions executed {total_exec}, samples {total_samp}")
    print("mix:", ", ".join(f"{k} {100 * v / total_exec:.1f}%" for k, v in mix.most_common(22)))
    print("hottest lines (samples, executed, SASS, top stalls):")
    for sa, ex, src, st in sorted(lines, reverse=True)[:top]:
        print(f"  {sa:6d} {ex:9d}  {src[:70]:70s} {st}")


if __name__ == "__main__":
    main()
