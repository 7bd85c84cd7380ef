"""Top stall-sample lines of one kernel from `ncu -i rep --page source --csv --launch-skip K --launch-count 1`.
usage: python tools/ncu_source_top.py <source.csv> [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sections, cur = [], None
for r in rows:
    if r and r[0] in ("Address", "#", "Line") or (r and "Source" in r and "# Samples" in r):
        cur = {"h": r, "d": []}
        sections.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["d"].append(r)
for sec in sections:
    idx = {c: i for i, c in enumerate(sec["h"])}
    si, ii, so = idx["# Samples"], idx.get("Instructions Executed"), idx["Source"]
    def val(r, i):
        try:
            return int(r[i])
        except Exception:
            return 0
    tot = sum(val(r, si) for r in sec["d"])
    print(f"--- section first column '{sec['h'][0]}', {len(sec['d'])} lines, {tot} samples")
    for r in sorted(sec["d"], key=lambda r: -val(r, si))[:n]:
        print(f"{val(r, si):7d} {100.0 * val(r, si) / max(tot, 1):5.1f}%  {r[ii] if ii is not None else '':>9}  {r[so][:120]}")
