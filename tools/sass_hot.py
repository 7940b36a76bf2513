"""Executed-instruction mix of a kernel from an ncu report's source page: which opcodes the warps actually issue.
Usage: python tools/sass_hot.py prof.ncu-rep [top N opcodes]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stall = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", src)
    if not m:
        continue
    n = int(r[ix["Instructions Executed"]] or 0)
    op = m.group(1)
    ops[op] += n
    stall[op] += int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    total += n
print(f"{total} warp instructions executed")
ts = sum(stall.values())
for op, n in ops.most_common(top):
    print(f"{op:12s} {n:12d}  {100.0 * n / total:5.1f} %   stall samples {100.0 * stall[op] / max(ts, 1):5.1f} %")
