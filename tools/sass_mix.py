"""Instruction mix of one kernel from `ncu -i rep --page source --csv --print-source sass`:
executed warp instructions and stall samples per opcode.  Usage: python tools/sass_mix.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("MUFU", "F2F", "PRMT", "IMAD", "LOP3", "SHF", "FMNMX", "SEL", "FSEL")) and "." in op else "")
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    ex[op] += n; st[op] += s; tot += n
ts = sum(st.values())
print(f"total warp instructions {tot}, samples {ts}")
for op, n in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:16s} {n:12d} {100*n/tot:6.2f}%   samples {100*st[op]/max(ts,1):6.2f}%")
