"""Executed warp instructions per CUDA source line of a kernel, from an ncu report captured with --import-source on.
Usage: python tools/src_lines.py prof.ncu-rep [top N lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
per, samp, src = collections.Counter(), collections.Counter(), {}
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-" and r[7].isdigit():
        k = (cur, int(r[0]))
        per[k] += int(r[7])
        samp[k] += int(r[4]) if r[4].isdigit() else 0
        src[k] = r[1].strip()
tot, ts = sum(per.values()), sum(samp.values())
print(f"{tot} warp instructions executed, {ts} stall samples")
for k, n in per.most_common(top):
    print(f"{k[0]}:{k[1]:<4d} {n:10d} {100 * n / tot:5.1f} %  stall {100 * samp[k] / max(ts, 1):5.1f} %  {src[k][:100]}")
