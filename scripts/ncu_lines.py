#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample shares of one ncu report
(needs -lineinfo + --import-source on):  python scripts/ncu_lines.py rep.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
by_samples = len(sys.argv) > 3 and sys.argv[3] == "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, hdr = None, None
inst = collections.Counter()
samp = collections.Counter()
src = {}
line_key = None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))  # the second "Source" column (SASS) overwrites the first
    if r[0].strip():
        line_key = (cur_file, int(r[0]))
        src[line_key] = r[1].strip()
    if line_key is None or not r[2].strip():
        continue
    try:
        inst[line_key] += float(d["Instructions Executed"])
        samp[line_key] += float(d["# Samples"])
    except ValueError:
        pass
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
order = samp.most_common(top) if by_samples else inst.most_common(top)
for k, _ in order:
    v = inst[k]
    print(f"{k[0]}:{k[1]:<5d} inst {v / ti:6.3f}  samples {samp[k] / ts:6.3f}  {src.get(k, '')[:100]}")
