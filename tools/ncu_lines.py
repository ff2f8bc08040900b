#!/usr/bin/env python
"""Per-source-line share of the executed warp instructions (and of the warp-state samples) of an .ncu-rep captured with
--import-source on:   python tools/ncu_lines.py x.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fname, hdr, lines = "", None, []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].strip().isdigit(): lines.append((fname, r))
def num(v):
    try: return float(v.replace(",", ""))
    except ValueError: return 0.0
ci, cx, cs = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tx = sum(num(r[cx]) for _, r in lines); ts = sum(num(r[ci]) for _, r in lines)
print(f"warp instructions {tx:.0f}, samples {ts:.0f};  columns: % instructions, % samples")
for f, r in sorted(lines, key=lambda fr: -num(fr[1][cx]))[:top]:
    print(f"{100*num(r[cx])/tx:5.1f} {100*num(r[ci])/ts:5.1f}  {f}:{r[0]}  {r[cs].strip()[:100]}")
