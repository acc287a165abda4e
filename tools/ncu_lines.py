"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: python tools/ncu_lines.py <file.csv> [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
data = []
hdr = None
for r in rows:
  if len(r) >= 2 and r[0] == "File Path":
    cur = r[1].split("/")[-1]
    continue
  if len(r) > 8 and r[0] == "Line No":
    hdr = r
    ii = hdr.index("Instructions Executed")
    wi = hdr.index("Warp Stall Sampling (All Samples)")
    continue
  if hdr is None or len(r) <= ii:
    continue
  if r[0].isdigit():
    try:
      data.append((cur, int(r[0]), r[1].strip()[:100], int(r[ii] or 0), int(r[wi] or 0)))
    except ValueError:
      pass
ti = sum(d[3] for d in data) or 1
ts = sum(d[4] for d in data) or 1
print(f"total inst {ti}  stall samples {ts}")
for d in sorted(data, key=lambda x: -x[4])[:top]:
  print(f"{d[3] / ti * 100:5.1f}% inst {d[4] / ts * 100:5.1f}% stall  {d[0]}:{d[1]:<4d} {d[2]}")
