"""Hottest SASS instructions (by stall samples) of an ncu report, with their executed counts."""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
c_src, c_s, c_i = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[hdr_i + 1:] if len(r) > c_i]
tot = sum(int(float(r[c_s] or 0)) for r in body)
idx = sorted(range(len(body)), key=lambda i: -int(float(body[i][c_s] or 0)))[:top]
for i in sorted(idx):
    r = body[i]
    print(f"{i:6d} {100*int(float(r[c_s]))/tot:5.2f}%  x{int(float(r[c_i])):>10d}  {r[c_src].strip()[:110]}")
