"""Instruction mix and stall samples per SASS opcode class from an ncu report (--page source --csv)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Warp Stall Sampling (Not-issued Samples)")}
inst = collections.Counter()
samp = collections.Counter()
nis = collections.Counter()
for r in rows[hdr_i + 1:]:
    if len(r) <= max(ci.values()):
        continue
    src = r[ci["Source"]].strip()
    toks = src.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    if not toks:
        continue
    op = toks[0].split(".")[0]
    full = toks[0]
    cls = op
    if op in ("FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2"):
        cls = "FP32 " + op
    elif op in ("LDS", "STS"):
        cls = "smem " + full.split(".")[0] + ("." + full.split(".")[-1] if "." in full else "")
    elif op in ("LDG", "STG", "LDGSTS", "LDGDEPBAR"):
        cls = "global " + op
    inst[cls] += int(float(r[ci["Instructions Executed"]] or 0))
    samp[cls] += int(float(r[ci["# Samples"]] or 0))
    nis[cls] += int(float(r[ci["Warp Stall Sampling (Not-issued Samples)"]] or 0))
ti, ts = sum(inst.values()), sum(samp.values())
print(f"{rep}: {ti} warp-instructions, {ts} samples")
for k, v in inst.most_common(28):
    print(f"  {k:22s} inst {v:12d} {100*v/ti:5.1f}%   samples {100*samp[k]/max(ts,1):5.1f}%  not-issued {100*nis[k]/max(ts,1):5.1f}%")
