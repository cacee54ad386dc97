#!/usr/bin/env python
"""Static SASS accounting of a kernel's loops (no GPU needed).

    python scripts/sass_loop_mix.py <lib.so> <kernel-name-substring> [--loops N]

Dumps the kernel with cuobjdump, finds the loops (backward branches), and prints for each loop its instruction
count and opcode mix.  Nested loops are listed with their parent so a per-iteration dynamic count can be put
together by hand (e.g. pair loop + 2 x the FFT pass loop)."""
import collections
import re
import subprocess
import sys


def dump(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat in name:
            yield name, f


def parse(body):
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        text = m.group(2).strip()
        pred = ""
        mm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
        if mm:
            pred, text = mm.group(1), mm.group(2)
        op = text.split()[0]
        ins.append((addr, op, text, pred))
    return ins


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    for name, body in dump(lib, pat):
        ins = parse(body)
        print(f"== {name}: {len(ins)} instructions, {len(ins) * 16 / 1024:.1f} KB")
        loops = []
        for addr, op, text, pred in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt <= addr:
                        loops.append((tgt, addr))
        loops = sorted(set(loops), key=lambda l: (l[0], -l[1]))
        for lo, hi in loops:
            sel = [i for i in ins if lo <= i[0] <= hi]
            inner = [l for l in loops if l != (lo, hi) and lo <= l[0] and l[1] <= hi]
            mix = collections.Counter(re.sub(r"\..*", "", i[1]) for i in sel)
            fine = collections.Counter(i[1] for i in sel)
            print(f"-- loop 0x{lo:x}..0x{hi:x}: {len(sel)} instr; inner loops: " +
                  ", ".join(f"0x{a:x}..0x{b:x}" for a, b in inner))
            print("   " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(28)))
            if "--fine" in sys.argv:
                print("   " + "  ".join(f"{k}:{v}" for k, v in fine.most_common(40)))


if __name__ == "__main__":
    main()
