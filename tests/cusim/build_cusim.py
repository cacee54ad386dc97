"""TEST INFRASTRUCTURE: compile the product's .cu sources with g++ against the cusim shim.

Produces tests/cusim/_build/libb200gate_cusim.so with the same C ABI, running every kernel on the
CPU fiber simulator.  Used only by tests (kernel-logic checks without a GPU); the Python package
never loads it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "noisereduce_b200", "csrc")
OUTDIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUTDIR, "libb200gate_cusim.so")
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
       [os.path.join(HERE, f) for f in ("cusim.h", "cusim.cpp")] + [os.path.join(ROOT, "include", "b200gate.h")]


def build(force=False):
    os.makedirs(OUTDIR, exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DB200_CUSIM_BUILD",
           "-Wno-unknown-pragmas", "-pthread", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           "-x", "c++", os.path.join(CSRC, "gate_host.cu"), os.path.join(HERE, "cusim.cpp"), "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("cusim build failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
