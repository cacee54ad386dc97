"""Build a compile-time variant of libb200gate for A/B runs:  python scripts/build_variant.py NAME -DFOO=1 ...
-> noisereduce_b200/libb200gate_NAME.so (git-ignored; travels with gpurun).  scripts/ab_variants.py times them."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisereduce_b200.csrc import build as B  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "noisereduce_b200", f"libb200gate_{name}.so")
cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs + \
      ["-I", os.path.join(ROOT, "include"), "-I", B.HERE, "-o", out] + [os.path.join(B.HERE, s) for s in B.SOURCES]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    raise SystemExit(1)
print(out)
