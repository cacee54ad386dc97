#!/bin/bash
# GPU job O: tests + A/B bench of the fused single-pass kernel vs the two-pass path.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/o_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/o_k_fused_full -f \
      python bench.py --steps 1 --warmup 1 --minutes 2 --no-cpu-baseline --no-e2e > gpurun_out/o_ncu.log 2>&1
tail -n 3 gpurun_out/o_pytest.log
