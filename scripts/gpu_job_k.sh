#!/bin/bash
# GPU job H: tests + bench after k_smooth batching and k1 at 16 warps/SM.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/k_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/k_bench.log
tail -n 2 gpurun_out/k_pytest.log
