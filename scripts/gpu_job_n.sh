#!/bin/bash
# GPU job N: GPU tests + bench after dtype templating.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n_bench.log
tail -n 3 gpurun_out/n_pytest.log
