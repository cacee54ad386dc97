#!/bin/bash
# GPU job D: full GPU test-suite (stationary, non-stationary, TorchGate), smoke, bench, kernel launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/d_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/d_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*_|k_" -c 40 --csv --log-file gpurun_out/d_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/d_ncu_launch.log 2>&1
tail -n 3 gpurun_out/d_pytest.log gpurun_out/d_bench.log
