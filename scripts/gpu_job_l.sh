#!/bin/bash
# GPU job L: full GPU tests incl. batching/config-5 geometry, bench, TorchGate config-4 timing.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/l_bench.log
timeout 300 python scripts/time_config4.py > gpurun_out/l_config4.log 2>&1
tail -n 3 gpurun_out/l_pytest.log gpurun_out/l_config4.log
