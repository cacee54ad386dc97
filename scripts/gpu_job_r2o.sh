#!/bin/bash
# Round-2 job O (1 GPU): GPU tests of the committed state, config-3 timing (smoothing prefetch, trimmed follower), where the
# pageable numpy path spends its time, default bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
timeout 300 python scripts/ab_config3.py 0 128 > gpurun_out/r2o_ab_config3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 40 --csv --log-file gpurun_out/r2o_launches_config3.csv \
    python scripts/ab_config3.py 0 > gpurun_out/r2o_ncu_launch3.log 2>&1
timeout 300 python scripts/trace_numpy_path.py 4 > gpurun_out/r2o_numpy_path.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r2o_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2o_bench.log
tail -n 3 gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_ab_config3.log gpurun_out/r2o_numpy_path.log | tail -8; tail -n 3 gpurun_out/r2o_bench.log | cut -c1-2500
