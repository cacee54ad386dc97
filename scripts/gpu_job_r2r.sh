#!/bin/bash
# Round-2 job R (1 GPU): final validation of the committed state: smoke, GPU tests, configs 3 / 4 timing, default bench,
# reference arm, launch lists of configs 2 / 3 / 4.
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2r_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 300 python scripts/time_configs.py 3 4 > gpurun_out/r2r_time_configs.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r2r_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2r_bench.log
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2r_bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/r2r_bench_ref.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 30 --csv --log-file gpurun_out/r2r_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2r_ncu_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 24 --csv --log-file gpurun_out/r2r_launches_config3.csv \
    python scripts/ab_config3.py 0 > gpurun_out/r2r_ncu_launch3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 140 --csv --log-file gpurun_out/r2r_launches_config4.csv \
    python scripts/time_configs.py 4 > gpurun_out/r2r_ncu_launch4.log 2>&1
tail -n 3 gpurun_out/r2r_pytest.log gpurun_out/r2r_smoke.log; tail -1 gpurun_out/r2r_time_configs.log; tail -n 3 gpurun_out/r2r_bench.log | cut -c1-2500; tail -n 4 gpurun_out/r2r_bench_ref.log | cut -c1-1200
