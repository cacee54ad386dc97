#!/bin/bash
# Round-2 job D: full default bench (parity, configs_extra, e2e, e2e_numpy, cpu_baseline), reference arm, GPU tests,
# configs 3/4 with the streaming smoothing / regenerating follower.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/r2d_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2d_bench.log
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2d_bench_ref.log 2>&1
timeout 600 python scripts/time_configs.py 3 4 > gpurun_out/r2d_configs.log 2>&1
nproc > gpurun_out/r2d_host.txt; free -g >> gpurun_out/r2d_host.txt; df -h /dev/shm >> gpurun_out/r2d_host.txt
tail -n 3 gpurun_out/r2d_pytest.log; tail -n 6 gpurun_out/r2d_bench.log | cut -c1-3000; tail -n 5 gpurun_out/r2d_bench_ref.log | cut -c1-1500; tail -n 2 gpurun_out/r2d_configs.log
