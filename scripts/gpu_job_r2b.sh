#!/bin/bash
# Round-2 job B: dual (two units per warp) analysis / synthesis kernels: GPU tests, bench, A/B against the single-unit
# kernels (path_flags 16), config 3 with the float32 follower, full ncu captures of k1d / k2d / k_smooth.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2b_bench.log
timeout 300 python scripts/ab_path_flags.py 0 16 > gpurun_out/r2b_ab.log 2>&1
timeout 600 python scripts/time_configs.py 3 > gpurun_out/r2b_configs.log 2>&1
for k in k2d_synthesize k1d_analyze k_smooth_packed; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2b_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2b_pytest.log; tail -n 2 gpurun_out/r2b_bench.log; cat gpurun_out/r2b_ab.log; tail -n 3 gpurun_out/r2b_configs.log
