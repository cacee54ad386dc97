#!/bin/bash
# Round-2 job A: new bulk-staged synthesis kernel (k2c) + async run: GPU tests, bench, A/B of k1 staging, configs 3/4 timing,
# full ncu captures of k2c and k1.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2a_bench.log
timeout 300 python scripts/ab_path_flags.py 0 8 2 > gpurun_out/r2a_ab.log 2>&1
timeout 600 python scripts/time_configs.py 3 4 > gpurun_out/r2a_configs.log 2>&1
for k in k2c_synthesize k1_analyze; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2a_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2a_pytest.log; tail -n 2 gpurun_out/r2a_bench.log; cat gpurun_out/r2a_ab.log; tail -n 3 gpurun_out/r2a_configs.log
