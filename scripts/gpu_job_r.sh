#!/bin/bash
# GPU job R: end-of-round validation: smoke, GPU tests, default bench (with cpu_baseline and e2e), reference arm,
# launch list and full-size ncu captures of the three main kernels.
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r_smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
timeout 900 python bench.py > gpurun_out/r_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r_bench.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*_|k_" -c 40 --csv --log-file gpurun_out/r_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r_ncu_launch.log 2>&1
for k in k2_synthesize k1_analyze k_smooth_packed; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r_pytest.log gpurun_out/r_smoke.log
