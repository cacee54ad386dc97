#!/bin/bash
# Round-2 job N (1 GPU): validation + evidence of the committed state: smoke, GPU tests, default bench, config-3 A/B and
# launch list, config-2 launch list, full ncu captures of the four config-3 kernels.
set -x
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled > gpurun_out/r2n_host.txt 2>&1; nproc >> gpurun_out/r2n_host.txt; free -g >> gpurun_out/r2n_host.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
timeout 400 python scripts/ab_config3.py 0 128 16 2 > gpurun_out/r2n_ab_config3.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r2n_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2n_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*d?_|k_" -c 40 --csv --log-file gpurun_out/r2n_launches_config3.csv \
    python scripts/ab_config3.py 0 > gpurun_out/r2n_ncu_launch3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*d?_|k_" -c 30 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2n_ncu_launch.log 2>&1
for k in k_smooth_box k_iir_sigmoid k2c_synthesize_2k k1nd_magnitude_2k; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2n_${k}_full -f \
      python scripts/ab_config3.py 0 > gpurun_out/r2n_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2n_pytest.log gpurun_out/r2n_smoke.log; cat gpurun_out/r2n_ab_config3.log | tail -5; tail -n 3 gpurun_out/r2n_bench.log | cut -c1-3000
