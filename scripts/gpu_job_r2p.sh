#!/bin/bash
# Round-2 job P (1 GPU): validation + evidence of the committed state: smoke, GPU tests, numpy-path breakdown, configs 2/3/4
# timing, default bench, launch lists (config 2, 3, 4), full ncu captures of the config-3 kernels (full batch).
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2p_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
timeout 300 python scripts/trace_numpy_path.py 4 > gpurun_out/r2p_numpy_path.log 2>&1
timeout 300 python scripts/time_configs.py 3 4 > gpurun_out/r2p_time_configs.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/r2p_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2p_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 40 --csv --log-file gpurun_out/r2p_launches_config3.csv \
    python scripts/ab_config3.py 0 > gpurun_out/r2p_ncu_launch3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 60 --csv --log-file gpurun_out/r2p_launches_config4.csv \
    python scripts/time_configs.py 4 > gpurun_out/r2p_ncu_launch4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 30 --csv --log-file gpurun_out/r2p_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2p_ncu_launch.log 2>&1
for k in k_smooth_box k_iir_sigmoid k2c_synthesize_2k k1n_magnitude_2k; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2p_${k}_full -f \
      python scripts/ab_config3.py 0 > gpurun_out/r2p_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2p_pytest.log gpurun_out/r2p_smoke.log; tail -4 gpurun_out/r2p_numpy_path.log; tail -3 gpurun_out/r2p_time_configs.log; tail -n 3 gpurun_out/r2p_bench.log | cut -c1-2500
