#!/bin/bash
# Round-2 job Q (1 GPU): GPU tests; config-3 A/B of the smoothing time loop / sqrt (default vs variant "oldwalk"); configs 3 / 4
# timing (time-split TorchGate statistics kernels); numpy path with its slab timeline; default bench; launch lists; full ncu
# captures of the kernels changed since job P.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
timeout 300 python scripts/ab_config3.py 0 v:oldwalk 0 v:oldwalk > gpurun_out/r2q_ab_config3.log 2>&1
timeout 300 python scripts/time_configs.py 4 > gpurun_out/r2q_time_configs.log 2>&1
B200GATE_TRACE=1 timeout 300 python scripts/trace_numpy_path.py 3 > gpurun_out/r2q_numpy_path.log 2> gpurun_out/r2q_numpy_trace.txt
( time timeout 900 python bench.py ) > gpurun_out/r2q_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2q_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 24 --csv --log-file gpurun_out/r2q_launches_config3.csv \
    python scripts/ab_config3.py 0 > gpurun_out/r2q_ncu_launch3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*[cd]?_|k_" -c 60 --csv --log-file gpurun_out/r2q_launches_config4.csv \
    python scripts/time_configs.py 4 > gpurun_out/r2q_ncu_launch4.log 2>&1
for k in k_smooth_box k1n_magnitude_2k; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2q_${k}_full -f \
      python scripts/ab_config3.py 0 > gpurun_out/r2q_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_ab_config3.log; tail -3 gpurun_out/r2q_numpy_path.log; tail -2 gpurun_out/r2q_time_configs.log; tail -n 3 gpurun_out/r2q_bench.log | cut -c1-2500
