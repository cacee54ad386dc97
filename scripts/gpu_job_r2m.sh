#!/bin/bash
# Round-2 job M (1 GPU): evidence for the committed state: smoke, GPU tests, default bench, launch list, full ncu captures.
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2m_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r2m_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2m_bench.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k[0-9n]*d?_|k_" -c 30 --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2m_ncu_launch.log 2>&1
for k in k2d_synthesize k1d_analyze k_smooth_packed; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2m_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2m_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/r2m_pytest.log gpurun_out/r2m_smoke.log; tail -n 5 gpurun_out/r2m_bench.log | cut -c1-3000
