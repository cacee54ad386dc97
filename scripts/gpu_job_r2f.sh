#!/bin/bash
# Round-2 job F (1 GPU): restructured k1d timing, config-3 per-kernel launch times, pageable-host probes.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2f_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2f_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k1n|k_iir|k_smooth|k2_syn|k2c" -c 8 --csv --log-file gpurun_out/r2f_cfg3_launches.csv \
    python scripts/time_configs.py 3 > gpurun_out/r2f_cfg3.log 2>&1
timeout 300 python scripts/probe_host_register.py > gpurun_out/r2f_hostreg.log 2>&1
grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2f_bench.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2f_bench.log | head -1
cut -d, -f5,12- gpurun_out/r2f_cfg3_launches.csv | tail -8; cat gpurun_out/r2f_hostreg.log
