#!/bin/bash
# Round-2 job C: k1d variants (warps / staging) A/B; config 3 per-kernel launch times.
set -x
mkdir -p gpurun_out
timeout 400 python scripts/ab_variants.py default k1ns k1w12 > gpurun_out/r2c_ab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k1n|k_iir|k_smooth_f|k2_syn|k2c" -c 12 --csv --log-file gpurun_out/r2c_cfg3_launches.csv \
    python scripts/time_configs.py 3 > gpurun_out/r2c_cfg3.log 2>&1
cat gpurun_out/r2c_ab.log; cut -d, -f5,12- gpurun_out/r2c_cfg3_launches.csv | tail -14
