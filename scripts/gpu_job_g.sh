#!/bin/bash
# GPU job G: f32x2 microbenchmark + bench after reverting the k1 pre-load.
set -x
mkdir -p gpurun_out
./scripts/microbench/f32x2 > gpurun_out/g_f32x2.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/g_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/g_bench.log
cat gpurun_out/g_f32x2.log
