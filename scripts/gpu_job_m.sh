#!/bin/bash
# GPU job M: A/B of the k2 lock-step variant (alternative nvcc build selected with B200GATE_LIB).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/m_bench_base.log 2>&1
B200GATE_LIB=$PWD/scripts/microbench/libb200gate_lockstep.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/m_bench_lockstep.log 2>&1
tail -c 300 gpurun_out/m_bench_lockstep.log
