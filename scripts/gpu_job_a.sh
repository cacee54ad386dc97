#!/bin/bash
# GPU job A (round 1): smoke + parity tests + first bench + ncu launch list + ncu --set full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/a_gpu.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/a_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 2 --warmup 1 --minutes 2 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_launch.log 2>&1
for k in k2_synthesize k1_analyze k_smooth; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/a_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --minutes 2 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_$k.log 2>&1
done
tail -3 gpurun_out/a_smoke.log gpurun_out/a_pytest.log gpurun_out/a_bench.log
