#!/bin/bash
# GPU job P: tests + bench with the spectrum cache; ncu of k1/k2 in that mode.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/q_bench.log
for k in k2_synthesize; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/q_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --minutes 2 --no-cpu-baseline --no-e2e > gpurun_out/q_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/q_pytest.log
