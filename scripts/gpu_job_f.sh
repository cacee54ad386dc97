#!/bin/bash
# GPU job B: parity tests + bench after the code-size restructure, one ncu capture of k2/k1.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/f_bench.log
for k in k2_synthesize k1_analyze; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/f_${k}_full -f \
      python bench.py --steps 1 --warmup 1 --minutes 2 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_$k.log 2>&1
done
tail -n 3 gpurun_out/f_pytest.log gpurun_out/f_bench.log
