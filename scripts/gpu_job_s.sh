#!/bin/bash
# GPU job S (2 GPUs): N=1 and N=2 bench with the spectrum cache (driver-style launch), plus GPU tests.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s_bench2.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/s_bench2.log
tail -n 2 gpurun_out/s_pytest.log
