#!/bin/bash
# GPU job E (2 GPUs): full GPU tests (incl. n_fft=2048), N=1 bench with the pipelined host path, N=2 torchrun bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench1.log 2>&1; echo "bench rc=$?" >> gpurun_out/e_bench1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/e_bench2.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/e_bench2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 1 --warmup 1 --impl reference > gpurun_out/e_bench2_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/e_bench2_ref.log
tail -n 3 gpurun_out/e_pytest.log gpurun_out/e_bench1.log gpurun_out/e_bench2.log
