#!/bin/bash
# GPU job I (4 GPUs): N=2 and N=4 torchrun bench with the overlapped all-gather.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/i_topo.txt 2>&1
for n in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/i_bench$n.log 2>&1; echo "bench$n rc=$?" >> gpurun_out/i_bench$n.log
done
tail -n 2 gpurun_out/i_bench2.log gpurun_out/i_bench4.log | cut -c1-600
