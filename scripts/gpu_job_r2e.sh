#!/bin/bash
# Round-2 job E (2 GPUs): gather transports A/B at N=2: kernel-issued NVLink stores vs copy-engine pushes vs NCCL.
set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2e_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_$tag.log; tail -n 3 gpurun_out/r2e_$tag.log | cut -c1-600; }
run store B200GATE_GATHER=store
run store_r8 B200GATE_GATHER=store B200GATE_RESERVE_SMS=8
run store_r20 B200GATE_GATHER=store B200GATE_RESERVE_SMS=20
run peer B200GATE_GATHER=peer
run nccl B200GATE_GATHER=nccl
