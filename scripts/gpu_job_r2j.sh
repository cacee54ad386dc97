#!/bin/bash
# Round-2 job J (8 GPUs): SMs given to k_peer_push at N=8 (32 / 48), config 5 with 40.
set -x
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 6 --warmup 3 --no-e2e > gpurun_out/r2j_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_$tag.log; }
run 8 n8_store_r32 B200GATE_GATHER=store B200GATE_RESERVE_SMS=32
run 8 n8_store_r48 B200GATE_GATHER=store B200GATE_RESERVE_SMS=48
env B200GATE_GATHER=store B200GATE_RESERVE_SMS=40 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    scripts/bench_config5.py --steps 1 --warmup 1 > gpurun_out/r2j_config5_n8_r40.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_config5_n8_r40.log
python - <<'PY'
import json
for t in ('n8_store_r32','n8_store_r48'):
    for l in open(f'gpurun_out/r2j_{t}.log'):
        if l.startswith('{'):
            d=json.loads(l); print(t, round(d['ms_per_step'],2), round(d['value']/1e9,1), d['gather_verified'], d['roofline']['kernel_ms'])
for l in open('gpurun_out/r2j_config5_n8_r40.log'):
    if l.startswith('{'):
        d=json.loads(l); print('config5 n8 r40', round(d['ms_per_step'],1), round(d['value']/1e9,1), d['checksums_agree_across_ranks'], d['all_gather_GBps_per_rank'], d['peak_memory_GB'])
PY
