#!/bin/bash
# Round-2 job I (8 GPUs): gather transports at N=8 and N=4 (config 2 per rank), config 5 at N=8.
set -x
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 6 --warmup 3 --no-e2e > gpurun_out/r2i_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_$tag.log; }
run 8 n8_store B200GATE_GATHER=store B200GATE_RESERVE_SMS=16
run 8 n8_peer B200GATE_GATHER=peer
run 8 n8_nccl B200GATE_GATHER=nccl
run 4 n4_store B200GATE_GATHER=store B200GATE_RESERVE_SMS=16
run 4 n4_peer B200GATE_GATHER=peer
env B200GATE_GATHER=store B200GATE_RESERVE_SMS=16 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    scripts/bench_config5.py --steps 1 --warmup 1 > gpurun_out/r2i_config5_n8_store.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_config5_n8_store.log
python - <<'PY'
import json
for t in ('n8_store','n8_peer','n8_nccl','n4_store','n4_peer'):
    try:
        for l in open(f'gpurun_out/r2i_{t}.log'):
            if l.startswith('{'):
                d=json.loads(l); print(t, round(d['ms_per_step'],2), round(d['value']/1e9,1), d['gather_verified'], d['roofline']['kernel_ms'])
    except Exception as e: print(t, e)
for l in open('gpurun_out/r2i_config5_n8_store.log'):
    if l.startswith('{'):
        d=json.loads(l); print('config5 n8', round(d['ms_per_step'],1), round(d['value']/1e9,1), d['checksums_agree_across_ranks'], d['all_gather_GBps_per_rank'], d['peak_memory_GB'])
PY
tail -3 gpurun_out/r2i_config5_n8_store.log | cut -c1-400
