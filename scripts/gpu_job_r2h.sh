#!/bin/bash
# Round-2 job H (2 GPUs): SM-filling k_peer_push at N=2 (reserve sweep) vs copy engines; e2e_numpy with the staged pageable path.
set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2h_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_$tag.log; }
run store_r8 B200GATE_GATHER=store B200GATE_RESERVE_SMS=8
run store_r12 B200GATE_GATHER=store B200GATE_RESERVE_SMS=12
run store_r16 B200GATE_GATHER=store B200GATE_RESERVE_SMS=16
run peer B200GATE_GATHER=peer
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench1.log 2>&1
python - <<'PY'
import json
for t in ('store_r8','store_r12','store_r16','peer'):
    for l in open(f'gpurun_out/r2h_{t}.log'):
        if l.startswith('{'):
            d=json.loads(l); print(t, round(d['ms_per_step'],2), d['gather_verified'], d['roofline']['kernel_ms'])
for l in open('gpurun_out/r2h_bench1.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_numpy', d.get('e2e_numpy')); print(d.get('parity'))
PY
