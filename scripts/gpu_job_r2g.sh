#!/bin/bash
# Round-2 job G (2 GPUs): the reworked k_peer_push at N=2 (reserve / CTA sweeps); k1d without its cache store (timing experiment).
set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e > gpurun_out/r2g_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_$tag.log; }
run store B200GATE_GATHER=store
run store_r24 B200GATE_GATHER=store B200GATE_RESERVE_SMS=24
run store_r6 B200GATE_GATHER=store B200GATE_RESERVE_SMS=6 B200GATE_PUSH_CTAS=24
B200GATE_DBG_NOCACHE_STORE=1 timeout 300 python scripts/ab_variants.py default > gpurun_out/r2g_nostore.log 2>&1
python - <<'PY'
import json
for t in ('store','store_r24','store_r6'):
    for l in open(f'gpurun_out/r2g_{t}.log'):
        if l.startswith('{'):
            d=json.loads(l); print(t, round(d['ms_per_step'],2), d['gather_verified'], d['roofline']['kernel_ms'])
PY
cat gpurun_out/r2g_nostore.log
