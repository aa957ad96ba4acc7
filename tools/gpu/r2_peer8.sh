#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_peer8; mkdir -p $O
for ex in nccl peer nccl peer; do
NTX_EXCHANGE=$ex timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/bench_n8_$ex.json 2> $O/bench_n8_$ex.err
python - $O/bench_n8_$ex.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1],'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3)); print('   ',d['per_rank_ms'][0])
except Exception as e: print('FAILED',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
