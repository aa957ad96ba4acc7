#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_peer2; mkdir -p $O
timeout 700 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -12 $O/pytest.log
for ex in peer nccl; do
NTX_EXCHANGE=$ex timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > $O/bench_n2_$ex.json 2> $O/bench_n2_$ex.err
python - $O/bench_n2_$ex.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1],'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),d['config']['workload'][-90:]); print('   ',d['per_rank_ms'][0])
except Exception as e: print('FAILED',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
