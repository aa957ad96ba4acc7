#!/bin/bash
# round-2 scaling run on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (the driver's SCALE launch line), per-rank breakdown in each line
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_scale${1:-}; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt 2>&1
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-extras > $O/bench_n$n.json 2> $O/bench_n$n.err
done
python - "$O" <<'PY'
import json,sys,os
O=sys.argv[1]; base=None
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open(os.path.join(O,'bench_n%d.json'%n)) if l.startswith('{')][-1])
    except Exception as e:
        print(n,'FAILED',e); continue
    if n==1: base=d['ms_per_step']
    print('N=%d ms %.3f e2e %.3f speedup %.2f iters %s' % (n,d['ms_per_step'],d['e2e']['ms_per_step'],base/d['ms_per_step'] if base else 0,d['config']['loop_iterations']))
    for r in d.get('per_rank_ms',[])[:2]: print('   ',r)
PY
