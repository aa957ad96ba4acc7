#!/bin/bash
# final check of the committed state on a 2-GPU box: smoke, the whole GPU suite (incl. the 2-GPU NCCL / peer-memory tests), bench N=1 and N=2
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_final2; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final2/bench_n1.json')); print('N=1 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['gpu_launches'])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29670 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_final2/bench_n2.json') if l.startswith('{')][-1]); print('N=2 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['config']['workload'][-80:])
PY
