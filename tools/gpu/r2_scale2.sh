#!/bin/bash
# N = 8 only, two settings of the frame loop's run-ahead depth (8-GPU box)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_scale_c; mkdir -p $O
for a in 1 3; do
  NTX_FRAME_AHEAD=$a timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + a)) bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/bench_n8_ahead$a.json 2> $O/bench_n8_ahead$a.err
  python - "$O/bench_n8_ahead$a.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1].split('/')[-1],'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3)); print('   ',d['per_rank_ms'][0])
PY
done
NTX_FRAME_AHEAD=3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras > $O/bench_n4_ahead3.json 2> $O/bench_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_scale_c/bench_n4_ahead3.json') if l.startswith('{')][-1]); print('n4 ahead3 ms',round(d['ms_per_step'],3))
PY
