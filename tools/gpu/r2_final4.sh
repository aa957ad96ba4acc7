#!/bin/bash
# final check of the committed state: smoke, the whole GPU suite, bench N=1 (with the mesh / product-model extras), bench --impl reference
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_final4; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final4/bench_n1.json')); print('N=1 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['gpu_launches']); print('mesh',json.dumps(d.get('mesh'))[:1800])
PY
