#!/bin/bash
# final check of the committed state with the mesh front end: smoke, the whole GPU suite, bench N=1 (with the mesh extras), launch list + ncu of the mesh kernels
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_final3; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final3/bench_n1.json')); print('N=1 ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['gpu_launches']); print('mesh',json.dumps(d.get('mesh'))[:1500])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mesh_ -c 60 --csv --log-file $O/launches_mesh.csv python tools/bench_mesh.py --cpu-samples 0 --iters 3 > $O/launches_mesh.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_ -c 3 -o $O/mesh_full python tools/bench_mesh.py --profile > $O/ncu.log 2>&1
ncu -i $O/mesh_full.ncu-rep --page raw --csv > $O/mesh_full_raw.csv 2>/dev/null
ls -la $O | tail -12
