#!/bin/bash
# round-2 GPU session 4: compact marcher + ray-mode field kernel + prekill: frame parity, timings, launch list
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_4; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 300 python tools/tune.py "" > $O/tune.log 2>&1; cat $O/tune.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_4/bench.json'))
print('ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'launches',d['gpu_launches'],'iters',d['config']['loop_iterations'])
r=d['roofline']; print('field avg us',r['avg_launch_us'],'frac',r['frac'],'march ms',r['march_kernel_ms'],'share',r['kernel_share_of_step'])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_frame.csv python tools/profile_frame.py frame > $O/ncu_frame.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_4/launches_frame.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:60].ljust(60), r[-1])
PY
