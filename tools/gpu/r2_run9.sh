#!/bin/bash
# round-2 GPU session 9: barrier-free marcher; shard sweep + launch list of a 1/8 shard (where does a rank's time go at N = 8)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_9; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_raymarching.py tests/test_gpu_compat.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python tools/shard_sweep.py > $O/shard_sweep.log 2>&1; cat $O/shard_sweep.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_shard8.csv python tools/profile_shard.py > $O/ncu_shard.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_9/launches_shard8.csv')) if len(r)>5 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split('(')[0][-40:]; agg.setdefault(k,[]).append(float(r[-1])/1000)
tot=0
for k,v in agg.items(): print(k.ljust(42), round(sum(v),1), [round(x) for x in v]); tot+=sum(v)
print('sum of kernels us', round(tot,1))
PY
timeout 300 python tools/tune.py "" > $O/tune.log 2>&1; cat $O/tune.log
