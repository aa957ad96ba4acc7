#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_mlp.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python tools/bench_cfg5.py --backend ntx > $O/cfg5_ntx.log 2>&1; tail -1 $O/cfg5_ntx.log | cut -c1-700
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ngp_field --csv --log-file $O/field_dram.csv python tools/profile_frame.py frame > $O/ncu_dram.log 2>&1
python - <<'PY'
import csv,json
rows=[r for r in csv.reader(open('gpurun_out/r2_12/field_dram.csv')) if len(r)>5 and r[0].isdigit()]
tot=0; per={}
for r in rows:
    if 'dram__bytes' in r[-3]:
        v=float(r[-1]); u=r[-2]
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
        tot+=v*mult; per.setdefault(r[0],0); per[r[0]]+=v*mult
json.dump({"dram_bytes_all_field_launches_of_one_frame": tot, "per_launch": per, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:ngp_field python tools/profile_frame.py frame (one steady-state frame, all its field-kernel launches)"}, open('gpurun_out/r2_12/field_kernel_traffic.json','w'), indent=1)
print(tot, per)
PY
