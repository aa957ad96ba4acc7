#!/bin/bash
# round-2 final single-GPU session: smoke, full parity suite, full bench line, ncu captures for profiles/ (tools/ncu_summary.py r2_final r02)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_final; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 300 $O/bench_reference_arm.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_frame.csv python tools/profile_frame.py frame > $O/ncu_frame.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ngp_field --launch-skip 1 --launch-count 1 -o $O/prof_field_frame python tools/profile_frame.py frame > $O/ncu_field.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:march_rays_compact --launch-skip 0 --launch-count 2 -o $O/prof_march_frame python tools/profile_frame.py frame > $O/ncu_march.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:mlp_pipe|grid_fwd_pair" -c 2 -o $O/prof_cfg2 python tools/profile_frame.py cfg2 > $O/ncu_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:wgrad_tc|grid_bwd|dgrad" -s 12 -c 6 -o $O/prof_cfg5 python tools/bench_cfg5.py --backend ntx --iters 2 > $O/ncu_cfg5.log 2>&1
NTX_LIB_PATH=$PWD/nerf_texture_b200/lib/libntx_probe.so timeout 400 python tools/field_probe.py 0 1 2 > $O/probe.log 2>&1; cat $O/probe.log
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ngp_field --csv --log-file $O/field_dram.csv python tools/profile_frame.py frame > $O/ncu_dram.log 2>&1
python - <<'PY'
import csv,json
rows=[r for r in csv.reader(open('gpurun_out/r2_final/field_dram.csv')) if len(r)>5 and r[0].isdigit()]
tot=0; per={}
for r in rows:
    if 'dram__bytes' in r[-3]:
        v=float(r[-1]); mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(r[-2],1)
        tot+=v*mult; per[r[0]]=per.get(r[0],0)+v*mult
json.dump({"dram_bytes_all_field_launches_of_one_frame": tot, "per_launch": per, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:ngp_field python tools/profile_frame.py frame (one steady-state frame, all its field-kernel launches)"}, open('gpurun_out/r2_final/field_kernel_traffic.json','w'), indent=1)
PY
