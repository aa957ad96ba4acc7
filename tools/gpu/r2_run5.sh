#!/bin/bash
# round-2 GPU session 5: per-ray useful-end precompute (no in-loop mip scans), fused epilogue, full ncu capture of field + march kernels
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_5; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 300 python tools/tune.py "" > $O/tune.log 2>&1; cat $O/tune.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_frame.csv python tools/profile_frame.py frame > $O/ncu_frame.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_5/launches_frame.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:60].ljust(60), r[-1])
PY
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ngp_field --launch-skip 1 --launch-count 1 -o $O/prof_field python tools/profile_frame.py frame > $O/ncu_field.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:march_rays_compact --launch-skip 0 --launch-count 2 -o $O/prof_march python tools/profile_frame.py frame > $O/ncu_march.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:mlp_pipe -c 1 -o $O/prof_mlp python tools/profile_frame.py cfg2 > $O/ncu_mlp.log 2>&1
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json
