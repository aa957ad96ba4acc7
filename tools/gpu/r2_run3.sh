#!/bin/bash
# round-2 GPU session 3: density-grid update parity, reference-files test (both arms), cfg5 launch list, compat numbers
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_3; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_density.py tests/test_gpu_reference_files.py tests/test_gpu_training.py tests/test_gpu_grid.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
for b in ntx ref; do timeout 300 python tools/bench_cfg5.py --backend $b > $O/cfg5_$b.log 2>&1; tail -1 $O/cfg5_$b.log | cut -c1-900; done
for b in ntx ref; do timeout 600 python tools/run_reference_files.py --backend $b --size 1024 --time 5 > $O/compat_$b.log 2>&1; tail -1 $O/compat_$b.log | cut -c1-600; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/launches_cfg5.csv python tools/bench_cfg5.py --backend ntx --iters 3 > $O/ncu_cfg5.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_3/launches_cfg5.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:40]: print(r[4][:70].ljust(70), r[-1])
PY
