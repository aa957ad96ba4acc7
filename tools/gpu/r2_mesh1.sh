#!/bin/bash
# mesh front end, first GPU session: parity tests, smoke, the bench tool, launch list + one ncu --set full of the fused projection
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_mesh1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_mesh.py -q --timeout 300 > $O/pytest_mesh.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mesh.log
tail -25 $O/pytest_mesh.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python tools/bench_mesh.py > $O/bench_mesh.log 2>&1; tail -3 $O/bench_mesh.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_ -c 3 -o $O/mesh_full python tools/bench_mesh.py --profile > $O/ncu.log 2>&1
ncu -i $O/mesh_full.ncu-rep --page raw --csv > $O/mesh_full_raw.csv 2>/dev/null
ls -la $O | tail -8
