#!/bin/bash
# mesh front end: 8-lanes-per-query neighbour search vs one thread per query (A/B via NTX_MESH_THREAD_KNN), parity tests, ncu of the new kernels
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_mesh2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_mesh.py -q --timeout 300 > $O/pytest_mesh.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mesh.log
tail -25 $O/pytest_mesh.log
timeout 600 python tools/bench_mesh.py > $O/bench_mesh.log 2>&1; tail -1 $O/bench_mesh.log
NTX_MESH_THREAD_KNN=1 timeout 600 python tools/bench_mesh.py --cpu-samples 0 > $O/bench_mesh_thread_knn.log 2>&1; tail -1 $O/bench_mesh_thread_knn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_ -c 3 -o $O/mesh_full python tools/bench_mesh.py --profile > $O/ncu.log 2>&1
ncu -i $O/mesh_full.ncu-rep --page raw --csv > $O/mesh_full_raw.csv 2>/dev/null
ls -la $O | tail -8
