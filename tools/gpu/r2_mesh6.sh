#!/bin/bash
# mesh front end: queries visited in Morton order (Python-side ordering, same kernels): parity + timing
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_mesh6; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_mesh.py -q --timeout 300 > $O/pytest_mesh.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mesh.log
tail -4 $O/pytest_mesh.log
timeout 600 python tools/bench_mesh.py --cpu-samples 0 > $O/bench_mesh.log 2>&1; tail -1 $O/bench_mesh.log
