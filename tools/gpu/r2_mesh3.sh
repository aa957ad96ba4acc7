#!/bin/bash
# mesh front end: compacted insertion in the leaf step; A/B of the node-loop form of the neighbour search (NTX_MESH_NODE_LOOP)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_mesh3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_mesh.py -q --timeout 300 > $O/pytest_mesh.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mesh.log
tail -3 $O/pytest_mesh.log
timeout 600 python tools/bench_mesh.py --cpu-samples 0 > $O/bench_mesh.log 2>&1; tail -1 $O/bench_mesh.log
NTX_MESH_NODE_LOOP=1 timeout 600 python -m pytest tests/test_gpu_mesh.py -q --timeout 300 -k "knn or project" > $O/pytest_mesh_node_loop.log 2>&1; tail -1 $O/pytest_mesh_node_loop.log
NTX_MESH_NODE_LOOP=1 timeout 600 python tools/bench_mesh.py --cpu-samples 0 > $O/bench_mesh_node_loop.log 2>&1; tail -1 $O/bench_mesh_node_loop.log
