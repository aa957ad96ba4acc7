#!/bin/bash
# mesh kernels: threads per block A/B (NTX_MESH_BLOCK): the kernels are latency-bound, block retirement granularity decides how many warps stay resident
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_mesh4; mkdir -p $O
for b in 128 64 32 96; do
  NTX_MESH_BLOCK=$b timeout 300 python tools/bench_mesh.py --cpu-samples 0 --iters 7 > $O/bench_mesh_block$b.log 2>&1
  echo "block $b: $(python - <<PY
import json
l=[x for x in open('$O/bench_mesh_block$b.log') if x.startswith('RESULT ')]
d=json.loads(l[-1][7:]) if l else {}
print({k:round(d[k],3) for k in ('trace_ms','knn_ms','project_fused_ms','project_reference_chain_ms') if k in d})
PY
)"
done
