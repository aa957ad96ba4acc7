#!/bin/bash
# round-2 GPU session 1: parity of the TMEM-operand kernels, A/B against the round-1 MLP, probe accounting, short bench
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python tools/tune.py "" "NTX_MLP_IMPL=1" > $O/tune.log 2>&1; cat $O/tune.log
NTX_LIB_PATH=$PWD/nerf_texture_b200/lib/libntx_probe.so timeout 400 python tools/field_probe.py 0 1 2 > $O/probe.log 2>&1; cat $O/probe.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json
