#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_8; mkdir -p $O
L=$PWD/nerf_texture_b200/lib
timeout 60 tools/microbench/mlp_chain > $O/mlp_chain.txt 2>&1; cat $O/mlp_chain.txt
timeout 600 python tools/tune.py "" "NTX_LIB_PATH=$L/libntx_nohint.so" > $O/tune.log 2>&1; cat $O/tune.log
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_grid.py tests/test_gpu_frame.py tests/test_gpu_field.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python tools/bench_cfg5.py --backend ntx > $O/cfg5_ntx.log 2>&1; tail -1 $O/cfg5_ntx.log | cut -c1-700
