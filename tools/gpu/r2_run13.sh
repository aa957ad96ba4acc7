#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_13; mkdir -p $O
L=$PWD/nerf_texture_b200/lib
timeout 900 python -m pytest tests/test_gpu_frame.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python tools/tune.py "" "NTX_LIB_PATH=$L/libntx_p8.so" "NTX_LIB_PATH=$L/libntx_p10.so" "NTX_LIB_PATH=$L/libntx_p14.so" > $O/tune.log 2>&1; cat $O/tune.log
