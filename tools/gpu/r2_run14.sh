#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_14; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_frame.py tests/test_gpu_density.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python tools/tune.py "" > $O/tune.log 2>&1; cat $O/tune.log
NTX_LIB_PATH=$PWD/nerf_texture_b200/lib/libntx_probe.so timeout 400 python tools/field_probe.py 0 1 > $O/probe.log 2>&1; cat $O/probe.log
