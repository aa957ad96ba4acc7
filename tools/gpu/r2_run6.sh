#!/bin/bash
# round-2 GPU session 6: two tile contexts in the field kernel + suspend-hint mbarrier waits: parity, A/B of the wait variants / ring depth, probe
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_6; mkdir -p $O
L=$PWD/nerf_texture_b200/lib
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 900 python tools/tune.py "" "NTX_LIB_PATH=$L/libntx_nohint.so" "NTX_LIB_PATH=$L/libntx_h500.so" "NTX_LIB_PATH=$L/libntx_st3.so" > $O/tune.log 2>&1; cat $O/tune.log
NTX_LIB_PATH=$L/libntx_probe.so timeout 400 python tools/field_probe.py 0 1 2 > $O/probe.log 2>&1; cat $O/probe.log
