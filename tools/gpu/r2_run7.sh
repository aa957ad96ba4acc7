#!/bin/bash
# round-2 GPU session 7: TMEM read microbenchmark, full parity run of the final single-GPU build, tune, probe
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_7; mkdir -p $O
timeout 60 tools/microbench/tmem_ld > $O/tmem_ld.txt 2>&1; cat $O/tmem_ld.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
timeout 300 python tools/tune.py "" > $O/tune.log 2>&1; cat $O/tune.log
