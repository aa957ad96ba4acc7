#!/bin/bash
# round-2 GPU session 2: tcgen05 wgrad + privatised grid backward parity, reference-files test, cfg5 / compat numbers
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
for b in ntx ref; do timeout 300 python tools/bench_cfg5.py --backend $b --dump $O/cfg5_$b.npz > $O/cfg5_$b.log 2>&1; tail -2 $O/cfg5_$b.log; done
for b in ntx ref; do timeout 600 python tools/run_reference_files.py --backend $b --size 1024 --time 5 > $O/compat_$b.log 2>&1; tail -2 $O/compat_$b.log; done
