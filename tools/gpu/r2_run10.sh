#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_10; mkdir -p $O
timeout 600 python tools/schedule_sweep.py 8 2 1 > $O/schedule_sweep.log 2>&1; cat $O/schedule_sweep.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:march_rays_compact --launch-skip 1 --launch-count 1 -o $O/prof_march_shard python tools/profile_shard.py > $O/ncu_march.log 2>&1
