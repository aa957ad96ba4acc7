#!/bin/bash
# round-2 GPU session 11: final single-GPU build: smoke, full parity suite, full bench line, ncu captures for profiles/
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_11; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 300 $O/bench_reference_arm.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_frame.csv python tools/profile_frame.py frame > $O/ncu_frame.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ngp_field --launch-skip 1 --launch-count 1 -o $O/prof_field_frame python tools/profile_frame.py frame > $O/ncu_field.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:march_rays_compact --launch-skip 0 --launch-count 2 -o $O/prof_march_frame python tools/profile_frame.py frame > $O/ncu_march.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:mlp_pipe|grid_fwd_pair" -c 2 -o $O/prof_cfg2 python tools/profile_frame.py cfg2 > $O/ncu_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:wgrad_tc|grid_bwd|dgrad" -s 12 -c 6 -o $O/prof_cfg5 python tools/bench_cfg5.py --backend ntx --iters 2 > $O/ncu_cfg5.log 2>&1
NTX_LIB_PATH=$PWD/nerf_texture_b200/lib/libntx_probe.so timeout 400 python tools/field_probe.py 0 1 2 > $O/probe.log 2>&1; cat $O/probe.log
