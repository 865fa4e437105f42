#!/bin/bash
# round 2, 1 GPU: ncu full capture of one layered specialised pass of the final generator (BT_JIT_OPT=7), same launch index as r2_bt_jit_pass_layer_full
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
BT_TILE_JIT_AFTER=1 timeout 140 ncu --set full --clock-control none --import-source on -k regex:bt_jit_pass -s 230 -c 1 -o gpurun_out/r2_bt_jit_pass_layer_opt7_full python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" > gpurun_out/r2_ncu_full3.log 2>&1
tail -2 gpurun_out/r2_ncu_full3.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
