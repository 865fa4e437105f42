#!/bin/bash
# round 2, 1 GPU: merged diagonal factors and the opaque coefficient offset: correctness (device cross-check) and C2 step time
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
P8=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_nvrtc/lib/libnvrtc.so.12
echo "== device cross-check of every specialised launch (QFT alone at 24 and 28 qubits; C2)"
for v in "" "BT_JIT_VARIANT=34" "BT_JIT_VARIANT=0" "BT_JIT_VARIANT=32"; do for spec in "24 0 qft" "28 0 qft" "28 100 qft+layers"; do env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 $v timeout 300 python tools/jit_verify.py $spec 2>&1 | tail -1 | cut -c1-220; done; done
echo "== C2 step time"
for v in "BT_JIT_MERGE_DIAG=0" "BT_JIT_MERGE_DIAG=1" "BT_JIT_VARIANT=34" "BT_JIT_VARIANT=32" "BT_JIT_VARIANT=0" "BT_JIT_VARIANT=0 BT_NVRTC_LIB=$P8" "BT_JIT_VARIANT=32 BT_NVRTC_LIB=$P8"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
echo "== ncu: one layered pass of the default build (launch 230 of the specialised kernel)"
BT_TILE_JIT_AFTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bt_jit_pass -s 230 -c 1 -o gpurun_out/r2_bt_jit_pass_layer_full python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" > gpurun_out/r2_ncu_full2.log 2>&1
tail -2 gpurun_out/r2_ncu_full2.log | cut -c1-200
