#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for spec in "24 0 qft" "28 0 qft" "28 100 qft+layers"; do env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 timeout 300 python tools/jit_verify.py $spec 2>&1 | tail -1 | cut -c1-220; done
for v in "BT_JIT_SHEAR=0" "BT_JIT_SHEAR=1" "BT_JIT_SHEAR=1 BT_JIT_VARIANT=0"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
timeout 300 python -m pytest tests/test_gpu_programs.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
