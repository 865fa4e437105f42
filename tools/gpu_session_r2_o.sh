#!/bin/bash
# round 2, 1 GPU: the wide (256 threads, no group loop) form of the specialised pass: device cross-check on QFT / C2 and C2 step time
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 128 130; do for spec in "24 0 qft" "28 0 qft" "28 100 qft+layers"; do env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 BT_JIT_VARIANT=$v timeout 300 python tools/jit_verify.py $spec 2>&1 | tail -1 | cut -c1-220; done; done
for n in 20 24; do env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=128 timeout 300 python tools/jit_bisect.py $n qft 2>/dev/null | head -1 | cut -c1-150; done
for v in "BT_JIT_VARIANT=2" "BT_JIT_VARIANT=128" "BT_JIT_VARIANT=130" "BT_JIT_VARIANT=160" "BT_JIT_VARIANT=0"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
