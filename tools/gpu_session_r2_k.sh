#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in "BT_JIT_VARIANT=2" "BT_JIT_VARIANT=66" "BT_JIT_VARIANT=66 BT_JIT_PREFETCH_DIST=148" "BT_JIT_VARIANT=66 BT_JIT_PREFETCH_DIST=296" "BT_JIT_VARIANT=66 BT_JIT_PREFETCH_DIST=888" "BT_JIT_VARIANT=98"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 BT_JIT_VARIANT=66 timeout 300 python tools/jit_verify.py 28 100 2>&1 | tail -1 | cut -c1-220
