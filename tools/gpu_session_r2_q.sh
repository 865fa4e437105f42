#!/bin/bash
# round 2, 1 GPU: T = 11 tiles with more resident CTAs per SM (register cap of the specialised pass)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in "BT_TILE_BITS=11 BT_JIT_MINB=4" "BT_TILE_BITS=11 BT_JIT_MINB=5" "BT_TILE_BITS=11 BT_JIT_MINB=6" "BT_TILE_BITS=11 BT_JIT_MINB=5 BT_JIT_VARIANT=0" "BT_TILE_BITS=10 BT_JIT_MINB=6"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 BT_TILE_BITS=11 BT_JIT_MINB=5 timeout 300 python tools/jit_verify.py 28 100 2>&1 | tail -1 | cut -c1-220
