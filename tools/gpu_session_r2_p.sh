#!/bin/bash
# round 2, 1 GPU: T = 11 tiles with the pipelined (two-buffer, K tiles per CTA) frame of the specialised pass
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in "BT_TILE_BITS=11 BT_TILE_PIPE=8" "BT_TILE_BITS=11 BT_TILE_PIPE=8 BT_JIT_VARIANT=0"; do env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 $v timeout 300 python tools/jit_verify.py 28 100 2>&1 | tail -1 | cut -c1-220; env BT_JIT_CACHE_DIR= BT_JIT_VERIFY=1 $v timeout 300 python tools/jit_verify.py 28 0 qft 2>&1 | tail -1 | cut -c1-220; done
for v in "BT_TILE_BITS=12" "BT_TILE_BITS=11" "BT_TILE_BITS=11 BT_TILE_PIPE=4" "BT_TILE_BITS=11 BT_TILE_PIPE=8" "BT_TILE_BITS=11 BT_TILE_PIPE=16" "BT_TILE_BITS=11 BT_TILE_PIPE=8 BT_JIT_VARIANT=0" "BT_TILE_BITS=11 BT_TILE_PIPE=8 BT_FUSE_MAX_GATES=30" "BT_TILE_BITS=10 BT_TILE_PIPE=8"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
