#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 16 2; do for n in 20 24; do echo "== N=$n variant $v"; env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=$v timeout 300 python tools/jit_bisect.py $n qft 2>/dev/null | head -1 | cut -c1-150; done; done
echo "-- perf BT_JIT_VARIANT=16"; env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=16 timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200
echo "-- verify mode, variants 0 / 2 / 16 on C2 (28 qubits, NVRTC 12.9)"
for v in 0 2 16; do env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=$v BT_JIT_VERIFY=1 timeout 300 python tools/jit_verify.py 28 100 2>&1 | tail -2 | cut -c1-250; done
