#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { echo "== $*"; env BT_JIT_CACHE_DIR= "$@" timeout 300 python tools/jit_locate.py 20 2>/dev/null | grep -E "pass [0-9]+:|first bad" | cut -c1-100 | head -7; }
run BT_JIT_VARIANT=4
for n in 24; do echo "== N=$n variant 0/1/2/4"; for v in 0 1 2 4; do env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=$v timeout 300 python tools/jit_bisect.py $n qft 2>/dev/null | head -1 | cut -c1-150; done; done
echo "== perf (NVRTC 12.9)"
for v in 0 1 2 4; do echo "-- BT_JIT_VARIANT=$v"; env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=$v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
echo "== perf (NVRTC 12.8 from the torch wheel)"
for v in 0 2; do echo "-- BT_JIT_VARIANT=$v"; env BT_JIT_CACHE_DIR= BT_JIT_VARIANT=$v BT_NVRTC_LIB=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_nvrtc/lib/libnvrtc.so.12 timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
