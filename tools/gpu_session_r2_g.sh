#!/bin/bash
# round 2, 1 GPU: which change of the generated code / compile options makes the specialised QFT passes agree with the interpreter
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { echo "== $*"; env BT_JIT_CACHE_DIR= "$@" timeout 300 python tools/jit_locate.py 20 2>/dev/null | grep -E "pass [0-9]+:|first bad" | cut -c1-200 | head -8; }
run BT_JIT_VARIANT=1
run BT_JIT_VARIANT=2
run BT_JIT_VARIANT=3
run BT_JIT_EXTRA_OPTS=-Xptxas=-O1
run BT_JIT_EXTRA_OPTS=-Xptxas=-O2
run BT_NVRTC_LIB=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_nvrtc/lib/libnvrtc.so.12
run BT_NVRTC_LIB=/usr/local/cuda/lib64/libnvrtc.so.12
echo "== perf"
for v in "BT_JIT_VARIANT=0" "BT_JIT_VARIANT=2" "BT_JIT_VARIANT=3" "BT_JIT_EXTRA_OPTS=-Xptxas=-O1" "BT_JIT_EXTRA_OPTS=-Xptxas=-O2"; do echo "-- $v"; env BT_JIT_CACHE_DIR= $v timeout 300 python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" 2>&1 | tail -1 | cut -c1-200; done
