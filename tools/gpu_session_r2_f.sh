#!/bin/bash
# round 2, 1 GPU: locate the specialised pass that disagrees with the interpreter on the QFT (N = 20), under a few knob variations
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/jit_locate.py 20 > gpurun_out/r2_jit_locate_20.txt 2> gpurun_out/r2_jit_locate_20.err
grep -v "^//\|^ *{\|^ *if\|^ *}" gpurun_out/r2_jit_locate_20.txt | head -40 | cut -c1-300
BT_TILE_LOWB=5 timeout 300 python tools/jit_locate.py 20 2>/dev/null | head -12 | cut -c1-300
BT_FUSE_SCHED=0 timeout 300 python tools/jit_locate.py 20 2>/dev/null | head -12 | cut -c1-300
BT_JIT_CACHE_DIR= BT_JIT_EXTRA_OPTS="--fmad=false" timeout 300 python tools/jit_locate.py 20 2>/dev/null | head -12 | cut -c1-300
BT_JIT_CACHE_DIR= BT_JIT_EXTRA_OPTS="-Xptxas=-O0" timeout 300 python tools/jit_locate.py 20 2>/dev/null | head -12 | cut -c1-300
grep -c . gpurun_out/r2_jit_locate_20.err
