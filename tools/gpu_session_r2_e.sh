#!/bin/bash
# round 2, 1 GPU: locate the QFT disagreement between specialised passes and the interpreter, then the tests touched since session B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for n in 16 20 24 28; do timeout 300 python tools/jit_bisect.py $n qft > gpurun_out/r2_jit_bisect_$n.txt 2>&1; grep -v "^\[tile\]" gpurun_out/r2_jit_bisect_$n.txt | cut -c1-300; grep "^\[tile\]" gpurun_out/r2_jit_bisect_$n.txt | tail -4 | cut -c1-400; done
timeout 900 python -m pytest tests/test_gpu_programs.py tests/test_gpu_path.py tests/test_gpu_linalg.py tests/test_gpu_golden.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -m gpu -q --durations=5 > gpurun_out/r2_pytest_gpu_e.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_e.log | cut -c1-300
