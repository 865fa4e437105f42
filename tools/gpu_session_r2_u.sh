#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_edge.py tests/test_gpu_analytic.py -m gpu -q -x 2>&1 | tail -3
C4_TRAJECTORIES=512 C4_REPS=3 timeout 300 python tools/c4_multi.py 2>&1 | tail -1 | cut -c1-700
