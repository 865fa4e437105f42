#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_gates.py tests/test_gpu_fullsize.py -m gpu -q -x -k "dm or rho or density or 3q or kraus or c3" 2>&1 | tail -3
timeout 600 python tools/microbench.py 28 > gpurun_out/r2_microbench_28q.txt 2>&1; grep -E "^DM|3q|CCX" gpurun_out/r2_microbench_28q.txt | cut -c1-160
