#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_path.py tests/test_gpu_golden.py tests/test_gpu_analytic.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python tools/linalg_bench.py > gpurun_out/r2_linalg_bench.txt 2>&1; cat gpurun_out/r2_linalg_bench.txt | cut -c1-200
timeout 600 python tools/microbench.py 28 > gpurun_out/r2_microbench_28q.txt 2>&1; grep -E "sample|measure_z" gpurun_out/r2_microbench_28q.txt | cut -c1-160
