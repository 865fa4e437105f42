#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/linalg_bench.py > gpurun_out/r2_linalg_bench.txt 2>&1; cat gpurun_out/r2_linalg_bench.txt | cut -c1-200
timeout 600 python tools/microbench.py 28 > gpurun_out/r2_microbench_28q.txt 2>&1; tail -45 gpurun_out/r2_microbench_28q.txt | cut -c1-160
