#!/bin/bash
# round 2, 1 GPU: full GPU parity suite, smoke, default bench line (driver's command), reference arm
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_pytest_gpu_a.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu_a.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 6000 gpurun_out/r2_bench_n1_a.json
tail -5 gpurun_out/r2_bench_n1_a.err
