#!/bin/bash
# round 2, 1 GPU: ncu launch list of the driver's bench command with the final generator (shares of the step per kernel)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_c2_opt7.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-blocks --no-cpu > gpurun_out/r2_ncu_bench_opt7.log 2>&1
tail -c 300 gpurun_out/r2_ncu_bench_opt7.log; wc -l gpurun_out/r2_launches_bench_c2_opt7.csv
