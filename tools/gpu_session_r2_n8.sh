#!/bin/bash
# round 2, 8 GPUs (charged 8x: keep it short): world-8 multi-process parity, bench line at N=8, C4 with 4096 trajectories on 8 ranks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r2_pytest_multi_n8.log 2>&1
tail -6 gpurun_out/r2_pytest_multi_n8.log
BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err
cut -c1-1800 gpurun_out/r2_bench_c5_n8.json
tail -3 gpurun_out/r2_bench_c5_n8.err
C4_REPS=2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29553 tools/c4_multi.py > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err
cut -c1-1500 gpurun_out/r2_c4_n8.json; tail -3 gpurun_out/r2_c4_n8.err
