#!/bin/bash
# round 2, 4 GPUs, after the remap fixes: multi-process parity (back-to-back steps), short remap diagnostics, bench line, C4 on 4 ranks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r2_pytest_multi_n4.log 2>&1
tail -6 gpurun_out/r2_pytest_multi_n4.log
DIAG_QUICK=1 DIAG_STEPS=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29551 tools/remap_diag.py > gpurun_out/r2_remap_diag_n4_after.log 2>&1
grep -E "^\[pull\]|^\[step\]" gpurun_out/r2_remap_diag_n4_after.log | cut -c1-420
BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n4.json 2> gpurun_out/r2_bench_c5_n4.err
cut -c1-2500 gpurun_out/r2_bench_c5_n4.json
tail -3 gpurun_out/r2_bench_c5_n4.err
C4_REPS=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29553 tools/c4_multi.py > gpurun_out/r2_c4_n4.json 2> gpurun_out/r2_c4_n4.err
cut -c1-2000 gpurun_out/r2_c4_n4.json; tail -3 gpurun_out/r2_c4_n4.err
