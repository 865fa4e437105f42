#!/bin/bash
# 4-GPU session after the remap fixes: parity, short diagnostics, bench line (per-rank debug output)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r2_pytest_multi_n4b.log 2>&1
tail -6 gpurun_out/r2_pytest_multi_n4b.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29551 tools/remap_diag.py > gpurun_out/r2_remap_diag_n4b.log 2>&1
grep -E "^\[pull\]|^\[step\]" gpurun_out/r2_remap_diag_n4b.log | cut -c1-420
BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/r2_bench_c5_n4.json 2> gpurun_out/r2_bench_c5_n4.err
cat gpurun_out/r2_bench_c5_n4.json | cut -c1-1500
for r in 0 1 2 3; do tail -3 gpurun_out/bench_rank$r.err | cut -c1-600; done
