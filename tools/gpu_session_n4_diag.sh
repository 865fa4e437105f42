#!/bin/bash
# 4-GPU session: multi-process parity (incl. back-to-back steps), remap diagnostics, bench line with per-rank debug output
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r2_pytest_multi_n4.log 2>&1
tail -8 gpurun_out/r2_pytest_multi_n4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29551 tools/remap_diag.py > gpurun_out/r2_remap_diag_n4.log 2>&1
tail -40 gpurun_out/r2_remap_diag_n4.log
