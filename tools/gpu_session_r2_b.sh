#!/bin/bash
# round 2, 1 GPU: new linear-algebra tests, fixed VQA test, scheduler sweep on C2, then the whole GPU suite with the new scheduler default
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_vqa.py -m gpu -x -q --durations=8 > gpurun_out/r2_pytest_linalg.log 2>&1
tail -30 gpurun_out/r2_pytest_linalg.log
timeout 600 python tools/sched_sweep.py > gpurun_out/r2_sched_sweep.txt 2>&1
cat gpurun_out/r2_sched_sweep.txt
timeout 900 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_linalg.py --deselect tests/test_gpu_vqa.py > gpurun_out/r2_pytest_gpu_b.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu_b.log
