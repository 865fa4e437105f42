#!/bin/bash
# round 2, 1 GPU: JIT / scheduler cross-check at 28 qubits, sampling + wide-load tests, microbench of the touched kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/jit_check.py 28 100 > gpurun_out/r2_jit_check.txt 2>&1
cat gpurun_out/r2_jit_check.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_gates.py tests/test_gpu_programs.py tests/test_gpu_golden.py tests/test_gpu_edge.py -m gpu -q -x --durations=5 > gpurun_out/r2_pytest_gpu_c.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu_c.log
