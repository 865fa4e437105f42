#!/bin/bash
# Second B200 session of the round: the variational callers (Pauli-sum kernels) -- parity tests, memcheck on the small cases,
# full-size timing -- and the T = 11 tile variants of the specialised fused pass.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/session_vqa.log; }
stamp "pytest vqa"
timeout 300 python -m pytest tests/test_gpu_vqa.py -q > gpurun_out/pytest_vqa.log 2>&1; stamp "pytest vqa rc=$?"
tail -25 gpurun_out/pytest_vqa.log
stamp "vqa bench"
timeout 300 python tools/vqa_bench.py 28 > gpurun_out/vqa_bench.txt 2> gpurun_out/vqa_bench.err; stamp "vqa bench rc=$?"
cat gpurun_out/vqa_bench.txt; tail -3 gpurun_out/vqa_bench.err
stamp "memcheck"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vqa.py -q -k "chain_hamiltonians and (3 or 12) or rejects" > gpurun_out/memcheck_vqa.log 2>&1; stamp "memcheck rc=$?"
tail -6 gpurun_out/memcheck_vqa.log
stamp "T=11 variants"
BT_TILE_BITS=11 timeout 200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_t11.json 2>/dev/null; stamp "t11 rc=$?"
BT_TILE_BITS=11 BT_TILE_LOWB=4 timeout 200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_t11_lowb4.json 2>/dev/null; stamp "t11 lowb4 rc=$?"
stamp "full gpu suite"
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; stamp "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu2.log
cat gpurun_out/session_vqa.log
