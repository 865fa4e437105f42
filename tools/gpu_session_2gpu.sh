#!/bin/bash
# 2-GPU session: sharded bench line (weak scaling workload C5 at 31 local qubits per GPU) as the driver launches it, the
# multi-process GPU test, and the variational tests / timing after the work-state change.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/session_2gpu.log; }
stamp "bench n2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; stamp "bench n2 rc=$?"
cat gpurun_out/bench_c5_n2.json | cut -c1-1500; tail -5 gpurun_out/bench_c5_n2.err
stamp "reference arm under torchrun"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --cpu-budget 8 > gpurun_out/bench_reference_n2.json 2> gpurun_out/bench_reference_n2.err; stamp "ref n2 rc=$?"
cut -c1-300 gpurun_out/bench_reference_n2.json
stamp "multi tests"
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_vqa.py -q > gpurun_out/pytest_multi.log 2>&1; stamp "pytest rc=$?"
tail -5 gpurun_out/pytest_multi.log
stamp "vqa bench"
timeout 200 python tools/vqa_bench.py 28 > gpurun_out/vqa_bench.txt 2> gpurun_out/vqa_bench.err; stamp "vqa bench rc=$?"
cat gpurun_out/vqa_bench.txt
cat gpurun_out/session_2gpu.log
