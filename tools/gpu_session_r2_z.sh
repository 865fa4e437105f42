#!/bin/bash
# round 2, 2-GPU box: multi-process parity (world 2) and the C5 bench line at N = 2 with the final generator (BT_JIT_OPT=7)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_n2_opt7.json 2> gpurun_out/r2_bench_c5_n2_opt7.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_c5_n2_opt7.json').read().strip().splitlines()[-1])
print(2, round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['remap']['ms_per_step'], d['remap']['GBps_per_rank'])
P
tail -3 gpurun_out/r2_bench_c5_n2_opt7.err
