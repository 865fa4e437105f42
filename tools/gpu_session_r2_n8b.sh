#!/bin/bash
# round 2, 8 GPUs, final code: bench line at N = 8 and C4 with 4096 trajectories on 8 ranks
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_c5_n8.json').read().strip().splitlines()[-1])
print(8, d['value'], d['ms_per_step'], d['e2e']['value'], d['remap']['ms_per_step'], d['remap']['GBps_per_rank'])
P
C4_REPS=3 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29553 tools/c4_multi.py > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err
cut -c1-300 gpurun_out/r2_c4_n8.json; grep -o '"parity".*' gpurun_out/r2_c4_n8.json | cut -c1-300
