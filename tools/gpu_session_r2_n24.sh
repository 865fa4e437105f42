#!/bin/bash
# round 2, 4-GPU box: the C5 bench line at N = 2 and N = 4 with the final code (N = 1 and 8 were measured separately)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n$n.json 2> gpurun_out/r2_bench_c5_n$n.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_c5_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['remap']['ms_per_step'], d['remap']['GBps_per_rank'])
P
done
