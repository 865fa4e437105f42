#!/bin/bash
# 2-GPU validation of the device-side remap synchronisation (flags in peer memory): multi-process parity check, then the sharded
# bench with per-rank step times.  Short in-kernel timeout so that a protocol error traps instead of hanging the box.
set +e
mkdir -p gpurun_out
export BT_REMAP_TIMEOUT_S=15
timeout 70 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi_flags.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/pytest_multi_flags.log
BENCH_DEBUG=1 timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_c5_n2_flags.json 2> gpurun_out/bench_c5_n2_flags.err; echo "bench rc=$?"
grep "bench rank" gpurun_out/bench_rank0.err | cut -c1-260; tail -3 gpurun_out/bench_c5_n2_flags.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c5_n2_flags.json').read()); print(round(d['value']), d['ms_per_step'], d['e2e']['seconds_per_step'], d['remap'], d['checksum'])"
