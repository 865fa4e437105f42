#!/bin/bash
# round 2, 1 GPU: whole GPU suite, driver-style bench line, C5 at 31 qubits on one GPU, ncu launch list + one full capture of the fused pass
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2_pytest_gpu_d.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_d.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err
cut -c1-1200 gpurun_out/r2_bench_n1_d.json; tail -3 gpurun_out/r2_bench_n1_d.err
timeout 600 python bench.py --gpus 1 --workload c5 --steps 8 --warmup 3 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
cut -c1-800 gpurun_out/r2_bench_c5_n1.json
# ncu: launch list of the bench command (short), then one full capture of a specialised pass in steady state
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_c2.csv python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-300
BT_TILE_JIT_AFTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bt_jit_pass -s 200 -c 1 -o gpurun_out/r2_bt_jit_pass_full python tools/sched_sweep.py 28 100 "look-ahead, LOWB=3, cost cap 40" > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep 2>/dev/null
