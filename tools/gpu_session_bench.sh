#!/bin/bash
# One B200 session of the round: GPU parity suite, the bench line, the ncu launch list of the bench command, one full ncu
# capture of the dominant kernel, the reference arm and a small knob sweep.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/session_bench.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
stamp "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; stamp "pytest rc=$?"
stamp "bench n1"
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; stamp "bench rc=$?"
stamp "ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; stamp "ncu list rc=$?"
stamp "ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bt_jit_pass -s 60 -c 1 -f -o gpurun_out/jit_pass_full_v2 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; stamp "ncu full rc=$?"
ncu -i gpurun_out/jit_pass_full_v2.ncu-rep --page raw --csv > gpurun_out/jit_pass_full_v2.csv 2>/dev/null
stamp "reference arm"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget 12 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; stamp "ref rc=$?"
stamp "sweeps"
BT_FUSE_MAX_GATES=44 timeout 200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_cap44.json 2>/dev/null; stamp "cap44 rc=$?"
BT_TILE_LOWB=3 timeout 200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_lowb3.json 2>/dev/null; stamp "lowb3 rc=$?"
BT_TILE_LOWB=3 BT_FUSE_MAX_GATES=44 timeout 200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_lowb3_cap44.json 2>/dev/null; stamp "lowb3cap44 rc=$?"
if [ "$1" = "configs" ]; then timeout 300 python tools/config_runs.py > gpurun_out/configs.txt 2>&1; stamp "configs rc=$?"; fi
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/session_bench.log
