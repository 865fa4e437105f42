#!/bin/bash
# final evidence of the round on one B200: bench line, reference arm, ncu launch list of the bench command, smoke()
set +e
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/session_final.log; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stamp "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; stamp "bench rc=$?"
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget 12 > gpurun_out/bench_reference_final.json 2>/dev/null; stamp "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_final.log 2>&1; stamp "ncu list rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_final.json').read()); print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['kernels'], d['clocks'])"
