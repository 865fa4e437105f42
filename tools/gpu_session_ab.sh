#!/bin/bash
# after the 16-box tensor copies and the structured-form slack: full GPU suite, C2 and C5 (31 qubits, one GPU) bench lines, A/B of the slack
set +e
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/session_ab.log; }
stamp "full gpu suite"
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; stamp "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu5.log
stamp "bench c2"
timeout 300 python bench.py > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err; stamp "bench rc=$?"
stamp "bench c5 n1"
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n1_v2.json 2> gpurun_out/bench_c5_n1_v2.err; stamp "c5 rc=$?"
stamp "A/B slack 100"
BT_FUSE_STRUCT_SLACK_PCT=100 timeout 300 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/bench_n1_slack100.json 2>/dev/null; stamp "c2 slack100 rc=$?"
BT_FUSE_STRUCT_SLACK_PCT=100 timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n1_slack100.json 2>/dev/null; stamp "c5 slack100 rc=$?"
BT_FUSE_STRUCT_SLACK_PCT=200 timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n1_slack200.json 2>/dev/null; stamp "c5 slack200 rc=$?"
for f in bench_n1_v2 bench_c5_n1_v2 bench_n1_slack100 bench_c5_n1_slack100 bench_c5_n1_slack200; do python - <<PY
import json
d=json.loads(open("gpurun_out/$f.json").read())
k=d.get("kernels") or d.get("kernels_rank0")
j=d.get("jit") or d.get("jit_rank0")
print("$f", round(d["value"]), round(d["ms_per_step"],1), k["tile"], j["modules_compiled"], j["specialised_launches"], d["clocks"])
PY
done
cat gpurun_out/session_ab.log
