#!/bin/bash
# round 2, last 1-GPU session: device cross-check of the reshaped specialised passes (BT_JIT_VERIFY), A/B against BT_JIT_OPT=0,
# whole GPU suite, smoke, the driver's bench command
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
BT_JIT_VERIFY=1 timeout 200 python tools/jit_verify.py 28 100 > gpurun_out/r2w_jit_verify.txt 2>&1; tail -2 gpurun_out/r2w_jit_verify.txt
timeout 200 python tools/jit_opt_ab.py c2 > gpurun_out/r2w_jit_opt_ab.txt 2>&1; cat gpurun_out/r2w_jit_opt_ab.txt | tail -5
timeout 200 python tools/jit_opt_ab.py c5 30 >> gpurun_out/r2w_jit_opt_ab.txt 2>&1; tail -5 gpurun_out/r2w_jit_opt_ab.txt
timeout 600 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2w_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2w_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "fp64", round(d["roofline"]["fp64"]["frac"], 3),
      "jit", {k: d["jit"].get(k) for k in ("nvrtc", "code_shape_variant", "arith_opt", "modules_compiled", "fell_back")}, "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
P
