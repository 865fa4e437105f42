#!/bin/bash
# round 2, final 1-GPU session: whole GPU suite, smoke, the driver's bench command twice (second run = warm on-disk cubin cache), reference arm
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2_pytest_gpu_final.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final_warmcache.json 2> /dev/null
python - <<P
import json
for f in ("gpurun_out/r2_bench_n1_final.json", "gpurun_out/r2_bench_n1_final_warmcache.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 3), "fp64", round(d["roofline"]["fp64"]["frac"], 3),
          "interp ms", round(d["interpreter"]["ms_per_step"], 1), "cold first step s", round(d["cold_first_step_s"], 3), "jit", {k: d["jit"][k] for k in ("modules_compiled", "compile_seconds", "all_modules_ready_s", "disk_cache_hits")}, "clocks", d["clocks"]["sm_mhz"])
P
timeout 300 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
