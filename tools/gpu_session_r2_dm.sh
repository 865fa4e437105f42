#!/bin/bash
# round 2, 1 GPU, last seconds of the budget: C3 block (fused superoperators, single DM kernels) after the plain k_dense form went back to
# row-by-row stores (machine code identical to the round-1 kernels), then the DM parity tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 16 python - > gpurun_out/r2_dm14_after_store_fix.json 2> gpurun_out/r2_dm14_after_store_fix.err <<P
import sys, json
sys.path.insert(0, ".")
import __graft_entry__ as ge, bench
bt = ge.load_package()
from importlib import import_module
wl = import_module(ge.PKG_NAME + ".workloads")
d = bench.dm14_block(bt, bt._lib, wl, 6549.8)
print(json.dumps({"fused": d["fused_superoperators"], "op_by_op_ms": d["op_by_op"]["ms"], "kernels": {k: round(v["ms"], 4) for k, v in d["kernels"].items()}}))
P
cat gpurun_out/r2_dm14_after_store_fix.json | cut -c1-900; tail -2 gpurun_out/r2_dm14_after_store_fix.err
timeout 14 python -m pytest tests/test_gpu_path.py -m gpu -q -x -k "dm_unitaries_match_oracle_and_sv or dm_controlled_and_nonadjacent or dm_channels_match_oracle or dm_fused_op_list" 2>&1 | tail -2
