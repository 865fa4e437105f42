#!/bin/bash
# debug session: pass lists (BT_TILE_DEBUG=1) of C5 at 31 qubits and of C2, to see which passes miss the specialiser
set +e
mkdir -p gpurun_out
BT_TILE_DEBUG=1 timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/dbg_c5.json 2> gpurun_out/dbg_c5.err
grep -c "^\[tile\] T=" gpurun_out/dbg_c5.err
BT_TILE_DEBUG=1 timeout 300 python bench.py --no-cpu --steps 1 --warmup 1 > gpurun_out/dbg_c2.json 2> gpurun_out/dbg_c2.err
grep -c "^\[tile\] T=" gpurun_out/dbg_c2.err
gzip -f gpurun_out/dbg_c5.err gpurun_out/dbg_c2.err
