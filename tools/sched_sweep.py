#!/usr/bin/env python
"""C2 step (28-qubit QFT + 100 layers) under different pass-scheduler / tile knobs: ms per step (CUDA events on the handle's stream,
steady state: specialised passes compiled during a warm-up step), kernel launches per step, and agreement of the final state
with the first configuration (<a|b>, <Z_q>).  Usage: python tools/sched_sweep.py [N] [layers]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 100
specs = wl.qft(N) + wl.layered(N, layers, 28)
arr = bt.pack_gates(wl.to_ops(bt, specs))
ng = len(arr)
CONFIGS = [
    ("first fit (round 1)", {"BT_FUSE_SCHED": 0}),
    ("look-ahead", {"BT_FUSE_SCHED": 1}),
    ("look-ahead, cost cap 20", {"BT_FUSE_SCHED": 1, "BT_FUSE_MAX_GATES": 20}),
    ("look-ahead, cost cap 40", {"BT_FUSE_SCHED": 1, "BT_FUSE_MAX_GATES": 40}),
    ("look-ahead, LOWB=4", {"BT_FUSE_SCHED": 1, "BT_TILE_LOWB": 4}),
    ("look-ahead, LOWB=3", {"BT_FUSE_SCHED": 1, "BT_TILE_LOWB": 3}),
    ("look-ahead, LOWB=3, cost cap 40", {"BT_FUSE_SCHED": 1, "BT_TILE_LOWB": 3, "BT_FUSE_MAX_GATES": 40}),
    ("look-ahead, interpreter (BT_TILE_JIT=0)", {"BT_FUSE_SCHED": 1, "BT_TILE_JIT": 0}),
    ("first fit, interpreter (BT_TILE_JIT=0)", {"BT_FUSE_SCHED": 0, "BT_TILE_JIT": 0}),
]
if len(sys.argv) > 3:
    CONFIGS = [c for c in CONFIGS if any(k in c[0] for k in sys.argv[3].split("|"))]
os.environ["BT_TILE_JIT_AFTER"] = "1"
ref = None
a = bt.zero_state(N)
for name, env in CONFIGS:
    for k, v in env.items():
        os.environ[k] = str(v)

    def step():
        L.check(a.lib.bt_sv_set_basis(a.h, 0))
        L.check(a.lib.bt_sv_apply_circuit(a.h, L.ptr(arr), ng, 1))

    t0 = time.perf_counter()
    step(); a.sync()
    L.check(a.lib.bt_jit_wait(None))
    step(); a.sync()
    warm = time.perf_counter() - t0
    l0 = a.launch_count()
    ms = C.c_float()
    reps = 3
    L.check(a.lib.bt_sv_timer_start(a.h))
    for _ in range(reps):
        step()
    L.check(a.lib.bt_sv_timer_stop(a.h, C.byref(ms)))
    launches = (a.launch_count() - l0) // reps
    ez = bt.expect(a, "Z")
    if ref is None:
        ref = a.copy()
        ez0 = ez
        agree = "reference"
    else:
        agree = f"|<a|b>|-1 = {abs(bt.inner(ref, a)) - 1:+.1e}, max|dZ| = {np.max(np.abs(ez - ez0)):.1e}"
    print(f"{name:42s} {ms.value / reps:7.1f} ms/step  {ng / (ms.value / reps) * 1e3:7.0f} gates/s  {launches:4d} launches/step  {ms.value / reps / max(1, launches - 1):.3f} ms/launch  warm-up {warm:.1f} s  {agree}", flush=True)
    for k in env:
        os.environ.pop(k)
