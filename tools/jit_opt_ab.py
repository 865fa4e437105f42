#!/usr/bin/env python
"""A/B of the specialiser's arithmetic reshaping (BT_JIT_OPT=0 against the default) in one process: ms per step (CUDA events on the
handle's stream, steady state), launches per step, and agreement of the final states (<a|b>, max |d<Z_q>|).
Usage: python tools/jit_opt_ab.py [c2|c5] [N]"""
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
work = sys.argv[1] if len(sys.argv) > 1 else "c2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else (28 if work == "c2" else 30)
specs = wl.c2_qft_layered(N) if work == "c2" else wl.c5_random(N)
arr = bt.pack_gates(wl.to_ops(bt, specs))
ng = len(arr)
os.environ["BT_TILE_JIT_AFTER"] = "1"
cfg = [C.c_int() for _ in range(4)]
L.check(L.load().bt_jit_config(*[C.byref(x) for x in cfg]))
print(f"{work} N={N} gates={ng} nvrtc {cfg[0].value}.{cfg[1].value} code shape {cfg[2].value}", flush=True)
a = bt.zero_state(N)
ref = None
CONFIGS = [{"BT_JIT_OPT": 0}, {"BT_JIT_OPT": 7}, {"BT_JIT_OPT": 0}, {"BT_JIT_OPT": 7}]
for env in CONFIGS:
    for k in ("BT_JIT_OPT", "BT_FUSE_MAX_GATES"):
        os.environ.pop(k, None)
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
    ms = C.c_float()
    reps = 5
    l0 = a.launch_count()
    L.check(a.lib.bt_sv_timer_start(a.h))
    for _ in range(reps):
        step()
    L.check(a.lib.bt_sv_timer_stop(a.h, C.byref(ms)))
    launches = (a.launch_count() - l0) // reps
    ez = bt.expect(a, "Z")
    if ref is None:
        ref, ez0, agree = a.copy(), ez, "reference"
    else:
        agree = f"|<a|b>|-1 = {abs(bt.inner(ref, a)) - 1:+.1e}, max|dZ| = {np.max(np.abs(ez - ez0)):.1e}"
    print(f"{' '.join(f'{k}={v}' for k, v in env.items()):40s} {ms.value / reps:8.2f} ms/step  {ng / (ms.value / reps) * 1e3:8.0f} gates/s  {launches} launches/step  warm-up {warm:.1f} s  {agree}", flush=True)
