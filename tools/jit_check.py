#!/usr/bin/env python
"""Cross-check of the fused path at full size: the same circuit under {first fit, look-ahead} x {interpreter, specialised passes};
every pair of final states must agree (|<a|b>| = 1 to rounding).  Usage: python tools/jit_check.py [N] [layers]"""
import ctypes as C
import os
import sys

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
parts = {"qft": wl.qft(N), "layers": wl.layered(N, layers, 28), "qft+layers": wl.qft(N) + wl.layered(N, layers, 28)}
os.environ["BT_TILE_JIT_AFTER"] = "1"
CONF = [("ff-interp", {"BT_FUSE_SCHED": 0, "BT_TILE_JIT": 0}), ("ff-jit", {"BT_FUSE_SCHED": 0}), ("la-interp", {"BT_FUSE_SCHED": 1, "BT_TILE_JIT": 0}), ("la-jit", {"BT_FUSE_SCHED": 1}),
        ("unfused", None)]
for pname, specs in parts.items():
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    states = {}
    for name, env in CONF:
        if name == "unfused" and N > 26 and pname != "qft":
            continue
        for k, v in (env or {}).items():
            os.environ[k] = str(v)
        s = bt.zero_state(N)
        for rep in range(3 if env and "BT_TILE_JIT" not in env else 1):  # specialised: interpreter + compile, wait, then twice on the modules
            L.check(s.lib.bt_sv_set_basis(s.h, 5))
            L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 0 if env is None else 1))
            s.sync()
            L.check(s.lib.bt_jit_wait(None))
        states[name] = s
        for k in (env or {}):
            os.environ.pop(k)
    names = list(states)
    ref = states[names[0]]
    line = [f"{pname:11s} N={N} gates={len(arr)} norm2[{names[0]}]={bt.norm2(ref):.12f}"]
    for nm in names[1:]:
        ov = bt.inner(ref, states[nm])
        line.append(f"|<{names[0]}|{nm}>|-1={abs(ov) - 1:+.2e} arg={np.angle(ov):+.1e}")
    print("  ".join(line), flush=True)
    del states, ref
c, l, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
L.load().bt_jit_stats(C.byref(c), C.byref(l), C.byref(f), None)
print("jit: compiled", c.value, "specialised launches", l.value, "fell back", f.value)
