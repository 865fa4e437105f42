#!/usr/bin/env python
"""Locate a disagreement between a specialised (NVRTC) pass and the interpreter: run the circuit with exactly ONE pass specialised
(bt_jit_debug_only) for every pass in turn, compare with the all-interpreter run, and for the first bad pass print which index bits
the wrong amplitudes have in common, the ratio wrong/right, and the generated source.  Usage: python tools/jit_locate.py [N] [prefix]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
lib = L.load()
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
specs = wl.qft(N)
if len(sys.argv) > 2:
    specs = specs[: int(sys.argv[2])]
arr = bt.pack_gates(wl.to_ops(bt, specs))


def run(only, dump=0):
    os.environ["BT_TILE_JIT"] = "0" if only is None else "2"
    lib.bt_jit_debug_only(-1 if only is None else only, dump)
    s = bt.zero_state(N)
    L.check(s.lib.bt_sv_set_basis(s.h, 5))
    L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    v = s.to_numpy()
    lib.bt_jit_debug_only(-1, 0)
    return v


ref = run(None)
print(f"N={N} gates={len(arr)} extra NVRTC options: {os.environ.get('BT_JIT_EXTRA_OPTS', '')!r} knobs: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("BT_") and k not in ("BT_TILE_JIT",)))
bad = None
for k in range(32):
    v = run(k)
    d = float(np.max(np.abs(v - ref)))
    c, l, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.bt_jit_stats(C.byref(c), C.byref(l), C.byref(f), None)
    print(f"  pass {k}: max |specialised - interpreter| = {d:.3e}   (modules {c.value}, specialised launches so far {l.value})")
    if d > 1e-10 and bad is None:
        bad = (k, v)
    if k > 0 and d == 0.0 and l.value == prev_l:
        break
    prev_l = l.value
if bad:
    k, v = bad
    idx = np.nonzero(np.abs(v - ref) > 1e-12)[0]
    print(f"first bad pass {k}: {len(idx)} of {len(v)} amplitudes differ")
    allone = np.bitwise_and.reduce(idx)
    allzero = np.bitwise_and.reduce(~idx) & ((1 << N) - 1)
    print(f"  index bits set in every wrong amplitude: {[b for b in range(N) if (allone >> b) & 1]}, clear in every one: {[b for b in range(N) if (allzero >> b) & 1]}")
    r = v[idx[:8]] / ref[idx[:8]]
    print("  wrong / right for the first few:", " ".join(f"{abs(z):.6f}*exp({np.angle(z):+.6f}i)" for z in r))
    print("  their indices:", [bin(int(i)) for i in idx[:8]])
    os.environ["BT_TILE_DEBUG"] = "1"
    run(k, dump=1)
