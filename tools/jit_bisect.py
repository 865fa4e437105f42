#!/usr/bin/env python
"""Find the first gate of a circuit at which the specialised (NVRTC) passes and the interpreter disagree: binary search over the
prefix length, states compared through |<a|b>| and max |amplitude difference| (N <= 24).  Usage: python tools/jit_bisect.py [N] [qft|layers]"""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = sys.argv[2] if len(sys.argv) > 2 else "qft"
specs = wl.qft(N) if kind == "qft" else wl.layered(N, 10, 28)
arr_all = bt.pack_gates(wl.to_ops(bt, specs))


def run(L_, jit):
    os.environ["BT_TILE_JIT"] = "2" if jit else "0"
    s = bt.zero_state(N)
    L.check(s.lib.bt_sv_set_basis(s.h, 5))
    arr = arr_all[:L_]
    L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    s.sync()
    return s


def diff(L_):
    a, b = run(L_, False), run(L_, True)
    ov = bt.inner(a, b)
    d = abs(abs(ov) - 1) + abs(np.angle(ov))
    if N <= 24:
        d = float(np.max(np.abs(a.to_numpy() - b.to_numpy())))
    return d


full = diff(len(arr_all))
print(f"{kind} N={N} gates={len(arr_all)}: interpreter vs specialised passes differ by {full:.3e}")
if full > 1e-10:
    lo, hi = 0, len(arr_all)  # diff(lo) small, diff(hi) large
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if diff(mid) > 1e-10:
            hi = mid
        else:
            lo = mid
    print(f"first bad prefix length {hi}: gate #{hi - 1} = {specs[hi - 1]}  (diff {diff(hi):.3e}; with one gate less {diff(lo):.3e})")
    for k in range(max(0, hi - 6), hi):
        print("   ", k, specs[k])
    os.environ["BT_TILE_DEBUG"] = "1"
    run(hi, True)
    os.environ.pop("BT_TILE_DEBUG")
c, l, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
L.load().bt_jit_stats(C.byref(c), C.byref(l), C.byref(f), None)
print("jit: compiled", c.value, "specialised launches", l.value, "fell back", f.value)
