#!/usr/bin/env python
"""BT_JIT_VERIFY run of C2: every specialised launch of the 28-qubit QFT + layers circuit is cross-checked on the device against the
interpreter kernel; prints launches checked / disagreements.  Usage: BT_JIT_VERIFY=1 python tools/jit_verify.py [N] [layers]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 100
os.environ.setdefault("BT_JIT_VERIFY", "1")
os.environ["BT_TILE_JIT_AFTER"] = "1"
kind = sys.argv[3] if len(sys.argv) > 3 else "qft+layers"
specs = wl.qft(N) if kind == "qft" else (wl.layered(N, layers, 28) if kind == "layers" else wl.qft(N) + wl.layered(N, layers, 28))
arr = bt.pack_gates(wl.to_ops(bt, specs))
s = bt.zero_state(N)
for rep in range(2):
    L.check(s.lib.bt_sv_set_basis(s.h, 5))
    L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    s.sync()
    L.check(s.lib.bt_jit_wait(None))
chk, bad = C.c_uint64(), C.c_uint64()
L.check(s.lib.bt_jit_verify_stats(C.byref(chk), C.byref(bad)))
c, l, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
s.lib.bt_jit_stats(C.byref(c), C.byref(l), C.byref(f), None)
cfg = [C.c_int() for _ in range(4)]
L.check(s.lib.bt_jit_config(*[C.byref(x) for x in cfg]))
print(f"{kind} N={N} gates={len(arr)} nvrtc {cfg[0].value}.{cfg[1].value} code shape {cfg[2].value} BT_JIT_OPT {cfg[3].value}: modules {c.value}, specialised launches {l.value}, cross-checked {chk.value}, disagreements {bad.value}, norm2 {bt.norm2(s):.12f}")
