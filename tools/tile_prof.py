import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
specs = wl.layered(N, 6, 28)
arr = bt.pack_gates(wl.to_ops(bt, specs))
s = bt.plus_state(N)
L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1)); s.sync()
print("launches", s.launch_count())
