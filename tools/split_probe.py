import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
N = 28
s = bt.zero_state(N); lib = s.lib
for name, specs in (("qft", wl.qft(N)), ("layered100", wl.layered(N, 100, 28)), ("layered10", wl.layered(N, 10, 28))):
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    for dbg in (0, -1, -2):
        os.environ["BT_TILE_STAGGER_NS"] = str(dbg)
        L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1)); s.sync()
        ms = C.c_float(); n0 = s.launch_count()
        p0, b0 = C.c_uint64(), C.c_uint64(); L.check(lib.bt_fusion_stats(C.byref(p0), C.byref(b0)))
        L.check(lib.bt_sv_timer_start(s.h))
        L.check(lib.bt_sv_set_basis(s.h, 0)); L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
        L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
        p1, b1 = C.c_uint64(), C.c_uint64(); L.check(lib.bt_fusion_stats(C.byref(p1), C.byref(b1)))
        print(f"{name:10s} dbg={dbg:2d}: gates={len(arr)} launches={s.launch_count()-n0} tile passes={p1.value-p0.value} blocks={b1.value-b0.value} ms={ms.value:.1f} ms/pass={ms.value/max(1,(s.launch_count()-n0-1)):.2f}")
