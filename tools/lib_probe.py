import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
N, depth = 28, 100
specs = wl.c2_qft_layered(N, depth, 28)
arr = bt.pack_gates(wl.to_ops(bt, specs))
s = bt.zero_state(N); lib = s.lib
cfgs = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("PROBE_CFGS", "12,0,8;12,0,10").split(";")]
for (tb, db, mg) in cfgs:
    os.environ["BT_TILE_BITS"] = str(tb); os.environ["BT_FUSE_MAX_GATES"] = str(mg)
    for _ in range(2):  # the pass specialiser compiles a structure the second time it sees it
        L.check(lib.bt_sv_set_basis(s.h, 0)); L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1)); s.sync()
    ms = C.c_float(); n0 = s.launch_count()
    L.check(lib.bt_sv_timer_start(s.h))
    L.check(lib.bt_sv_set_basis(s.h, 0)); L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
    jc, jl, jf, jt = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double(); lib.bt_jit_stats(C.byref(jc), C.byref(jl), C.byref(jf), C.byref(jt))
    print(f"{os.environ.get('BLUETANGLE_CUDA_LIB','default')[-24:]} T={tb} db={db} mg={mg}: launches={s.launch_count()-n0} ms={ms.value:.1f} gates/s={len(arr)/ms.value*1e3:.0f} norm={bt.norm2(s):.9f} jit: {jc.value} modules, {jl.value} launches, {jf.value} failed, {jt.value:.1f} s compiling")
