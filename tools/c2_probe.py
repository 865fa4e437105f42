import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")

def run(N, depth, fuse, reps=1):
    specs = wl.c2_qft_layered(N, depth, 28)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    s = bt.zero_state(N)
    lib = s.lib
    L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), fuse)); s.sync()
    ms = C.c_float()
    n0 = s.launch_count()
    L.check(lib.bt_sv_timer_start(s.h))
    for _ in range(reps):
        L.check(lib.bt_sv_set_basis(s.h, 0))
        L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), fuse))
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
    n1 = s.launch_count()
    return s, len(arr), ms.value / reps, (n1 - n0) / reps

N = 16
a, ng, _, _ = run(N, 10, 0)
b, _, _, _ = run(N, 10, 1)
print("N=16 fused vs unfused max diff", np.max(np.abs(a.to_numpy() - b.to_numpy())), "norm", bt.norm2(b))
del a, b
N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for fuse in (0, 1):
    s, ng, ms, nl = run(N, depth, fuse)
    print(f"N={N} depth={depth} fuse={fuse}: gates={ng} launches/step={nl:.0f} ms={ms:.1f} gates/s={ng/ms*1e3:.0f} ms/launch={ms/nl:.3f} norm={bt.norm2(s):.12f}")
    del s
for pers in (0, 1):
    os.environ["BT_TILE_PERSISTENT"] = str(pers)
    for mg in (4, 5, 6, 8, 10, 12):
        os.environ["BT_FUSE_MAX_GATES"] = str(mg)
        s, ng, ms, nl = run(N, depth, 1)
        print(f"  persistent={pers} max_gates={mg}: launches/step={nl:.0f} ms={ms:.1f} gates/s={ng/ms*1e3:.0f} ms/launch={ms/nl:.3f}")
        del s
