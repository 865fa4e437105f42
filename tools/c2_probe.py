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
    p0, b0 = C.c_uint64(), C.c_uint64(); L.check(lib.bt_fusion_stats(C.byref(p0), C.byref(b0)))
    L.check(lib.bt_sv_timer_start(s.h))
    for _ in range(reps):
        L.check(lib.bt_sv_set_basis(s.h, 0))
        L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), fuse))
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
    n1 = s.launch_count()
    p1, b1 = C.c_uint64(), C.c_uint64(); L.check(lib.bt_fusion_stats(C.byref(p1), C.byref(b1)))
    return s, len(arr), ms.value / reps, (n1 - n0) / reps, (p1.value - p0.value) / reps, (b1.value - b0.value) / reps

N = 16
a = run(N, 10, 0)[0]
for cl in (0, 1):
    os.environ["BT_TILE_CLUSTERS"] = str(cl)
    b = run(N, 10, 1)[0]
    print(f"N=16 clusters={cl} fused vs unfused max diff", np.max(np.abs(a.to_numpy() - b.to_numpy())), "norm", bt.norm2(b))
    del b
del a
N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for (cl, tb, db, tpc) in ((1, 12, 0, 1), (0, 12, 0, 1), (1, 11, 1, 8), (1, 11, 0, 1)):
    os.environ["BT_TILE_CLUSTERS"] = str(cl)
    os.environ["BT_TILE_BITS"] = str(tb); os.environ["BT_TILE_DB"] = str(db); os.environ["BT_TILE_PER_CTA"] = str(tpc)
    for mg in (6, 8, 10, 12):
        os.environ["BT_FUSE_MAX_GATES"] = str(mg)
        s, ng, ms, nl, npass, nblk = run(N, depth, 1)
        print(f"  clusters={cl} T={tb} db={db} tiles/cta={tpc} max_gates={mg}: tile passes={npass:.0f} blocks={nblk:.0f} ms={ms:.1f} gates/s={ng/ms*1e3:.0f} ms/launch={ms/nl:.3f}")
        del s
