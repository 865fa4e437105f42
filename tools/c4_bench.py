"""C4 (BASELINE configs[3], per-GPU share): 512 trajectories of the 20-qubit monitored brickwork circuit as one batched state."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
N, T = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 512
specs, M = wl.c4_monitored(N, 20, 20)
ops = wl.to_ops(bt, specs)
U = np.random.Generator(np.random.PCG64(20)).random((T, M))
for rep in range(3):
    st = bt.zero_state(N, T)
    st.sync()
    t0 = time.perf_counter()
    ms = C.c_float()
    L.check(st.lib.bt_sv_timer_start(st.h))
    n0 = st.launch_count()
    _, mids = bt.apply(ops, st, rng=bt.BatchDraws(U), track_measurements=True)
    L.check(st.lib.bt_sv_timer_stop(st.h, C.byref(ms)))
    dt = time.perf_counter() - t0
    nl = st.launch_count() - n0
    ngates = len(ops) - M
    print(f"rep {rep}: {T} trajectories x {N}q, {ngates} gates + {M} measurements: device {ms.value:.1f} ms, host {dt*1e3:.1f} ms, launches {nl}, "
          f"{T/dt:.0f} trajectories/s, {T*len(ops)/dt:.0f} trajectory-ops/s, state {16*T*2**N/2**30:.1f} GiB")
    del st

# trajectory-major execution: chunks small enough to stay in the 126 MB L2 run the whole circuit before the next chunk starts
for chunk in (2, 4, 6, 8, 16, 64):
    t0 = time.perf_counter()
    outs = []
    for c0 in range(0, T, chunk):
        c1 = min(T, c0 + chunk)
        st = bt.zero_state(N, c1 - c0)
        _, mids = bt.apply(ops, st, rng=bt.BatchDraws(U[c0:c1]), track_measurements=True)
        outs.append(np.stack([np.atleast_1d(np.asarray(m)) for m in mids], axis=1))
        del st
    dt = time.perf_counter() - t0
    print(f"chunk {chunk:3d} ({16*chunk*2**N/2**20:.0f} MiB): host {dt*1e3:.1f} ms, {T/dt:.0f} trajectories/s")
