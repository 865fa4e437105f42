"""BASELINE.json configs C1, C3, C4 at full size on one GPU: timings + invariants (C2 and C5 are bench.py lines)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
from oracle import bt_oracle as O

def wall(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))

# ---- C1: 12-qubit brickwork H/CNOT/RZ depth 20 + 4096 shots ------------------------------------------------------
specs = wl.c1_brickwork(12, 20, 12)
ops = wl.to_ops(bt, specs); arr = bt.pack_gates(ops)
us = np.random.Generator(np.random.PCG64(12)).random(4096)
def c1():
    s = bt.zero_state(12)
    L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    return bt.sample(s, 4096, uniforms=us)
t_gpu = wall(c1, 20)
oo = wl.to_ops(O, specs)
t0 = time.perf_counter(); ref = O.apply_ops(O.zero_state(12), oo); smp = O.sample(ref, us); t_ref = time.perf_counter() - t0
same = bool(np.array_equal(c1(), smp))
print(f"C1 12q depth 20 ({len(arr)} gates) + 4096 shots: device path {t_gpu*1e3:.2f} ms end to end ({len(arr)/t_gpu:.0f} gates/s); "
      f"reference-algorithm restatement (kron chain + SpMV, 1 thread) {t_ref*1e3:.0f} ms ({len(arr)/t_ref:.0f} gates/s); samples identical: {same}")

# ---- C3: 14-qubit density matrix, depolarizing + amplitude damping after every gate ------------------------------
N = 14
ops3 = []
for e in wl.c3_noisy_dm(N, 20, 14):
    if e[0] == "gate":
        name, q, t, c = e[1]; ops3.append(bt.Op(name, q, t, control=c))
    else:
        _, model, p, q, t = e; ops3.append(bt.OpQC(model, p, q, t))
ngate = sum(1 for o in ops3 if isinstance(o, bt.Op)); nch = len(ops3) - ngate
for fused in (True, False):
    rho = bt.CuRho(N); rho.sync()
    ms = C.c_float(); n0 = rho.launch_count()
    L.check(rho.lib.bt_dm_timer_start(rho.h))
    if fused:
        bt.apply(ops3, rho)
    else:
        for o in ops3: bt.apply(rho, o)
    L.check(rho.lib.bt_dm_timer_stop(rho.h, C.byref(ms)))
    nl = rho.launch_count() - n0
    tr = bt._lib.bt_c64(); L.check(rho.lib.bt_dm_trace(rho.h, C.byref(tr)))
    ez = bt.expect(rho, "Z")
    print(f"C3 14q DM depth 20: {ngate} gates + {nch} Kraus channels, {'fused superoperators' if fused else 'op by op'}: {ms.value:.1f} ms, {nl} passes, "
          f"{len(ops3)/ms.value*1e3:.0f} ops/s, {32*4**N*nl/ms.value/1e6:.0f} GB/s algorithmic, trace={tr.re:.12f}, sum<Z>={ez.sum():.9f}")
    del rho

# ---- C4: 512 trajectories x 20 qubits, monitored brickwork ---------------------------------------------------------
N, T = 20, 512
specs4, M = wl.c4_monitored(N, 20, 20)
ops4 = wl.to_ops(bt, specs4)
U = np.random.Generator(np.random.PCG64(20)).random((T, M))
best = 1e9
for _ in range(3):
    st = bt.zero_state(N, T); st.sync()
    t0 = time.perf_counter()
    _, mids = bt.apply(ops4, st, rng=bt.BatchDraws(U), track_measurements=True)
    st.sync(); best = min(best, time.perf_counter() - t0)
    nrm = bt.norm2(st); del st
print(f"C4 20q monitored brickwork depth 20 ({len(ops4)-M} gates, {M} mid-circuit measurements), {T} trajectories in one batch (8 GiB): "
      f"{best*1e3:.0f} ms -> {T/best:.0f} trajectories/s per GPU ({8*T/best:.0f}/s on 8 GPUs, no communication); max |norm-1| = {np.max(np.abs(nrm-1)):.1e}")
