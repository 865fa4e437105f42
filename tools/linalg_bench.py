#!/usr/bin/env python
"""Timings of the round-2 device paths (host wall clock around the C-ABI call, best of 3): Schmidt spectrum / entanglement entropy,
partial_trace with many kept qubits, expect of a controlled op, multi-qubit measurement vs sequential, batched sampling."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r


def scrambled(N, nb=1):
    s = bt.zero_state(N, nb)
    arr = bt.pack_gates(wl.to_ops(bt, wl.layered(N, 8, 5)))
    L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
    s.sync()
    return s


print("# entanglement_entropy (one-sided Jacobi on the device; sweeps to convergence)")
for N, nb in ((12, 1), (16, 1), (20, 1), (22, 1), (24, 1), (12, 512), (16, 64), (20, 16)):
    s = scrambled(N, nb)
    sw = C.c_int()
    spec = np.empty((nb, 1 << (N // 2)))
    ms, _ = best(lambda: L.check(s.lib.bt_sv_schmidt_spectrum(s.h, N // 2, L.pdouble(spec), C.byref(sw))))
    ent = float(np.sum(-spec[0][spec[0] > 0] * np.log(spec[0][spec[0] > 0])))
    print(f"  N={N:2d} batch={nb:4d}: {ms:9.2f} ms  ({sw.value} sweeps, {(1 << (N // 2)) - 1} rounds per sweep)  S[0] = {ent:.6f}")
    del s
print("# partial_trace(state, keep) with k kept qubits at 28 qubits (tiled Gram kernel; k <= 3: k_rdm)")
s = scrambled(28)
for k in (2, 3, 4, 6, 8, 10):
    keep = list(range(3, 3 + 2 * k, 2))[:k]
    ms, rho = best(lambda: bt.partial_trace(s, keep))
    print(f"  k={k:2d} keep={keep}: {ms:9.2f} ms  tr = {np.trace(rho).real:.12f}   ({16 * (1 << 28) / ms / 1e6:7.0f} GB/s of state read)")
print("# expect(state, controlled op) at 28 qubits")
ms, v = best(lambda: bt.expect(s, bt.Op("X", 9, control=20)))
print(f"  <C-X>: {ms:.2f} ms  value {v:+.3e}")
print("# k Z-measurements on distinct qubits at 28 qubits: one multi call vs k sequential calls")
for k in (1, 2, 4):
    qs = [5, 11, 17, 23][:k]
    u = np.random.default_rng(k).random((1, k))
    out = np.empty((1, k), dtype=np.int32)

    def multi():
        t = s.copy()
        t.sync()
        t0 = time.perf_counter()
        L.check(t.lib.bt_sv_measure_z_multi(t.h, k, (C.c_int * k)(*qs), L.pdouble(u), out.ctypes.data_as(C.POINTER(C.c_int32)), None))
        return (time.perf_counter() - t0) * 1e3

    def seq():
        t = s.copy()
        t.sync()
        t0 = time.perf_counter()
        o = np.empty(1, dtype=np.int32)
        for j in range(k):
            L.check(t.lib.bt_sv_measure_z(t.h, qs[j], L.pdouble(np.ascontiguousarray(u[:, j])), o.ctypes.data_as(C.POINTER(C.c_int32)), None, 0))
        return (time.perf_counter() - t0) * 1e3

    a, b = min(multi() for _ in range(3)), min(seq() for _ in range(3))
    print(f"  k={k}: multi {a:6.2f} ms ({(16 + 32) * (1 << 28) / a / 1e6:5.0f} GB/s counted as one read + one write pass)   sequential {b:6.2f} ms")
del s
print("# batched sampling: 512 trajectories x 20 qubits, 64 shots each (one launch set)")
sb = scrambled(20, 512)
us = np.random.default_rng(0).random((512, 64))
ms, _ = best(lambda: bt.sample(sb, 64, uniforms=us))
print(f"  {ms:.2f} ms  ({16 * 512 * (1 << 20) / ms / 1e6:.0f} GB/s of state read)")
