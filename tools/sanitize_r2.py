"""Round-2 kernels under compute-sanitizer (memcheck / racecheck): specialised passes in both frames (single tile and pipelined),
the interpreter, multi-qubit measurement, segmented sampling, the Jacobi / Gram / density-matrix reductions, wide-load gate kernels."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

bt = ge.load_package()
L = bt._lib
from importlib import import_module

wl = import_module(ge.PKG_NAME + ".workloads")
N = 14
rng = np.random.default_rng(1)
for env in ({"BT_TILE_JIT": "0"}, {"BT_TILE_JIT": "2"}, {"BT_TILE_JIT": "2", "BT_JIT_VARIANT": "0"}, {"BT_TILE_JIT": "2", "BT_TILE_BITS": "11", "BT_TILE_PIPE": "2"}):
    os.environ.update(env)
    s = bt.zero_state(N)
    bt.apply(wl.to_ops(bt, wl.qft(N) + wl.layered(N, 4, 3)), s)
    print(env, "norm", float(bt.norm2(s)), flush=True)
    for k in env:
        os.environ.pop(k)
v = rng.normal(size=(3, 1 << 10)) + 1j * rng.normal(size=(3, 1 << 10))
v /= np.linalg.norm(v, axis=1, keepdims=True)
b = bt.CuState.from_numpy(v)
print("entropy", bt.entanglement_entropy(b), flush=True)
print("rdm5 trace", np.trace(bt.partial_trace(b, [1, 3, 4, 8, 10])[0]).real, flush=True)
out = np.empty((3, 3), dtype=np.int32)
L.check(b.lib.bt_sv_measure_z_multi(b.h, 3, (C.c_int * 3)(2, 10, 5), L.pdouble(np.ascontiguousarray(rng.random((3, 3)))), out.ctypes.data_as(C.POINTER(C.c_int32)), (C.c_int * 3)(0, 1, 0)))
print("multi measure", out.tolist(), flush=True)
print("batched samples", bt.sample(b, 5, uniforms=rng.random((3, 5))).tolist(), flush=True)
for q in (10, 9, 1):
    bt.apply(b, bt.Op("H", q))
    bt.apply(b, bt.Op("CNOT", q, 10 if q != 10 else 9))
rho = bt.CuRho.from_numpy(np.outer(v[0][:64], v[0][:64].conj()) / np.vdot(v[0][:64], v[0][:64]).real)
bt.apply(rho, bt.Op("CNOT", 5, 6))
bt.apply(rho, bt.OpQC("depolarizing", 0.1, 5, 6))
print("rho: <CX>", bt.expect(rho, bt.Op("X", 6, control=5)), "S", bt.entanglement_entropy(rho)[0], "tr rdm", np.trace(bt.partial_trace_rho(rho, [2, 5])).real, flush=True)
