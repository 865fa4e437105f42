"""oracle/strided.py -- TEST INFRASTRUCTURE ONLY (see oracle/bt_oracle.py header).

Large-N CPU oracle: the same reference semantics as oracle/bt_oracle.py (op descriptors, draw order, Kraus
selection incl. the qubit>target quirk, measurement projection + own-norm normalisation), with the 2^N x 2^N
sparse-matrix arithmetic replaced by the in-place strided C routines of oracle/strided_cpu.c.  Validated against
the kron-chain restatement at small N in tests/test_oracle_strided.py; then used where the kron chain cannot go
(N up to ~26) and as bench.py's "port" CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import bt_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbt_oracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "oracle"], cwd=os.path.dirname(_HERE), check=True)
        l = C.CDLL(_SO)
        l.bto_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        l.bto_rdm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        l.bto_norm2.argtypes = [C.c_void_p, C.c_uint64]
        l.bto_norm2.restype = C.c_double
        l.bto_scale.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
        l.bto_scale.restype = None
        l.bto_probs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        l.bto_probs.restype = C.c_double
        l.bto_axpy.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        l.bto_axpy.restype = None
        l.bto_num_threads.restype = C.c_int
        l.bto_set_threads.argtypes = [C.c_int]
        l.bto_set_threads.restype = None
        _lib = l
    return _lib


def num_threads() -> int:
    return int(lib().bto_num_threads())


def use_all_cores() -> int:
    """OpenMP threads := the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().bto_set_threads(max(1, n))
    return num_threads()


def _ints(v: Sequence[int]):
    return (C.c_int * max(1, len(v)))(*v) if len(v) else (C.c_int * 1)(0)


def apply_bits(vec: np.ndarray, n_bits: int, tb: Sequence[int], m: np.ndarray, cb: Sequence[int] = ()) -> None:
    """In place: matrix m (row-major, index bit t <-> tb[t]) on vec, restricted to control bits cb all 1."""
    mm = np.ascontiguousarray(m, dtype=np.complex128)
    rc = lib().bto_apply(vec.ctypes.data, n_bits, len(tb), _ints(list(tb)), mm.ctypes.data, len(cb), _ints(list(cb)))
    assert rc == 0


def rdm_bits(vec: np.ndarray, n_bits: int, tb: Sequence[int]) -> np.ndarray:
    D = 1 << len(tb)
    out = np.empty((D, D), dtype=np.complex128)
    rc = lib().bto_rdm(vec.ctypes.data, n_bits, len(tb), _ints(list(tb)), out.ctypes.data)
    assert rc == 0
    return out


def _gate_bits(N: int, op) -> Tuple[List[int], List[int]]:
    """(target bits, control bits) with matrix index = 2*b_qubit + b_target (src/hilbert.jl:30)."""
    if op.target_qubit == -1:
        tb = [N - op.qubit]
    else:
        tb = [N - op.target_qubit, N - op.qubit]
    cb = [N - op.control] if op.control != -2 else []
    return tb, cb


class SV:
    """State vector with the reference's semantics and strided arithmetic."""

    def __init__(self, N: int, vec: Optional[np.ndarray] = None):
        self.N = N
        if vec is None:
            vec = np.zeros(1 << N, dtype=np.complex128)
            vec[0] = 1
        self.v = np.ascontiguousarray(vec, dtype=np.complex128).copy()

    def gate(self, mat: np.ndarray, tb, cb=()):
        apply_bits(self.v, self.N, tb, mat, cb)

    def norm2(self) -> float:
        return float(lib().bto_norm2(self.v.ctypes.data, self.v.size))

    def normalize(self):
        lib().bto_scale(self.v.ctypes.data, self.v.size, 1.0 / np.sqrt(self.norm2()))

    # src/linalg.jl:167-230, :83-86 ------------------------------------------------------------------------
    def partial_trace(self, qubits: Sequence[int], ordered_min_max: bool = True) -> np.ndarray:
        qs = sorted(qubits) if ordered_min_max else list(qubits)
        # matrix index MSB = first listed qubit
        tb = [self.N - q for q in reversed(qs)]
        return rdm_bits(self.v, self.N, tb)

    # src/hilbert.jl:682-696 ----------------------------------------------------------------------------------
    def born_measure_Z(self, qubit: int, draws: O.Draws) -> int:
        prob0 = float(np.real(self.partial_trace([qubit])[0, 0]))
        ind = 0 if draws.uniform() < prob0 else 1
        self.gate(O.GATE["P0"] if ind == 0 else O.GATE["P1"], [self.N - qubit])
        self.normalize()
        return ind

    # src/struct.jl:9-41 -----------------------------------------------------------------------------------------
    def channel(self, kraus, qubit: int, target: int, draws: O.Draws) -> int:
        if target == -1:
            pA = self.partial_trace([qubit])
        else:
            pA = self.partial_trace([qubit, target])  # (min,max) order whatever (qubit,target) is: the reference's quirk
        probs = [float(np.real(np.trace(K @ pA @ K.conj().T))) for K in kraus]
        ind = O.weighted_sample(probs, draws)
        if ind is None:
            raise RuntimeError("_weighted_sample returned nothing")
        K = kraus[ind]
        tb = [self.N - qubit] if target == -1 else [self.N - target, self.N - qubit]
        self.gate(K, tb)
        self.normalize()
        return ind

    def channel3(self, kraus, first: int, draws: O.Draws) -> int:
        pA = self.partial_trace([first, first + 1, first + 2])
        probs = [float(np.real(np.trace(K @ pA @ K.conj().T))) for K in kraus]
        ind = O.weighted_sample(probs, draws)
        self.gate(kraus[ind], [self.N - first - 2, self.N - first - 1, self.N - first])
        self.normalize()
        return ind

    # src/hilbert.jl:469-515 ------------------------------------------------------------------------------------------
    def apply(self, op, noise=False, draws: Optional[O.Draws] = None, mids: Optional[list] = None):
        N = self.N
        if isinstance(op, O.OpQC):
            if op.name.upper() in ("RES", "RESET"):
                ind = self.born_measure_Z(op.qubit, draws)
                if ind == 1:
                    self.gate(O.GATE["X"], [N - op.qubit])
            elif op.q == 3:
                self.channel3(op.kraus, op.qubit, draws)
            else:
                self.channel(op.kraus, op.qubit, op.target_qubit, draws)
        elif op.type == "🔬":
            rname = O.resolve_measurement_name(op.name, draws)
            rot = O.measurement_mat(rname)
            self.gate(rot, [N - op.qubit])
            ind = self.born_measure_Z(op.qubit, draws)
            self.gate(rot.conj().T, [N - op.qubit])
            if isinstance(op, O.ifOp):
                for o in op.if0 if ind == 0 else op.if1:
                    self.apply(o, noise=noise, draws=draws)
            if mids is not None:
                mids.append(ind)
        else:
            tb, cb = _gate_bits(N, op)
            self.gate(op.mat, tb, cb)
        if isinstance(noise, O.NoiseModel) and getattr(op, "noisy", False) is True:
            if op.q == 1:
                if op.control == -2:
                    self.channel(noise.q1.kraus, op.qubit, -1, draws)
                else:
                    self.channel(noise.q2.kraus, op.control, op.qubit, draws)
            elif op.q == 2:
                self.channel(noise.q2.kraus, op.qubit, op.target_qubit, draws)

    def apply_ops(self, ops, noise=False, draws=None, track_measurements=False):
        mids: list = []
        for o in ops:
            self.apply(o, noise=noise, draws=draws, mids=mids)
        return (self, mids) if track_measurements else self

    def probs(self) -> np.ndarray:
        p = np.empty(self.v.size, dtype=np.float64)
        lib().bto_probs(self.v.ctypes.data, self.v.size, p.ctypes.data)
        return p

    def expect_z_all(self) -> np.ndarray:
        p = self.probs()
        out = np.empty(self.N)
        idx = np.arange(p.size, dtype=np.uint64)
        for q in range(1, self.N + 1):
            sign = 1.0 - 2.0 * ((idx >> np.uint64(self.N - q)) & np.uint64(1)).astype(np.float64)
            out[q - 1] = float(np.dot(sign, p))
        return out


class DM:
    """Density matrix as a 2N-bit vector (column-major rho): U on row bits, conj(U) on column bits; a channel is the
    explicit sum over Kraus operators, as src/struct.jl:58-76 does it."""

    def __init__(self, N: int, rho: Optional[np.ndarray] = None):
        self.N = N
        if rho is None:
            v = np.zeros(1 << (2 * N), dtype=np.complex128)
            v[0] = 1
        else:
            v = np.asarray(rho, dtype=np.complex128).reshape(-1, order="F").copy()
        self.v = np.ascontiguousarray(v)

    def to_matrix(self) -> np.ndarray:
        d = 1 << self.N
        return self.v.reshape((d, d), order="F")

    def _two_sided(self, vec, mat, tb, cb=()):
        apply_bits(vec, 2 * self.N, tb, mat, cb)
        apply_bits(vec, 2 * self.N, [b + self.N for b in tb], np.conj(mat), [b + self.N for b in cb])

    def apply(self, op, noise=False):
        N = self.N
        if isinstance(op, O.OpQC):
            self.channel(op.kraus, op.qubit, op.target_qubit)
        elif op.type == "🔬":
            raise RuntimeError("fix this:")
        else:
            tb, cb = _gate_bits(N, op)
            self._two_sided(self.v, op.mat, tb, cb)
        if isinstance(noise, O.NoiseModel) and getattr(op, "noisy", False) is True:
            if op.q == 1:
                if op.control == -2:
                    self.channel(noise.q1.kraus, op.qubit, -1)
                else:
                    self.channel(noise.q2.kraus, op.control, op.qubit)
            elif op.q == 2:
                self.channel(noise.q2.kraus, op.qubit, op.target_qubit)

    def channel(self, kraus, qubit: int, target: int = -1):
        tb = [self.N - qubit] if target == -1 else [self.N - target, self.N - qubit]
        acc = np.zeros_like(self.v)
        for K in kraus:
            t = self.v.copy()
            self._two_sided(t, K, tb)
            lib().bto_axpy(acc.ctypes.data, t.ctypes.data, acc.size)
        self.v = acc
