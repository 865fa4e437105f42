"""Host-side mirror of BlueTangle.jl's front end for DEVICE-RESIDENT states.

The reference dispatches on the state type (``apply(::AbstractVectorS, op)`` src/hilbert.jl:469,
``apply(::SparseMatrixCSC, op)`` :639, ``apply(::MPS, op)`` :566).  ``CuState`` / ``CuRho`` are the new state
types; every function below keeps the reference's name, argument meaning and error behaviour and forwards to
libbluetangle_cuda.so through the C ABI (``_lib.py``).  Julia is not available in this container, so this Python
layer is what the parity tests drive; ``julia/BlueTangleCUDA.jl`` is the same layer written as Julia methods.

No CPU fallback anywhere: a missing library or device raises.
Random numbers: the reference calls Julia's global ``rand()``; here every function that consumes draws takes
``rng`` (anything with ``uniform()`` / ``randint(n)``), default = the module-level generator (``seed!``).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib as L
from .gates import gate, gates, gates_with_phase, clean_name, is_measurement, noise_model, is_valid_quantum_channel

c128 = np.complex128


# ----------------------------------------------------------------------------------------------------------
# draws
# ----------------------------------------------------------------------------------------------------------
class Draws:
    """Uniform source.  ``uniform()`` <-> Julia ``rand()``; ``randint(n)`` <-> ``rand(1:n)`` (0-based here)."""

    def __init__(self, seed=None):
        self.g = seed if isinstance(seed, np.random.Generator) else np.random.Generator(np.random.PCG64(seed))

    def uniform(self) -> float:
        return float(self.g.random())

    def randint(self, n: int) -> int:
        return int(self.g.integers(0, n))


class BatchDraws:
    """Per-trajectory draw streams for batched states: trajectory t consumes U[t, 0], U[t, 1], ... in order,
    exactly what a sequential per-shot loop (src/ops.jl:671-676) fed with U[t, :] would consume."""

    def __init__(self, U: np.ndarray):
        self.U = np.ascontiguousarray(U, dtype=np.float64)
        self.cnt = np.zeros(self.U.shape[0], dtype=np.int64)

    def take(self, mask: Optional[np.ndarray] = None) -> np.ndarray:
        T = self.U.shape[0]
        idx = np.minimum(self.cnt, self.U.shape[1] - 1)
        u = self.U[np.arange(T), idx].copy()
        if mask is None:
            if np.any(self.cnt >= self.U.shape[1]):
                raise IndexError("BatchDraws exhausted")
            self.cnt += 1
        else:
            if np.any(self.cnt[mask] >= self.U.shape[1]):
                raise IndexError("BatchDraws exhausted")
            self.cnt[mask] += 1
        return u


_global_rng = Draws()


def seed(n: int) -> None:
    """``Random.seed!(n)`` counterpart for the module-level generator."""
    global _global_rng
    _global_rng = Draws(n)


def _rng(r):
    return _global_rng if r is None else r


# ----------------------------------------------------------------------------------------------------------
# bit helpers (src/bit.jl:9-48)
# ----------------------------------------------------------------------------------------------------------
def int2bin(number: int, N: int) -> List[int]:
    return [(number >> (N - i)) & 1 for i in range(1, N + 1)]


def bin2int(v: Sequence[int]) -> int:
    r = 0
    for b in v:
        r = (r << 1) + int(b)
    return r


fock_basis = int2bin


def mag_basis(number: int, N: int) -> List[int]:
    return [1 - 2 * b for b in int2bin(number, N)]


# ----------------------------------------------------------------------------------------------------------
# device-resident states
# ----------------------------------------------------------------------------------------------------------
class CuState:
    """Device-resident state vector(s): ``n_batch`` trajectories of 2^N ComplexF64 amplitudes."""

    def __init__(self, N: int, n_batch: int = 1):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.bt_sv_create(int(N), int(n_batch), C.byref(h)))
        self.h = h
        self.N = int(N)
        self.n_batch = int(n_batch)
        self.mask: Optional[np.ndarray] = None  # batched states: trajectories the next ops act on (ifOp branches), None = all

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.bt_sv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_mask(self, mask: Optional[np.ndarray]) -> None:
        """bt_sv_set_mask: gates, Kraus steps and measurements then touch only the trajectories with mask[t] true; draws of a
        BatchDraws source are consumed for those trajectories only."""
        if mask is None:
            L.check(self.lib.bt_sv_set_mask(self.h, None))
            self.mask = None
            return
        m = np.ascontiguousarray(np.asarray(mask).astype(bool))
        if m.shape != (self.n_batch,):
            raise ValueError("mask must have one entry per trajectory")
        mi = m.astype(np.int32)
        L.check(self.lib.bt_sv_set_mask(self.h, mi.ctypes.data_as(C.POINTER(C.c_int32))))
        self.mask = m

    # transfers -------------------------------------------------------------------------------------------
    @staticmethod
    def from_numpy(vec: np.ndarray) -> "CuState":
        v = np.ascontiguousarray(vec, dtype=c128)
        if v.ndim == 1:
            nb, n = 1, v.shape[0]
        else:
            nb, n = v.shape
        N = int(round(math.log2(n)))
        if 1 << N != n:
            raise ValueError("state length must be a power of two")
        s = CuState(N, nb)
        L.check(s.lib.bt_sv_upload(s.h, L.ptr(v), v.size))
        return s

    def to_numpy(self) -> np.ndarray:
        out = np.empty((self.n_batch, 1 << self.N), dtype=c128)
        L.check(self.lib.bt_sv_download(self.h, L.ptr(out), out.size))
        return out[0] if self.n_batch == 1 else out

    def copy(self) -> "CuState":
        s = CuState(self.N, self.n_batch)
        L.check(self.lib.bt_sv_copy(s.h, self.h))
        return s

    def sync(self) -> None:
        L.check(self.lib.bt_sv_sync(self.h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        L.check(self.lib.bt_sv_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def __len__(self):
        return 1 << self.N


class CuRho:
    """Device-resident density matrix (2^N x 2^N, column-major like Julia)."""

    def __init__(self, N: int):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.bt_dm_create(int(N), C.byref(h)))
        self.h = h
        self.N = int(N)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.bt_dm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def from_numpy(rho: np.ndarray) -> "CuRho":
        r = np.asarray(rho, dtype=c128)
        N = int(round(math.log2(r.shape[0])))
        d = CuRho(N)
        f = np.asfortranarray(r)
        L.check(d.lib.bt_dm_upload(d.h, L.ptr(f), f.size))
        return d

    @staticmethod
    def from_state(state: CuState) -> "CuRho":
        """``state*state'`` (src/ops.jl:810)."""
        d = CuRho(state.N)
        L.check(d.lib.bt_dm_from_sv(d.h, state.h))
        return d

    def to_numpy(self) -> np.ndarray:
        dim = 1 << self.N
        out = np.empty((dim, dim), dtype=c128, order="F")
        L.check(self.lib.bt_dm_download(self.h, L.ptr(out), out.size))
        return out

    def sync(self) -> None:
        L.check(self.lib.bt_dm_sync(self.h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        L.check(self.lib.bt_dm_launch_count(self.h, C.byref(n)))
        return int(n.value)


State = Union[CuState, CuRho]


def get_N(x: State) -> int:
    """src/ops.jl:5-9."""
    return x.N


# state constructors (src/hilbert.jl:835-882) ---------------------------------------------------------------
def zero_state(N: int, n_batch: int = 1) -> CuState:
    return CuState(N, n_batch)


def basis_state(N: int, index: int, n_batch: int = 1) -> CuState:
    s = CuState(N, n_batch)
    L.check(s.lib.bt_sv_set_basis(s.h, int(index)))
    return s


def one_state(N: int, n_batch: int = 1) -> CuState:
    return basis_state(N, (1 << N) - 1, n_batch)


def plus_state(N: int, n_batch: int = 1) -> CuState:
    s = CuState(N, n_batch)
    L.check(s.lib.bt_sv_set_plus(s.h))
    return s


def product_state(list_of_qubits: Sequence[int], n_batch: int = 1) -> CuState:
    return basis_state(len(list_of_qubits), bin2int([1 if b > 0 else 0 for b in list_of_qubits]), n_batch)


def neel_state01(N: int) -> CuState:
    return product_state([0 if i % 2 == 1 else 1 for i in range(1, N + 1)])


def neel_state10(N: int) -> CuState:
    return product_state([1 if i % 2 == 1 else 0 for i in range(1, N + 1)])


def minus_state(N: int) -> CuState:
    s = one_state(N)
    for i in range(1, N + 1):
        apply(s, Op("H", i))
    return s


def random_state(N: int, gen: np.random.Generator) -> CuState:
    """src/hilbert.jl:882: normalised vector of uniform [0,1) real and imaginary parts."""
    v = gen.random(1 << N) + 1j * gen.random(1 << N)
    return CuState.from_numpy(v / np.linalg.norm(v))


# ----------------------------------------------------------------------------------------------------------
# op descriptors (src/struct.jl)
# ----------------------------------------------------------------------------------------------------------
class QuantumOps:
    pass


class Op(QuantumOps):
    """``Op(name, qubit[, target]; control, noisy)`` / ``Op(name, mat, qubit[, target])`` /
    ``Op("CCX", qubit, control, target)`` -- src/struct.jl:365-474."""

    def __new__(cls, name, *args, **kw):
        if isinstance(name, str) and name.upper() in ("RES", "RESET") and len(args) == 1 and isinstance(args[0], (int, np.integer)):
            # src/struct.jl:421-423: amplitude damping with gamma = 1, executed as measure-then-X on a state vector
            return OpQC("RES", [np.array([[1, 0], [0, 0]], dtype=c128), np.array([[0, 1], [0, 0]], dtype=c128)], int(args[0]))
        return super().__new__(cls)

    def __init__(self, name, *args, control: int = -2, noisy: bool = True, type: str = ""):
        if isinstance(name, (list, tuple)):  # Op(["RZ", 0.3], q) src/struct.jl:550-551
            arg = name[1]
            arg_s = ",".join(repr(float(a)) for a in arg) if isinstance(arg, (list, tuple)) else repr(float(arg))
            name = f"{name[0]}({arg_s})"
        mat = None
        rest = list(args)
        if rest and not isinstance(rest[0], (int, np.integer)):
            mat = np.asarray(rest.pop(0), dtype=c128)
        rest = [int(r) for r in rest]
        if len(rest) == 3:
            # three-qubit constructor src/struct.jl:436-449: (qubit, control_qubit, target_qubit)
            m3 = {"CCZ": "CZ", "CCX": "CX", "CCY": "CY", "CSWAP": "SWAP"}
            if name not in m3:
                raise ValueError("Unsupported three-qubit operation")
            name = m3[name]
            control = rest[1]
            rest = [rest[0], rest[2]]
        if len(rest) == 1:
            rest.append(-1)
        if len(rest) != 2:
            raise ValueError("Op needs a qubit (and optionally a target qubit)")
        self.name = name
        self.qubit, self.target_qubit = rest
        self.control = int(control)
        # _get_op_num_qubits src/struct.jl:455-474
        if self.target_qubit == -1:
            if self.qubit == self.control:
                raise ValueError("`qubit` must differ from `control` qubit")
            self.q = 1
        else:
            if self.qubit == self.target_qubit:
                raise ValueError("`qubit` and `target_qubit` must differ")
            if self.qubit == self.control and self.target_qubit == self.control:
                raise ValueError("either `qubit` or `target_qubit` must differ from `control` qubit")
            self.q = 2
        self.mat = gates(name) if mat is None else mat
        if self.mat.shape != (1 << self.q, 1 << self.q):
            raise ValueError(f"size of matrix {self.mat.shape} not compatible with {self.q}-qubit operation")
        self.ismeasure = self.q == 1 and is_measurement(name)
        if self.ismeasure:
            if self.control != -2:
                raise ValueError("measurement and control operations are incompatible.")
            self.type = "🔬"
            self.noisy = False
        else:
            self.type = type or ("phase" if clean_name(name) in gates_with_phase else "op")
            self.noisy = bool(noisy)

    def __repr__(self):
        return f"Op({self.name!r}, {self.qubit}, {self.target_qubit}, control={self.control})"


class OpQC(QuantumOps):
    """Quantum channel as an op -- src/struct.jl:182-249.  ``OpQC(name, kraus, qubit[, target])`` or
    ``OpQC(model, p, qubit[, target])``."""

    def __init__(self, name: str, kraus_or_p, qubit: int, target_qubit: int = -1, type: str = ""):
        if isinstance(kraus_or_p, (float, int)):
            p = float(kraus_or_p)
            kraus = noise_model(name, p, two_qubit=target_qubit > 0)
            type = str(p)
        else:
            kraus = [np.asarray(k, dtype=c128) for k in kraus_or_p]
        if not is_valid_quantum_channel(kraus):
            raise ValueError("not valid kraus operators: not CPTP!")
        self.q = int(round(math.log2(kraus[0].shape[0])))
        if self.q == 1 and target_qubit > 0:
            raise ValueError("for 1-qubit quantum channel, target_qubit should be -1")
        if self.q == 2 and target_qubit < 0:
            raise ValueError("for 2-qubit quantum channel, you must select the target_qubit")
        if self.q > 3:
            raise ValueError("Noise models are available only for up to 3 qubits!")
        self.name = name.lower()
        self.kraus = kraus
        self.qubit = int(qubit)
        self.target_qubit = int(target_qubit) if self.q == 2 else -1
        self.control = -2
        self.noisy = False
        self.type = type

    def prob(self, state: CuState) -> np.ndarray:
        """__calc_prob src/struct.jl:9-29, :44-49."""
        return _kraus_probs(state, self.q, self.qubit, self.target_qubit, self.kraus)


class QuantumChannel:
    """src/struct.jl:90-140."""

    def __init__(self, q_or_model, model_or_p=None, p=None, kraus=None):
        if isinstance(q_or_model, str):
            q, model, p = 1, q_or_model, float(model_or_p)
        else:
            q, model, p = int(q_or_model), str(model_or_p), float(p)
        if kraus is None:
            kraus = noise_model(model, p, two_qubit=(q == 2))
        kraus = [np.asarray(k, dtype=c128) for k in kraus]
        if not is_valid_quantum_channel(kraus):
            raise ValueError("not valid kraus operators: not CPTP!")
        self.q = int(round(math.log2(kraus[0].shape[0])))
        if self.q not in (1, 2):
            raise ValueError("Noise models are available only for 1 and 2 qubits!")
        self.name = model.lower()
        self.p = p
        self.kraus = kraus


def Noise1(model: str, p: float) -> QuantumChannel:
    return QuantumChannel(1, model, p)


def Noise2(model: str, p: float) -> QuantumChannel:
    return QuantumChannel(2, model, p)


class NoiseModel:
    """src/struct.jl:147-173."""

    def __init__(self, a, b):
        if isinstance(a, str):
            a, b = Noise1(a, float(b)), Noise2(a, float(b))
        if isinstance(a, QuantumChannel) and a.q != 1:
            raise ValueError("size error in noise model (1)")
        if isinstance(b, QuantumChannel) and b.q != 2:
            raise ValueError("size error in noise model (2)")
        self.q1, self.q2 = a, b


class ifOp(QuantumOps):
    """Mid-circuit measurement with conditional op lists -- src/struct.jl:666-696."""

    def __init__(self, name: str, qubit: int, if0="I", if1="I"):
        if not is_measurement(name):
            raise ValueError("select MX or MY or MZ or MR basis.")

        def as_list(x):
            if isinstance(x, str):
                return [Op(x, qubit)]
            if isinstance(x, Op):
                return [x]
            return list(x)

        self.name, self.qubit = name, int(qubit)
        self.if01 = (as_list(if0), as_list(if1))
        self.q, self.type = 1, "🔬"
        self.mat = gates(name)


# ----------------------------------------------------------------------------------------------------------
# low-level calls
# ----------------------------------------------------------------------------------------------------------
class OpF(QuantumOps):
    """``OpF(name, f)`` / ``OpF(name, ops)`` -- src/struct.jl:703-744: an op that is a function of the state.  ``f`` receives the
    device-resident state and may update it in place (returning None or the state) or return another device state; an op list
    is applied in order (the reference's ``o * state`` loop, :725-741) as one fused device call.  The third form of the
    reference, a full 2^N x 2^N matrix (:716-721), has no device counterpart -- the backend never builds operators of the
    whole register -- and raises."""

    def __init__(self, name: str, data):
        self.q = 1
        self.name = name
        self.type = ""
        self.noisy = False
        self.data = data
        if callable(data):
            self.apply = lambda state, **kw: data(state, **kw)
        elif isinstance(data, (list, tuple)) and all(isinstance(o, QuantumOps) for o in data):
            ops = list(data)
            self.apply = lambda state, **kw: apply(ops, state)
        else:
            raise NotImplementedError("OpF with a full-register matrix (src/struct.jl:716-721) is not available on the device backend: pass the op list or a function of the state")

    def __repr__(self):
        return f"OpF({self.name!r})"


def _apply_matrix(x: State, q: int, mat: np.ndarray, qubit: int, target: int, control: int, want: Optional[int] = None) -> None:
    lib = x.lib
    if isinstance(x, CuRho):
        if q == 1:
            L.check(lib.bt_dm_apply_1q(x.h, qubit, L.ptr(L.cmat(mat, 2)), control))
        else:
            L.check(lib.bt_dm_apply_2q(x.h, qubit, target, L.ptr(L.cmat(mat, 4)), control))
        return
    if want is None:
        if q == 1:
            L.check(lib.bt_sv_apply_1q(x.h, qubit, L.ptr(L.cmat(mat, 2)), control))
        else:
            L.check(lib.bt_sv_apply_2q(x.h, qubit, target, L.ptr(L.cmat(mat, 4)), control))
    else:
        if q == 1:
            L.check(lib.bt_sv_apply_1q_if(x.h, qubit, L.ptr(L.cmat(mat, 2)), control, want))
        else:
            L.check(lib.bt_sv_apply_2q_if(x.h, qubit, target, L.ptr(L.cmat(mat, 4)), control, want))


def _kraus_table(kraus: Sequence[np.ndarray], q: int) -> np.ndarray:
    D = 1 << q
    tab = np.empty((len(kraus), D * D), dtype=c128)
    for i, K in enumerate(kraus):
        tab[i] = L.cmat(K, D).reshape(-1, order="F")
    return np.ascontiguousarray(tab)


def _kraus_probs(state: CuState, q: int, qubit: int, target: int, kraus) -> np.ndarray:
    tab = _kraus_table(kraus, q)
    out = np.empty((state.n_batch, len(kraus)), dtype=np.float64)
    L.check(state.lib.bt_sv_kraus_probs(state.h, q, qubit, target, L.ptr(tab), len(kraus), L.pdouble(out)))
    return out[0] if state.n_batch == 1 else out


def _uniforms(state: CuState, rng) -> np.ndarray:
    if isinstance(rng, BatchDraws):
        return rng.take(state.mask)
    r = _rng(rng)
    return np.array([r.uniform() for _ in range(state.n_batch)], dtype=np.float64)


def _channel_apply(x: State, kraus, q: int, qubit: int, target: int, rng) -> Optional[np.ndarray]:
    """__QuantumChannel_new_apply: trajectory step on a state vector (src/struct.jl:31-55), exact channel on a
    density matrix (src/struct.jl:58-76)."""
    tab = _kraus_table(kraus, q)
    if isinstance(x, CuRho):
        L.check(x.lib.bt_dm_kraus(x.h, q, qubit, target, L.ptr(tab), len(kraus)))
        return None
    u = _uniforms(x, rng)
    chosen = np.empty(x.n_batch, dtype=np.int32)
    L.check(x.lib.bt_sv_kraus(x.h, q, qubit, target, L.ptr(tab), len(kraus), L.pdouble(u), chosen.ctypes.data_as(C.POINTER(C.c_int32))))
    return chosen


def _measurement_mat(name: str) -> np.ndarray:
    """src/struct.jl:554-563."""
    u = name.upper()
    if u in ("M(Z)", "MZ"):
        return gate["I"]
    if u in ("M(X)", "MX"):
        return gate["H"]
    if u in ("M(Y)", "MY"):
        return gate["HSP"]
    raise ValueError(name)


def _resolve_measurement_name(name: str, rng) -> str:
    """src/struct.jl:565-571: "MR" costs one discrete draw."""
    u = name.upper()
    if u in ("MR", "M(R)"):
        return ["MX", "MY", "MZ"][_rng(rng).randint(3)]
    return u


def born_measure_Z(state: CuState, qubit: int, rng=None, reset: bool = False):
    """src/hilbert.jl:682-696 (and _reset_Z :752-759 when ``reset``): returns (state, outcome[, per-trajectory array])."""
    if isinstance(state, CuRho):
        if reset:
            raise RuntimeError("reset on a density matrix is an OpQC channel")
        L.check(state.lib.bt_dm_dephase(state.h, qubit))  # src/hilbert.jl:784-796
        return state
    u = _uniforms(state, rng)
    out = np.empty(state.n_batch, dtype=np.int32)
    p0 = np.empty(state.n_batch, dtype=np.float64)
    L.check(state.lib.bt_sv_measure_z(state.h, qubit, L.pdouble(u), out.ctypes.data_as(C.POINTER(C.c_int32)), L.pdouble(p0), 1 if reset else 0))
    return state, (int(out[0]) if state.n_batch == 1 else out)


def _born_measure(state: CuState, o, rng=None):
    """src/hilbert.jl:669-679: rotate, measure Z, rotate back."""
    if isinstance(state, CuRho):
        raise RuntimeError("fix this:")  # src/hilbert.jl:772 -- unsupported in the reference
    rot = _measurement_mat(_resolve_measurement_name(o.name, rng))
    ident = np.array_equal(rot, gate["I"])
    if not ident:
        _apply_matrix(state, 1, rot, o.qubit, -1, -2)
    state, ind = born_measure_Z(state, o.qubit, rng)
    if not ident:
        _apply_matrix(state, 1, rot.conj().T, o.qubit, -1, -2)
    return state, ind


def _reset_Z(state: CuState, qubit: int, rng=None):
    return born_measure_Z(state, qubit, rng, reset=True)


def apply_noise(x: State, op, noise: NoiseModel, rng=None):
    """src/hilbert.jl:322-364."""
    if not (hasattr(op, "noisy") and op.noisy is True):
        return x
    if op.q == 1:
        if op.control == -2:
            if not isinstance(noise.q1, QuantumChannel):
                raise TypeError("noise.q1 is not a QuantumChannel (src/hilbert.jl:328-331)")
            _channel_apply(x, noise.q1.kraus, 1, op.qubit, -1, rng)
        else:
            _channel_apply(x, noise.q2.kraus, 2, op.control, op.qubit, rng)
    elif op.q == 2:
        _channel_apply(x, noise.q2.kraus, 2, op.qubit, op.target_qubit, rng)
    return x


def _ifop_apply(state: CuState, op: ifOp, noise, rng):
    """src/struct.jl:578-594."""
    if isinstance(state, CuRho):
        raise RuntimeError("error: fix this!")  # src/struct.jl:616
    rot = _measurement_mat(_resolve_measurement_name(op.name, rng))
    ident = np.array_equal(rot, gate["I"])
    if not ident:
        _apply_matrix(state, 1, rot, op.qubit, -1, -2)
    state, ind = born_measure_Z(state, op.qubit, rng)
    if not ident:
        _apply_matrix(state, 1, rot.conj().T, op.qubit, -1, -2)
    if state.n_batch == 1:
        for o in op.if01[0] if ind == 0 else op.if01[1]:
            apply(state, o, noise=noise, rng=rng)
    else:
        # batched trajectories: each branch runs under a trajectory mask (outcome == want, within the mask already active), so the
        # branch ops are the general apply(state, ifop; noise=noise) of src/struct.jl:587-590 -- gates, the branch's own noise draws
        # and nested measurements -- and a trajectory consumes draws only for the branch it took
        outer = state.mask
        ind = np.asarray(ind)
        try:
            for want in (0, 1):
                branch = [o for o in op.if01[want] if not (isinstance(o, Op) and not o.ismeasure and not (isinstance(noise, NoiseModel) and o.noisy)
                                                           and np.array_equal(o.mat, np.eye(1 << o.q)))]
                sub = ind == want  # masked-out trajectories carry outcome -1
                if not branch or not sub.any():
                    continue
                state.set_mask(sub)
                for o in branch:
                    apply(state, o, noise=noise, rng=rng)
        finally:
            state.set_mask(outer)
    return state, ind


def apply(a, b, noise=False, rng=None, track_measurements: bool = False):
    """``apply(state, op; noise, track_measurements)`` -- src/hilbert.jl:469-515 (state vector), :639-666 (density
    matrix), :517-557 (op lists and the (op, state) argument order).  The state is updated in place and returned."""
    if isinstance(a, (CuState, CuRho)):
        x, op = a, b
    else:
        x, op = b, a
    if isinstance(op, (list, tuple)) and not (isinstance(op, tuple) and op and isinstance(op[0], str)):
        if isinstance(x, CuRho) and FUSE_DEFAULT and all(isinstance(o, (Op, OpQC)) and not getattr(o, "ismeasure", False) and getattr(o, "q", 1) <= 2 for o in op):
            _dm_apply_ops(x, list(op), noise)
            return (x, []) if track_measurements else x
        mids: List = []
        groups = _coalesce(op, x, noise)
        # every measurement of the list is a plain Z measurement / reset on a state vector: outcomes are logged on the device and
        # read once at the end, so the host enqueues the whole monitored circuit without waiting for the GPU in between
        deferred = (DEFER_OUTCOMES and isinstance(x, CuState) and any(isinstance(g, _MeasureRun) for g in groups)
                    and not any(isinstance(g, (ifOp, OpF)) or (isinstance(g, Op) and g.ismeasure) or (isinstance(g, OpQC) and _z_measurement(g)) for g in groups))
        if deferred:
            L.check(x.lib.bt_sv_measure_log(x.h, 1))
        for o in groups:
            if isinstance(o, _GateRun):
                o.run(x)
                continue
            if isinstance(o, _MeasureRun):
                m = o.run(x, rng, deferred=deferred)
                if track_measurements:
                    mids.extend(m)
                continue
            if track_measurements and isinstance(x, CuState):
                x, m = apply(x, o, noise=noise, rng=rng, track_measurements=True)
                mids.extend(m)
            else:
                x = apply(x, o, noise=noise, rng=rng)
        if deferred and not track_measurements:
            L.check(x.lib.bt_sv_measure_log(x.h, 0))  # nobody asked for the outcomes: no read, no synchronisation
        elif deferred:
            runs = [g for g in groups if isinstance(g, _MeasureRun)]
            total = x.n_batch * sum(len(g.ops) for g in runs)
            log = np.empty(total, dtype=np.int32)
            n = C.c_uint64()
            L.check(x.lib.bt_sv_measure_log_read(x.h, log.ctypes.data_as(C.POINTER(C.c_int32)), total, C.byref(n)))
            L.check(x.lib.bt_sv_measure_log(x.h, 0))
            if n.value != total:
                raise RuntimeError(f"outcome log holds {n.value} entries, expected {total}")
            if track_measurements:
                pos = 0
                for g in runs:
                    k = len(g.ops)
                    blk = log[pos:pos + x.n_batch * k].reshape(x.n_batch, k)
                    pos += x.n_batch * k
                    mids.extend((int(blk[0, j]) if x.n_batch == 1 else blk[:, j].copy()) for j in range(k) if not isinstance(g.ops[j], OpQC))
        return (x, mids) if track_measurements else x
    if isinstance(op, tuple):
        op = Op(*op)
    mid: List = []
    if isinstance(op, OpF):  # src/hilbert.jl:486-487: state = op.apply(state)
        r = op.apply(x)
        if isinstance(r, (CuState, CuRho)):
            x = r
    elif isinstance(op, OpQC):
        if op.name.upper() in ("RES", "RESET") and isinstance(x, CuState):
            _reset_Z(x, op.qubit, rng)
        else:
            if isinstance(x, CuRho) and op.q == 3:
                raise NotImplementedError("3-qubit channels on density matrices are not defined in the reference")
            _channel_apply(x, op.kraus, op.q, op.qubit, op.target_qubit, rng)
    elif op.type == "🔬":
        if isinstance(op, ifOp):
            _, ind = _ifop_apply(x, op, noise, rng)
        else:
            _, ind = _born_measure(x, op, rng)
        if track_measurements:
            mid.append(ind)
    else:
        _apply_matrix(x, op.q, op.mat, op.qubit, op.target_qubit, op.control)
    if isinstance(noise, NoiseModel):
        apply_noise(x, op, noise, rng)
    return (x, mid) if track_measurements else x


def _dm_apply_ops(rho: CuRho, ops, noise) -> None:
    """apply(rho, op; noise) for a whole op list in one library call (bt_dm_apply_ops): unitaries, OpQC channels and the
    NoiseModel channels of apply_noise (src/hilbert.jl:346-364) become one fused superoperator per qubit (pair)."""
    entries, keep = [], []

    def add(kind, nq, qubit, target, control, mats):
        D = 1 << nq
        tab = np.ascontiguousarray(np.stack([L.cmat(m, D).reshape(-1, order="F") for m in mats]))
        keep.append(tab)
        entries.append((kind, nq, qubit, target, control, len(mats), tab.ctypes.data))

    for o in ops:
        if isinstance(o, OpQC):
            add(1, o.q, o.qubit, o.target_qubit, -2, o.kraus)
        else:
            add(0, o.q, o.qubit, o.target_qubit, o.control, [o.mat])
        if isinstance(noise, NoiseModel) and getattr(o, "noisy", False) is True:
            if o.q == 1 and o.control == -2:
                add(1, 1, o.qubit, -1, -2, noise.q1.kraus)
            elif o.q == 1:
                add(1, 2, o.control, o.qubit, -2, noise.q2.kraus)
            else:
                add(1, 2, o.qubit, o.target_qubit, -2, noise.q2.kraus)
    arr = (L.bt_dm_op * len(entries))()
    for i, e in enumerate(entries):
        arr[i].kind, arr[i].nq, arr[i].qubit, arr[i].target, arr[i].control, arr[i].nK, arr[i].mats = e
    L.check(rho.lib.bt_dm_apply_ops(rho.h, C.cast(arr, C.c_void_p), len(entries), 1))


class _GateRun:
    """A run of consecutive plain gates handed to the library in one call (bt_sv_apply_circuit)."""

    def __init__(self, ops: List[Op], fuse: bool):
        self.arr = pack_gates(ops)
        self.fuse = fuse

    def run(self, x: State) -> None:
        if isinstance(x, CuRho):
            L.check(x.lib.bt_dm_apply_circuit(x.h, L.ptr(self.arr), len(self.arr), int(self.fuse)))
        else:
            L.check(x.lib.bt_sv_apply_circuit(x.h, L.ptr(self.arr), len(self.arr), int(self.fuse)))


FUSE_DEFAULT = True


def pack_gates(ops: Sequence[Op]) -> np.ndarray:
    arr = np.zeros(len(ops), dtype=L.GATE_DTYPE)
    for i, o in enumerate(ops):
        D = 1 << o.q
        arr[i]["nq"], arr[i]["qubit"], arr[i]["target"], arr[i]["control"] = o.q, o.qubit, o.target_qubit, o.control
        arr[i]["m"][: D * D] = np.asarray(o.mat, dtype=c128).reshape(-1, order="F")
    return arr


class _MeasureRun:
    """Consecutive Z-basis measurements / resets on distinct qubits: ONE read pass (joint distribution of the measured bits per
    trajectory) and ONE collapse pass through ``bt_sv_measure_z_multi`` instead of one of each per measurement.  Outcomes are
    decided in op order, each on the state collapsed by the previous ones, and the draws are consumed in the same order as by
    the op-by-op loop (src/hilbert.jl:517-553 over born_measure_Z :682-696), so results and draw streams are unchanged."""

    MAXK = 4

    def __init__(self, ops):
        self.ops = list(ops)

    def run(self, x: "CuState", rng, deferred: bool = False) -> List:
        k = len(self.ops)
        u = np.empty((x.n_batch, k), dtype=np.float64)
        for j in range(k):
            u[:, j] = _uniforms(x, rng)
        qs = (C.c_int * k)(*[int(o.qubit) for o in self.ops])
        rs = (C.c_int * k)(*[1 if isinstance(o, OpQC) else 0 for o in self.ops])
        if deferred:  # outcomes go to the device-side log (bt_sv_measure_log): no host synchronisation here
            L.check(x.lib.bt_sv_measure_z_multi(x.h, k, qs, L.pdouble(u), None, rs))
            return []
        out = np.empty((x.n_batch, k), dtype=np.int32)
        L.check(x.lib.bt_sv_measure_z_multi(x.h, k, qs, L.pdouble(u), out.ctypes.data_as(C.POINTER(C.c_int32)), rs))
        # resets are channels (OpQC): they report no mid-circuit outcome (apply() only tracks type "🔬" ops)
        return [(int(out[0, j]) if x.n_batch == 1 else out[:, j].copy()) for j in range(k) if not isinstance(self.ops[j], OpQC)]


MEASURE_FUSE_DEFAULT = True
DEFER_OUTCOMES = True


def _z_measurement(o) -> bool:
    """plain Z-basis measurement (no rotation, no random basis, no branches) or a reset channel"""
    if isinstance(o, ifOp):
        return False
    if isinstance(o, OpQC):
        return o.name.upper() in ("RES", "RESET")
    return isinstance(o, Op) and o.ismeasure and o.name.upper() in ("MZ", "M(Z)")


def _coalesce(ops, x, noise):
    """Group maximal runs of plain gates (no noise attached, not measurements/channels) into single library calls, and runs of
    Z-basis measurements / resets on distinct qubits of a state vector into single multi-measurement calls."""
    out, run, mrun = [], [], []
    plain_ok = not isinstance(noise, NoiseModel)
    fuse_meas = MEASURE_FUSE_DEFAULT and isinstance(x, CuState)

    def flush_m():
        nonlocal mrun
        if mrun:
            out.append(_MeasureRun(mrun))
        mrun = []

    for o in ops:
        if isinstance(o, tuple):
            o = Op(*o)
        if plain_ok and isinstance(o, Op) and not o.ismeasure:
            flush_m()
            run.append(o)
            continue
        if run:
            out.append(_GateRun(run, FUSE_DEFAULT))
            run = []
        if fuse_meas and _z_measurement(o):
            if len(mrun) == _MeasureRun.MAXK or any(m.qubit == o.qubit for m in mrun):
                flush_m()
            mrun.append(o)
            continue
        flush_m()
        out.append(o)
    flush_m()
    if run:
        out.append(_GateRun(run, FUSE_DEFAULT))
    return out


# ----------------------------------------------------------------------------------------------------------
# reductions / observables
# ----------------------------------------------------------------------------------------------------------
def partial_trace(state: CuState, *keep) -> np.ndarray:
    """``partial_trace(state, q)`` src/linalg.jl:167-192, ``(state, q1, q2)`` :198-230 (adjacent only, like the
    reference), ``(state, [q..])`` :83-86 (kept qubits in ascending label order)."""
    lib = state.lib
    if len(keep) == 1 and isinstance(keep[0], (list, tuple)):
        qs = sorted(set(int(q) for q in keep[0]))
        general = True
    else:
        qs = [int(q) for q in keep]
        general = False
    if len(qs) == 1:
        out = np.empty((state.n_batch, 2, 2), dtype=c128)
        L.check(lib.bt_sv_rdm1(state.h, qs[0], L.ptr(out)))
        out = out.transpose(0, 2, 1)
    elif len(qs) == 2:
        if not general and abs(qs[0] - qs[1]) > 1:
            raise ValueError("must be local")
        out = np.empty((state.n_batch, 4, 4), dtype=c128)
        L.check(lib.bt_sv_rdm2(state.h, qs[0], qs[1], L.ptr(out)))
        out = out.transpose(0, 2, 1)
    elif len(qs) >= 3 and general:
        # any number of kept qubits (the library accepts up to 12): tiled Gram kernel over the gathered 2^k x 2^(N-k) matrix
        D = 1 << len(qs)
        out = np.empty((state.n_batch, D, D), dtype=c128)
        arr = (C.c_int * len(qs))(*qs)
        L.check(lib.bt_sv_rdm(state.h, len(qs), arr, L.ptr(out)))
        out = out.transpose(0, 2, 1)
    else:
        raise ValueError("partial_trace(state, q), (state, q1, q2) or (state, [q...])")
    return out[0] if state.n_batch == 1 else out


def partial_trace_rho(rho: "CuRho", keep: Sequence[int]) -> np.ndarray:
    """``partial_trace(rho, dims, trace_out)`` src/linalg.jl:88-140 for qubit registers, given the KEPT qubits (the reference
    passes the traced-out ones: ``setdiff(1:N, keep)``); kept qubits in ascending label order."""
    qs = sorted(set(int(q) for q in keep))
    D = 1 << len(qs)
    out = np.empty((D, D), dtype=c128)
    arr = (C.c_int * max(1, len(qs)))(*qs)
    L.check(rho.lib.bt_dm_rdm(rho.h, len(qs), arr, L.ptr(out)))
    return out.T


def norm2(state: CuState):
    out = np.empty(state.n_batch, dtype=np.float64)
    L.check(state.lib.bt_sv_norm2(state.h, L.pdouble(out)))
    return float(out[0]) if state.n_batch == 1 else out


def normalize(state: CuState) -> CuState:
    L.check(state.lib.bt_sv_normalize(state.h))
    return state


def inner(a: CuState, b: CuState):
    """<a|b> (src/tensor.jl:199 analogue)."""
    out = np.empty(a.n_batch, dtype=c128)
    L.check(a.lib.bt_sv_inner(a.h, b.h, L.ptr(out)))
    return complex(out[0]) if a.n_batch == 1 else out


def fidelity(a, b):
    """``fidelity(psi, psi2)`` = |<a|b>|^2 (src/tensor.jl:219); ``fidelity(rho, sigma)`` = real(tr(sqrt(sqrt(rho) sigma sqrt(rho)))^2)
    (src/tensor.jl:222-229) as dense linear algebra on the device (``bt_dm_fidelity``: up to 11 qubits)."""
    if isinstance(a, CuRho) and isinstance(b, CuRho):
        out = C.c_double()
        L.check(a.lib.bt_dm_fidelity(a.h, b.h, C.byref(out)))
        return float(out.value)
    if isinstance(a, CuRho) or isinstance(b, CuRho):
        raise TypeError("fidelity takes two state vectors or two density matrices (src/tensor.jl:219-229)")
    v = inner(a, b)
    return abs(v) ** 2


def same_up_to_global_phase(psi: CuState, phi: CuState):
    """src/linalg.jl:67-72: (abs(<phi|psi>) ≈ 1, angle(<phi|psi>)) -- one fused reduction on the device."""
    v = inner(phi, psi)
    return bool(np.isclose(abs(v), 1.0)), float(np.angle(v))


def prob(state: CuState) -> np.ndarray:
    """abs2.(state) (src/tensor.jl:150-158 wrapper of the same quantity)."""
    out = np.empty((state.n_batch, 1 << state.N), dtype=np.float64)
    L.check(state.lib.bt_sv_probs(state.h, L.pdouble(out)))
    return out[0] if state.n_batch == 1 else out


_PAULI = {"I", "X", "Y", "Z"}


def _expect_product(x: State, names: Sequence[str], qubits: Sequence[int]):
    """Re <(x)_j O_j> for gate-table names on the given qubits (expand_multi_op src/ops.jl:928-944)."""
    if len(names) != len(qubits):
        raise ValueError("qubit number does not match with operators")
    lib = x.lib
    N = x.N
    if all(n.upper() in _PAULI for n in names):
        s = ["I"] * N
        for n, q in zip(names, qubits):
            s[q - 1] = n.upper()
        ps = "".join(s).encode()
        if isinstance(x, CuRho):
            out = C.c_double()
            L.check(lib.bt_dm_expect_pauli(x.h, ps, C.byref(out)))
            return float(out.value)
        out = np.empty(x.n_batch, dtype=np.float64)
        L.check(lib.bt_sv_expect_pauli(x.h, ps, L.pdouble(out)))
        return float(out[0]) if x.n_batch == 1 else out
    mats = np.empty((len(names), 4), dtype=c128)
    for i, n in enumerate(names):
        mats[i] = L.cmat(gates(n), 2).reshape(-1, order="F")
    qs = (C.c_int * len(qubits))(*[int(q) for q in qubits])
    if isinstance(x, CuRho):
        out = C.c_double()
        L.check(lib.bt_dm_expect_product(x.h, len(names), qs, L.ptr(mats), C.byref(out)))
        return float(out.value)
    out = np.empty(x.n_batch, dtype=np.float64)
    L.check(lib.bt_sv_expect_product(x.h, len(names), qs, L.ptr(mats), L.pdouble(out)))
    return float(out[0]) if x.n_batch == 1 else out


def expect(x: State, what):
    """src/func.jl:91-101: ``expect(state, op::Op)`` -> scalar; ``expect(state, "Z")`` -> per-qubit vector."""
    lib = x.lib
    if isinstance(what, str):
        m = L.cmat(gates(what), 2)
        if isinstance(x, CuRho):
            out = np.empty(x.N, dtype=np.float64)
            L.check(lib.bt_dm_expect_1q_all(x.h, L.ptr(m), L.pdouble(out)))
            return out
        out = np.empty((x.n_batch, x.N), dtype=np.float64)
        L.check(lib.bt_sv_expect_1q_all(x.h, L.ptr(m), L.pdouble(out)))
        return out[0] if x.n_batch == 1 else out
    if isinstance(what, Op):
        if what.control != -2 or (what.q == 2 and isinstance(x, CuRho)):
            # controlled operators and 2-qubit operators on rho: trace against the reduced density matrix of the touched qubits
            m = L.cmat(what.mat, 1 << what.q)
            if isinstance(x, CuRho):
                out = C.c_double()
                L.check(lib.bt_dm_expect_op(x.h, what.q, what.qubit, what.target_qubit, what.control, L.ptr(m), C.byref(out)))
                return float(out.value)
            out = np.empty(x.n_batch, dtype=np.float64)
            L.check(lib.bt_sv_expect_op(x.h, what.q, what.qubit, what.target_qubit, what.control, L.ptr(m), L.pdouble(out)))
            return float(out[0]) if x.n_batch == 1 else out
        if what.q == 1:
            m = L.cmat(what.mat, 2)
            if isinstance(x, CuRho):
                out = C.c_double()
                qs = (C.c_int * 1)(what.qubit)
                L.check(lib.bt_dm_expect_product(x.h, 1, qs, L.ptr(m), C.byref(out)))
                return float(out.value)
            out = np.empty(x.n_batch, dtype=np.float64)
            qs = (C.c_int * 1)(what.qubit)
            L.check(lib.bt_sv_expect_product(x.h, 1, qs, L.ptr(m), L.pdouble(out)))
            return float(out[0]) if x.n_batch == 1 else out
        out = np.empty(x.n_batch, dtype=np.float64)
        L.check(lib.bt_sv_expect_matrix2q(x.h, what.qubit, what.target_qubit, L.ptr(L.cmat(what.mat, 4)), L.pdouble(out)))
        return float(out[0]) if x.n_batch == 1 else out
    raise TypeError("expect takes an Op or an operator name")


def correlation(x: State, list_of_operators, qubits_applied: Optional[Sequence[int]] = None):
    """src/func.jl:139-147 (operator string form) and :17-23 (``correlation(state, qubits)`` = Z-parity)."""
    if qubits_applied is None:
        qubits = list(list_of_operators)
        return _expect_product(x, ["Z"] * len(qubits), qubits)
    names = list_of_operators.split(",") if isinstance(list_of_operators, str) else list(list_of_operators)
    return _expect_product(x, names, list(qubits_applied))


def sample(x: State, shots: int, rng=None, uniforms: Optional[np.ndarray] = None) -> np.ndarray:
    """src/ops.jl:46-62 as inverse CDF on uniform draws; 0-based basis indices."""
    if uniforms is None:
        r = _rng(rng)
        uniforms = np.array([r.uniform() for _ in range(shots)], dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    if isinstance(x, CuState) and x.n_batch > 1:
        # batched trajectories: uniforms[t] are trajectory t's draws; one segmented launch set for the whole batch
        if u.ndim == 1:
            u = np.ascontiguousarray(u.reshape(x.n_batch, -1))
        if u.shape[0] != x.n_batch:
            raise ValueError("uniforms must have one row per trajectory")
        out = np.empty(u.shape, dtype=np.int64)
        L.check(x.lib.bt_sv_sample_batched(x.h, L.pdouble(u), u.shape[1], out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out
    out = np.empty(len(u), dtype=np.int64)
    if isinstance(x, CuRho):
        L.check(x.lib.bt_dm_sample(x.h, L.pdouble(u), len(u), out.ctypes.data_as(C.POINTER(C.c_int64))))
    else:
        L.check(x.lib.bt_sv_sample(x.h, L.pdouble(u), len(u), out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


def sample_exact(x: State) -> Tuple[np.ndarray, np.ndarray]:
    """src/ops.jl:98-101 / :129-132: indices with non-zero probability (ascending) and their probabilities."""
    if isinstance(x, CuRho):
        p = np.empty(1 << x.N, dtype=np.float64)
        L.check(x.lib.bt_dm_diag(x.h, L.pdouble(p)))
    else:
        p = prob(x)
    nz = np.nonzero(p)[0]
    return nz.astype(np.int64), p[nz]


def get_probs_from_sample(samples: Sequence[int], N: int):
    """src/ops.jl:76-93."""
    vals, counts = np.unique(np.asarray(samples, dtype=np.int64), return_counts=True)
    return vals, counts / float(len(samples))


def sample_state(x: State, shots: int, rng=None):
    return get_probs_from_sample(sample(x, shots, rng), x.N)


def _sample_to_expectation(bitstr, probv, N: int, qubits: Sequence[int]) -> float:
    """src/func.jl:166-182."""
    a = np.asarray(bitstr, dtype=np.int64)
    par = np.zeros(len(a), dtype=np.int64)
    for q in qubits:
        par ^= (a >> (N - q)) & 1
    p = np.asarray(probv, dtype=np.float64)
    return float(np.sum(np.where(par == 0, p, -p)))


def mag_moments(N: int, bitstr, sample_probs, moment_order: int) -> float:
    """src/func.jl:234-237."""
    a = np.asarray(bitstr, dtype=np.int64)
    ones = np.zeros(len(a), dtype=np.int64)
    for b in range(N):
        ones += (a >> b) & 1
    mag = (N - 2 * ones).astype(np.float64)
    return float(np.sum(mag ** moment_order * np.asarray(sample_probs)))


def cumulants_from_moments(moments: Sequence[float], n: Optional[int] = None):
    """src/func.jl:241-253: cumulant_n = moment_n - sum_{m<n} C(n-1, m-1) cumulant_m moment_{n-m}; all of them when ``n`` is None."""
    from math import comb

    if n is None:
        return [cumulants_from_moments(moments, k) for k in range(1, len(moments) + 1)]
    c: List[float] = []
    for k in range(1, n + 1):
        v = moments[k - 1]
        for m in range(1, k):
            v -= comb(k - 1, m - 1) * c[m - 1] * moments[k - m - 1]
        c.append(v)
    return c[n - 1]


def mag_cumulants(m: "Measurement", cumulant_order: Optional[int] = None, max_moment: int = 12):
    """src/func.jl:255-258: cumulants of the magnetisation from a Measurement's sample distribution."""
    moments = [mag_moments(m.number_of_qubits, m.bitstr, m.sample, i) for i in range(1, max_moment + 1)]
    return cumulants_from_moments(moments, cumulant_order)


class Measurement:
    """src/struct.jl:907-930 (fields the path fills)."""

    def __init__(self, bitstr, sample_probs, expect_v, mag_moments_v, basis, n_exp, name, N):
        self.bitstr, self.sample, self.expect, self.mag_moments = bitstr, sample_probs, expect_v, mag_moments_v
        self.measurement_basis, self.number_of_experiment, self.circuit_name, self.number_of_qubits = basis, n_exp, name, N


def measure(x, number_of_experiment: int = -1, rng=None, label: str = "state to measurement") -> Measurement:
    """``measure(state[, shots])`` src/ops.jl:496-513 and ``measure(sample, N)`` :516-529."""
    if isinstance(x, (CuState, CuRho)):
        N = x.N
        if number_of_experiment == -1:
            bitstr, avg = sample_exact(x)
        else:
            bitstr, avg = sample_state(x, number_of_experiment, rng)
    else:
        raise TypeError("measure(state, shots) takes a device state; use measure_samples for raw samples")
    ex = [_sample_to_expectation(bitstr, avg, N, [i]) for i in range(1, N + 1)]
    mm = [mag_moments(N, bitstr, avg, k) for k in range(1, 13)]
    return Measurement(bitstr, avg, ex, mm, "0", number_of_experiment, label, N)


def measure_samples(samples: Sequence[int], N: int) -> Measurement:
    bitstr, avg = get_probs_from_sample(samples, N)
    ex = [_sample_to_expectation(bitstr, avg, N, [i]) for i in range(1, N + 1)]
    mm = [mag_moments(N, bitstr, avg, k) for k in range(1, 13)]
    return Measurement(bitstr, avg, ex, mm, "0", len(samples), "sample to measurement", N)


# ----------------------------------------------------------------------------------------------------------
# circuit drivers (src/ops.jl:421-445, :599-692, :790-914) -- ops are applied raw (no layout SWAP insertion,
# the device applies any bit pair at the same cost)
# ----------------------------------------------------------------------------------------------------------
def get_N_ops(ops) -> int:
    """src/ops.jl:14-34."""
    m = 0
    for op in ops:
        if isinstance(op, ifOp):
            cur = op.qubit
        elif isinstance(op, OpQC):
            cur = max(op.qubit, op.target_qubit, op.qubit + 2 if op.q == 3 else 0)
        else:
            cur = max(op.qubit, op.target_qubit, op.control)
        m = max(m, cur)
    return m


class Options:
    """src/struct.jl:835-854 (fields the device path honours)."""

    def __init__(self, circuit_name="circuit", measurement_basis="Z", noise=False, density_matrix=False):
        self.circuit_name, self.measurement_basis, self.noise, self.density_matrix = circuit_name, measurement_basis, noise, density_matrix
        self.twirl = False
        self.readout_noise = False
        self.measurement_mitigate = False


class Circuit:
    """src/struct.jl:867-875."""

    def __init__(self, ops, options: Options, N: int):
        self.ops = list(ops)
        self.options = options
        self.N = N
        self.mid_measurement_count = sum(1 for o in self.ops if getattr(o, "type", "") == "🔬")


def ishermitian(op) -> bool:
    """src/linalg.jl:1-7."""
    m = getattr(op, "mat", None)
    return isinstance(m, np.ndarray) and bool(np.array_equal(m, m.conj().T))


def adjoint(x):
    """``adjoint(op)`` / ``adjoint(ops)`` / ``adjoint(circuit)`` -- src/linalg.jl:9-45: Hermitian ops, OpF and ifOp come back unchanged,
    any other Op gets the conjugate-transposed matrix and a "†" appended to its name (removed again by a second adjoint); a list or
    a circuit is reversed.  With it ``apply(adjoint(ops), apply(ops, state))`` returns the state (the encode -> decode round trip the
    full-size tests use)."""
    if isinstance(x, Circuit):
        return Circuit(adjoint(x.ops), x.options, x.N)
    if isinstance(x, (list, tuple)):
        return [adjoint(o) for o in reversed(list(x))]
    if isinstance(x, (OpF, ifOp)) or not isinstance(x, Op) or ishermitian(x):
        return x
    name = x.name[:-1] if x.name.endswith("†") else x.name + "†"
    args = (x.qubit,) if x.q == 1 else (x.qubit, x.target_qubit)
    return Op(name, x.mat.conj().T, *args, control=x.control, noisy=x.noisy, type=x.type)


def isunitary(mat) -> bool:
    """src/linalg.jl:51."""
    m = np.asarray(mat)
    return bool(np.allclose(m.conj().T @ m, np.eye(m.shape[0]), atol=1e-12, rtol=0))


def compile(ops, options: Optional[Options] = None) -> Circuit:
    """src/ops.jl:421-445 without the layout pass (the device needs no SWAP routing)."""
    return Circuit(ops, options or Options(), get_N_ops(ops))


def to_state(circuit: Circuit, rng=None, n_batch: int = 1) -> CuState:
    """src/ops.jl:790-795."""
    s = zero_state(circuit.N, n_batch)
    return apply(circuit.ops, s, noise=circuit.options.noise, rng=rng)


def to_rho(circuit: Circuit) -> CuRho:
    """src/ops.jl:806-844."""
    rho = CuRho(circuit.N)
    apply(circuit.ops, rho, noise=circuit.options.noise)
    return rho


def _final_measurement(state: CuState, options: Options, rng=None) -> CuState:
    """src/ops.jl:856-914 (basis rotation before sampling; readout noise is broken in the reference)."""
    b = options.measurement_basis
    for q in range(1, state.N + 1):
        if b == "Z":
            continue  # Op("ZBasis", I, q) is the identity
        if b == "X":
            m = gate["H"]
        elif b == "Y":
            m = gate["HSP"]
        elif b == "R":
            m = [gate["H"], gate["HSP"], gate["I"]][_rng(rng).randint(3)]
        else:
            raise ValueError("measurement_basis error!")
        _apply_matrix(state, 1, m, q, -1, -2)
    return state


def measure_circuit(circuit: Circuit, number_of_experiment: int, rng=None) -> Measurement:
    """``measure(circuit, shots)`` src/ops.jl:599-646."""
    N = circuit.N
    if circuit.options.noise is False and circuit.mid_measurement_count == 0:
        st = _final_measurement(to_state(circuit, rng), circuit.options, rng)
        samples = sample(st, number_of_experiment, rng)
    else:
        samples = np.empty(number_of_experiment, dtype=np.int64)
        for i in range(number_of_experiment):
            st = _final_measurement(to_state(circuit, rng), circuit.options, rng)
            samples[i] = sample(st, 1, rng)[0]
    m = measure_samples(samples, N)
    m.measurement_basis, m.circuit_name, m.number_of_experiment = circuit.options.measurement_basis, circuit.options.circuit_name, number_of_experiment
    return m


def run(circuit_or_ops, number_of_experiment: int = 1, rng=None, backend="cuda", batch: Optional[BatchDraws] = None):
    """``run(circuit, shots; backend)`` src/ops.jl:659-692: mid-circuit outcomes per shot.

    ``batch``: run all shots as one batched device state (SURVEY 8e trajectories); shot t consumes ``batch.U[t]``."""
    if number_of_experiment <= 0:
        raise ValueError("`number_of_experiment` must be positive.")
    if str(backend).lower() not in ("cuda", "gpu", "statevector", "sv", "state", "state_vector"):
        raise ValueError(f"Unsupported backend `{backend}`.")
    circuit = circuit_or_ops if isinstance(circuit_or_ops, Circuit) else compile(circuit_or_ops)
    nm = circuit.options.noise
    if batch is not None:
        st = zero_state(circuit.N, number_of_experiment)
        _, mids = apply(circuit.ops, st, noise=nm, rng=batch, track_measurements=True)
        if not mids:
            return [[] for _ in range(number_of_experiment)]
        arr = np.stack([np.asarray(m) for m in mids], axis=1)
        return [list(map(int, row)) for row in arr]
    out = []
    for _ in range(number_of_experiment):
        st = zero_state(circuit.N)
        _, mids = apply(circuit.ops, st, noise=nm, rng=rng, track_measurements=True)
        out.append([int(m) for m in mids])
    return out


# ----------------------------------------------------------------------------------------------------------
# remaining members of the path's API surface
# ----------------------------------------------------------------------------------------------------------
def born_measure_Z2(state: CuState, qubit1: int, qubit2: int, rng=None):
    """Two-qubit Born measurement, ``born_measure_Z(N, state, qubit1, qubit2)`` src/hilbert.jl:705-721: outcome index
    1..4 over the projectors [P0P0, P1P0, P0P1, P1P1] on (qubit1, qubit2), drawn with _weighted_sample (:810-819),
    state projected and normalised."""
    P0, P1 = gate["P0"], gate["P1"]
    # expand_multi_op("Pa,Pb",[qubit1,qubit2]): first name acts on qubit1; kron index = 2*b_qubit1 + b_qubit2
    proj = [np.kron(P0, P0), np.kron(P1, P0), np.kron(P0, P1), np.kron(P1, P1)]
    if qubit1 > qubit2:
        # bt_sv_kraus reproduces the CHANNEL quirk for qubit > target; this function has no such quirk, so hand the
        # pair over in ascending order with the projectors re-indexed
        swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=c128)
        proj = [swap @ p @ swap for p in proj]
        qa, qb = qubit2, qubit1
    else:
        qa, qb = qubit1, qubit2
    chosen = _channel_apply(state, proj, 2, qa, qb, rng)
    ind = chosen + 1
    return state, (int(ind[0]) if state.n_batch == 1 else ind)


def sample_bit(state: CuState, shots: int = 1, rng=None) -> List[List[int]]:
    """src/tensor.jl:164-176 for a state vector: sampled bit strings (qubit 1 first)."""
    return [int2bin(int(a), state.N) for a in sample(state, shots, rng)]


def hamiltonian_expect(x: State, terms) -> float:
    """<H> for H = sum_k c_k * (operator string on qubits): the Pauli-sum form of src/vqa.jl:36-67 evaluated term by term
    with the fused Pauli-string reduction (each term reads the state once).  terms: iterable of (coef, "Z,Z", [q1, q2])."""
    tot = 0.0
    for coef, names, qubits in terms:
        tot = tot + coef * correlation(x, names, qubits)
    return tot


def shadow(circuit, number_of_experiment: int, rng=None) -> np.ndarray:
    """``shadow(circuit, number_of_experiment)`` -- src/ops.jl:145-185: classical-shadow estimate of the circuit's density
    matrix.  Per experiment, in the reference's draw order: run the circuit (noise / mid-circuit draws), then for qubit 1..N
    draw a basis ``rand(1:3)`` and rotate with H / HSP / I (``m_list``, :148), then one ``sample(state, 1)`` shot; the
    snapshot is kron_i (3 U_i' |b_i><b_i| U_i - I) and rho is their mean.  The circuit, the N rotations (one fused call)
    and the shot run on the device; the 2^N x 2^N estimate is assembled on the host (dense; the reference returns a sparse
    matrix) -- as in the reference this is a small-N tool.  Throws like the reference when tr(rho) is not 1."""
    r = _rng(rng)
    circ = circuit if isinstance(circuit, Circuit) else compile(list(circuit))
    N = circ.N
    m_list = [gate["H"], gate["HSP"], gate["I"]]
    names = ["Xbasis", "Ybasis", "Zbasis"]
    eye = np.eye(2, dtype=np.complex128)
    rho = np.zeros((1 << N, 1 << N), dtype=np.complex128)
    for _ in range(number_of_experiment):
        state = to_state(circ, rng=r)
        basis = [r.randint(3) for _ in range(N)]
        apply([Op(names[m], m_list[m], q + 1) for q, m in enumerate(basis) if m != 2], state)
        bits = int2bin(int(sample(state, 1, rng=r)[0]), N)
        snap = np.ones((1, 1), dtype=np.complex128)
        for b, m in zip(bits, basis):
            bv = np.zeros((2, 1), dtype=np.complex128)
            bv[b, 0] = 1.0
            um = m_list[m]
            snap = np.kron(snap, 3.0 * (um.conj().T @ bv @ bv.conj().T @ um) - eye)
        rho += snap / number_of_experiment
    if not np.isclose(np.trace(rho), 1.0):
        raise ValueError("la.tr(rho)!≈1")
    return rho


def entanglement_entropy(x, spectrum_bool: bool = False):
    """src/func.jl:299-312 (state) and :323-328 (rho).  State: Schmidt spectrum across the cut between the first N - N/2 and
    the last N/2 qubits (Julia reshapes column-major, so its rows are the LOW N/2 index bits), sum(-s log s) over the squared
    singular values s > 0; with ``spectrum_bool`` also -log.(s).  rho: singular values of ``bipartition_trace(rho)`` (the last
    N/2 qubits; N must be even like the reference's ``Int(N/2)``), returns (entropy, -log.(spec)) like the reference.
    The spectrum comes from the device (one-sided Jacobi iteration, ``bt_sv_schmidt_spectrum`` / ``bt_dm_bipartition_spectrum``);
    a batched state returns one entropy per trajectory."""
    N = x.N
    sweeps = C.c_int()
    if isinstance(x, CuRho):
        if N % 2:
            raise ValueError("InexactError: Int(N/2)")  # src/linalg.jl:152
        spec = np.empty(1 << (N // 2), dtype=np.float64)
        L.check(x.lib.bt_dm_bipartition_spectrum(x.h, N // 2, L.pdouble(spec), C.byref(sweeps)))
        spec = spec[spec > 0]
        return float(np.sum(-spec * np.log(spec))), -np.log(spec)
    part_a = N // 2
    spec = np.empty((x.n_batch, 1 << part_a), dtype=np.float64)
    L.check(x.lib.bt_sv_schmidt_spectrum(x.h, part_a, L.pdouble(spec), C.byref(sweeps)))
    ent, logs = [], []
    for row in spec:
        r = row[row > 0]
        ent.append(float(np.sum(-r * np.log(r))))
        logs.append(-np.log(r))
    if x.n_batch == 1:
        return (ent[0], logs[0]) if spectrum_bool else ent[0]
    return (np.array(ent), logs) if spectrum_bool else np.array(ent)
