"""
oracle/bt_oracle.py -- CPU restatement of BlueTangle.jl's apply -> noise -> measure/sample -> expect path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module, and only as the checker or as the timed CPU baseline -- never as a fallback for the CUDA path.

What it is: a line-by-line numpy/scipy restatement of the reference's *algorithm* (a gate is embedded into a
2^N x 2^N sparse matrix by a chain of ``kron``s, or by a Pauli decomposition for non-adjacent pairs, and then
multiplied into the state; Born probabilities come from reshaped partial traces; Kraus channels are sampled
from tr(K rho_A K')), each function citing the reference file:line it follows (paths relative to
/root/reference/).

Parity pinning: Julia is not installed in the build container, so the reference itself cannot be executed.
The reference's tests hold no stored numeric amplitude vectors; what they hold (test/runtests.jl) are known
answers (:8 bit order, :10-17 exact operator identity, :195-231 deterministic mid-circuit outcomes) and
relational invariants (:19-35, :54-73, :89-133, :136-193).  tests/test_oracle_reference_kats.py re-states
every one of those against this oracle.  Multi-shot ``sample`` goes through StatsBase (un-vendored, version
unpinned, Project.toml:16) => that one function is "parity unpinned"; the contract used here is inverse-CDF
on caller supplied uniforms.

Random numbers: the reference calls Julia's global ``rand()``.  Here every draw is taken from a ``Draws``
object handed in by the caller, in exactly the order the reference consumes them, so the
device path and the oracle can be fed identical uniforms.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

C = np.complex128

# --------------------------------------------------------------------------------------------------
# bit helpers -- src/bit.jl:9-48
# --------------------------------------------------------------------------------------------------


def int2bin(number: int, N: int) -> List[int]:
    """src/bit.jl:9-15 -- MSB first: out[i] (1-based i) = (number >> (N-i)) & 1."""
    out = [0] * N
    for i in range(1, N + 1):
        out[N - i] = (number >> (i - 1)) & 1
    return out


def bin2int(v: Sequence[int]) -> int:
    """src/bit.jl:23-29."""
    r = 0
    for b in v:
        r = (r << 1) + int(b)
    return r


def fock_basis(number: int, N: int) -> List[int]:
    """src/bit.jl:33."""
    return int2bin(number, N)


def mag_basis(number: int, N: int) -> List[int]:
    """src/bit.jl:40."""
    return [1 - 2 * b for b in int2bin(number, N)]


def get_N(x) -> int:
    """src/ops.jl:5-9."""
    n = x.shape[0]
    N = int(round(math.log2(n)))
    assert 1 << N == n
    return N


# --------------------------------------------------------------------------------------------------
# gate tables -- src/gates.jl:22-59 (constants), :369-453 (parametrised)
# --------------------------------------------------------------------------------------------------


def _round_sig(z: complex, sig: int = 10) -> complex:
    """Julia ``round(x, sigdigits=10)`` applied to re and im separately (src/gates.jl:32-33,47-48)."""

    def r(x: float) -> float:
        if x == 0.0 or not math.isfinite(x):
            return x
        d = sig - int(math.floor(math.log10(abs(x)))) - 1
        return round(x, d)

    return complex(r(z.real), r(z.imag))


def _rs(m):
    m = np.array(m, dtype=C)
    out = np.empty_like(m)
    for i in np.ndindex(m.shape):
        out[i] = _round_sig(complex(m[i]))
    return out


_s2 = 1 / math.sqrt(2)
_im = 1j

GATE = {
    "I": np.array([[1, 0], [0, 1]], dtype=C),
    "X": np.array([[0, 1], [1, 0]], dtype=C),
    "SX": 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=C),
    "XSQRT": 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=C),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=C),
    "Z": np.array([[1, 0], [0, -1]], dtype=C),
    "H": _s2 * np.array([[1, 1], [1, -1]], dtype=C),
    "S": np.array([[1, 0], [0, 1j]], dtype=C),
    "SD": np.array([[1, 0], [0, -1j]], dtype=C),
    "T": _rs([[1, 0], [0, np.exp(1j * math.pi / 4)]]),
    "TD": _rs([[1, 0], [0, np.exp(-1j * math.pi / 4)]]),
    "HSP": _s2 * np.array([[1, -1j], [1, 1j]], dtype=C),
    "HY": _s2 * np.array([[1, 1j], [1, -1j]], dtype=C),
    "H2": np.array([[0.5, 0.5, 0.5, 0.5], [0.5, -0.5, 0.5, -0.5], [0.5, 0.5, -0.5, -0.5], [0.5, -0.5, -0.5, 0.5]], dtype=C),
    "P0": np.array([[1, 0], [0, 0]], dtype=C),
    "P1": np.array([[0, 0], [0, 1]], dtype=C),
    "SP": np.array([[0, 1], [0, 0]], dtype=C),
    "SM": np.array([[0, 0], [1, 0]], dtype=C),
    "CX": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=C),
    "CNOT": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=C),
    "CY": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]], dtype=C),
    "CZ": np.diag([1, 1, 1, -1]).astype(C),
    "CS": np.diag([1, 1, 1, 1j]).astype(C),
    "CT": _rs(np.diag([1, 1, 1, np.exp(1j * math.pi / 4)])),
    "CTD": _rs(np.diag([1, 1, 1, np.exp(-1j * math.pi / 4)])),
    "CSD": np.diag([1, 1, 1, -1j]).astype(C),
    "CI": np.diag([1, 1, 1, 1]).astype(C),
    "CH": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, _s2, _s2], [0, 0, _s2, -_s2]], dtype=C),
    "SWAP": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=C),
    "ISWAP": np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=C),
    "FSWAP": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, -1]], dtype=C),
    "SYC": np.array([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, np.exp(-1j * math.pi / 6)]], dtype=C),
    "ECR": _s2 * np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]], dtype=C),
}
_ccx = np.eye(8, dtype=C)
_ccx[6:8, 6:8] = [[0, 1], [1, 0]]
GATE["CCX"] = _ccx
_ccz = np.eye(8, dtype=C)
_ccz[7, 7] = -1
GATE["CCZ"] = _ccz

GATES_WITH_PHASE = ["P", "RX", "RY", "RZ", "U1", "U2", "U3", "CP", "GIVENS", "FSIM", "SWAPA", "RXX", "RYY", "RZZ", "RXY"]
ONE_QUBIT_GATES = ["I", "X", "Y", "Z", "SX", "XSQRT", "H", "T", "S", "SD", "P", "U2", "U3"]
TWO_QUBIT_GATES = ["CX", "CNOT", "CY", "CZ", "CP", "RXX", "RYY", "RZZ", "RXY", "GIVENS", "FSIM", "SWAP", "ISWAP", "FSWAP", "SYC", "ECR"]


def _P(l):
    return np.array([[1, 0], [0, np.exp(1j * l)]], dtype=C)


def _RX(t):
    return np.array([[math.cos(t / 2), -1j * math.sin(t / 2)], [-1j * math.sin(t / 2), math.cos(t / 2)]], dtype=C)


def _RY(t):
    return np.array([[math.cos(t / 2), -math.sin(t / 2)], [math.sin(t / 2), math.cos(t / 2)]], dtype=C)


def _RZ(t):
    return np.array([[np.exp(-1j * t / 2), 0], [0, np.exp(1j * t / 2)]], dtype=C)


def _U2(phi, lam):
    return _s2 * np.array([[1, -np.exp(1j * lam)], [np.exp(1j * phi), np.exp(1j * (phi + lam))]], dtype=C)


def _U3(t, phi, lam):
    return np.array(
        [[math.cos(t / 2), -np.exp(1j * lam) * math.sin(t / 2)], [np.exp(1j * phi) * math.sin(t / 2), np.exp(1j * (phi + lam)) * math.cos(t / 2)]],
        dtype=C,
    )


def _CP(l):
    return np.diag([1, 1, 1, np.exp(1j * l)]).astype(C)


def _GIVENS(t):
    return np.array([[1, 0, 0, 0], [0, math.cos(t), -math.sin(t), 0], [0, math.sin(t), math.cos(t), 0], [0, 0, 0, 1]], dtype=C)


def _FSIM(t, phi):
    return np.array(
        [[1, 0, 0, 0], [0, math.cos(t), -1j * math.sin(t), 0], [0, -1j * math.sin(t), math.cos(t), 0], [0, 0, 0, np.exp(1j * phi)]], dtype=C
    )


def _SWAPA(a):
    e = np.exp(1j * math.pi * a)
    return 0.5 * np.array([[2, 0, 0, 0], [0, 1 + e, 1 - e, 0], [0, 1 - e, 1 + e, 0], [0, 0, 0, 2]], dtype=C)


def _RXX(p):  # src/gates.jl:401 -- cos(phi), NOT cos(phi/2)
    c, s = math.cos(p), -1j * math.sin(p)
    return np.array([[c, 0, 0, s], [0, c, s, 0], [0, s, c, 0], [s, 0, 0, c]], dtype=C)


def _RYY(p):  # src/gates.jl:403 -- +i sin on the corners
    c, s = math.cos(p), 1j * math.sin(p)
    return np.array([[c, 0, 0, s], [0, c, -s, 0], [0, -s, c, 0], [s, 0, 0, c]], dtype=C)


def _RZZ(p):
    a, b = np.exp(-1j * p / 2), np.exp(1j * p / 2)
    return np.diag([a, b, b, a]).astype(C)


def _RXY(p):
    c, s = math.cos(p), -1j * math.sin(p)
    return np.array([[1, 0, 0, 0], [0, c, s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=C)


_PARAM = {
    "P": _P, "RX": _RX, "RY": _RY, "RZ": _RZ, "U1": _P, "U2": _U2, "U3": _U3, "CP": _CP, "GIVENS": _GIVENS,
    "SWAPA": _SWAPA, "FSIM": _FSIM, "RXX": _RXX, "RYY": _RYY, "RZZ": _RZZ, "RXY": _RXY,
}


def _eval_julia_number(expr: str) -> float:
    """Tiny stand-in for ``eval(Meta.parse(..))`` (src/gates.jl:348) on numeric literals such as
    ``.1pi``, ``0.5π``, ``-pi/4``, ``2*pi/3``.  Juxtaposition ``<number>pi`` means multiplication."""
    e = expr.strip().replace("π", "pi")
    e = re.sub(r"(\d|\.)\s*pi", r"\1*pi", e)
    if not re.fullmatch(r"[0-9eE\.\+\-\*/\(\) pi]*", e):
        raise ValueError(f"unsupported gate argument {expr!r}")
    return float(eval(e, {"__builtins__": {}}, {"pi": math.pi}))


def _clean_name(name: str) -> str:
    return name.split("(")[0].upper()


def is_measurement(name: str) -> bool:
    """src/gates.jl:331-335."""
    return name.upper() in ("MZ", "M(Z)", "MX", "M(X)", "MY", "M(Y)", "MR", "M(R)")


def gates(op_name: str) -> np.ndarray:
    """src/gates.jl:369-453."""
    clean = _clean_name(op_name)
    up = op_name.upper()
    if clean in _PARAM:
        args = op_name.split("(", 1)[1].rsplit(")", 1)[0].split(",")
        return _PARAM[clean](*[_eval_julia_number(a) for a in args])
    if up in ("M(Z)", "MZ", "M(R)", "MR", "RES"):
        return GATE["I"]
    if up in ("M(X)", "MX"):
        return GATE["H"]
    if up in ("M(Y)", "MY"):
        return GATE["HSP"]
    if up in GATE:
        return GATE[up]
    raise KeyError(f"Gate {op_name} not found")


# --------------------------------------------------------------------------------------------------
# Kraus sets -- src/noise.jl:52-131
# --------------------------------------------------------------------------------------------------


def noise_model(model: str, p: float, two_qubit: bool = False) -> List[np.ndarray]:
    """src/noise.jl:52-131.  Quirk kept: ``phase_flip`` yields Y (String compared with Symbol, :80)."""
    I, X, Y, Z = GATE["I"], GATE["X"], GATE["Y"], GATE["Z"]
    model = model.lower()
    if model == "amplitude_damping":
        ops = [np.array([[1, 0], [0, math.sqrt(1 - p)]], dtype=C), np.array([[0, math.sqrt(p)], [0, 0]], dtype=C)]
    elif model == "phase_damping":
        ops = [np.array([[1, 0], [0, math.sqrt(1 - p)]], dtype=C), np.array([[0, 0], [0, math.sqrt(p)]], dtype=C)]
    elif model in ("phase_flip", "bit_flip", "bit_phase_flip"):
        ops = [math.sqrt(1 - p) * I, math.sqrt(p) * (X if model == "bit_flip" else Y)]
    elif model == "depolarizing_amp":
        ops = [math.sqrt(1 - 3 * p / 4) * I, math.sqrt(p / 4) * X, math.sqrt(p / 4) * Y, math.sqrt(p / 4) * Z]
    elif model == "depolarizing":
        ops = [math.sqrt(1 - p) * I, math.sqrt(p / 3) * X, math.sqrt(p / 3) * Y, math.sqrt(p / 3) * Z]
    elif model == "rot_z":
        ops = [_RZ(p)]
    elif model == "rot_y":
        ops = [_RY(p)]
    elif model == "rot_x":
        ops = [_RX(p)]
    elif model == "rot_p":
        ops = [_P(p)]
    elif model == "rot_xyz":
        ops = [_RX(p) / math.sqrt(3), _RY(p) / math.sqrt(3), _RZ(p) / math.sqrt(3)]
    elif is_measurement(model):
        ops = [GATE["P0"], GATE["P1"]]
    else:
        raise ValueError("Unknown quantum error model")
    if not two_qubit:
        return ops
    return [np.kron(Ki, Kj) for Ki in ops for Kj in ops]  # src/noise.jl:126


def is_valid_quantum_channel(kraus: Sequence[np.ndarray]) -> bool:
    """src/struct.jl:291-309."""
    n = kraus[0].shape[0]
    s = sum(K.conj().T @ K for K in kraus)
    tp = np.allclose(s, np.eye(n), rtol=math.sqrt(np.finfo(float).eps), atol=0)
    choi = sum(np.outer(K.reshape(-1, order="F"), K.reshape(-1, order="F").conj()) for K in kraus)
    ev = np.linalg.eigvalsh(choi)
    cp = np.allclose(choi, choi.conj().T) and bool(np.all(np.round(ev, 10) >= 0))
    return bool(tp and cp)


# --------------------------------------------------------------------------------------------------
# operator embedding -- src/hilbert.jl:18-159, :191-237; src/decompose.jl:5-50,71-100
# --------------------------------------------------------------------------------------------------

_ID2 = sp.identity(2, dtype=C, format="csc")


def _foldl_kron(mats):
    out = mats[0]
    for m in mats[1:]:
        out = sp.kron(out, m, format="csc")
    return out


def _check_swap_invariant(m: np.ndarray) -> bool:
    """src/hilbert.jl:191-204 (exact == 0 test)."""
    S = GATE["SWAP"]
    return bool(np.all((m @ S - S @ m) == 0))


def _swap_control_target(m: np.ndarray) -> np.ndarray:
    """src/hilbert.jl:215-237."""
    if _check_swap_invariant(m):
        return m
    perm = [0, 2, 1, 3]
    return m[np.ix_(perm, perm)]


_PAULIS = [GATE["I"], GATE["X"], GATE["Y"], GATE["Z"]]


def pauli_decomposition(A: np.ndarray) -> np.ndarray:
    """src/decompose.jl:71-80 for n=2: coefficient list in Iterators.product order (first index fastest)."""
    A = np.asarray(A.todense()) if sp.issparse(A) else np.asarray(A)
    coeffs = np.zeros(16, dtype=C)
    for j in range(4):
        for i in range(4):
            Pm = np.kron(_PAULIS[i], _PAULIS[j])
            coeffs[i + 4 * j] = np.trace(A @ Pm) / 4
    return coeffs


def pauli_reconstruction(coeffs: np.ndarray, qubit: int, distance: int):
    """src/decompose.jl:96-100 + pauli_decomposition_tensor :5-50 (qubit != 0 branch):
    sum_ij c_ij  I^(qubit-1) (x) s_i (x) I^(distance-1) (x) s_j."""
    sP = [sp.csc_matrix(p) for p in _PAULIS]
    total = None
    for j in range(4):
        for i in range(4):
            c = coeffs[i + 4 * j]
            product = [sP[i], sP[j]]
            for _ in range(qubit - 1):
                product.insert(0, sP[0])
            for _ in range(distance - 1):
                product.insert(qubit, sP[0])  # Julia insert!(product, qubit+1, I) (1-based)
            term = c * _foldl_kron(product)
            total = term if total is None else total + term
    return total


def CCZX(sym: str, N: int, i: int, j: int, k: int):
    """src/hilbert.jl:73-103."""
    for q in (i, j, k):
        if q < 1 or q > N:
            raise ValueError("Qubit indices must be within the range 1 to N")
    i, j, k = N - i, N - j, N - k
    dim = 1 << N
    a = np.arange(dim)
    if sym == "CCZ":
        d = np.where(((a >> i) & (a >> j) & (a >> k) & 1) == 1, -1.0, 1.0)
        return sp.csc_matrix((d.astype(C), (a, a)), shape=(dim, dim))
    if sym == "CCX":
        cond = ((a >> i) & (a >> j) & 1) == 1
        b = np.where(cond, a ^ (1 << k), a)
        return sp.csc_matrix((np.ones(dim, dtype=C), (a, b)), shape=(dim, dim))
    raise ValueError("Unsupported operation. Use 'CCZ' or 'CCX'.")


def hilbert1(N: int, mat: np.ndarray, qubit: int, control: int = -2):
    """src/hilbert.jl:143-159."""
    if N < qubit or N < control:
        raise ValueError("N must be larger than qubit")
    if mat.shape[0] > 2:
        raise ValueError("only 2-qubit operations are supported")
    m = sp.csc_matrix(mat.astype(C))
    if control == -2:
        return _foldl_kron([m if x == qubit else _ID2 for x in range(1, N + 1)])
    P1, P0 = sp.csc_matrix(GATE["P1"]), sp.csc_matrix(GATE["P0"])
    lst = _foldl_kron([m if x == qubit else (P1 if x == control else _ID2) for x in range(1, N + 1)])
    return lst + _foldl_kron([P0 if x == control else _ID2 for x in range(1, N + 1)])


def hilbert2(N: int, mat: np.ndarray, qubit: int, target: int, control: int = -2):
    """src/hilbert.jl:18-70."""
    distance = abs(qubit - target)
    if N < qubit or N < target or N < control:
        raise ValueError("N must be larger than qubits")
    if mat.shape[0] > 4:
        raise ValueError("only 2-qubit operations are supported")
    final = sp.csc_matrix((_swap_control_target(mat) if qubit > target else mat).astype(C))
    if control == -2:
        if distance == 1:
            return _foldl_kron([final if x == qubit else _ID2 for x in range(1, N + 1) if x != target])
        coeffs = pauli_decomposition(final)
        out = pauli_reconstruction(coeffs, min(qubit, target), distance)
        for _ in range(N - max(qubit, target)):
            out = sp.kron(out, _ID2, format="csc")
        return out
    if distance == 1:
        P1, P0 = sp.csc_matrix(GATE["P1"]), sp.csc_matrix(GATE["P0"])
        lst = _foldl_kron([final if x == qubit else (P1 if x == control else _ID2) for x in range(1, N + 1) if x != target])
        return lst + _foldl_kron([P0 if x == control else _ID2 for x in range(1, N + 1)])
    if np.array_equal(mat, GATE["CX"]):
        return CCZX("CCX", N, qubit, control, target)
    if np.array_equal(mat, GATE["CZ"]):
        return CCZX("CCZ", N, qubit, control, target)
    raise ValueError("Unsupported operation. Use 'CCZ' or 'CCX'.")


def hilbert3(N: int, mat: np.ndarray, first_qubit: int):
    """src/hilbert.jl:106-128."""
    if N < first_qubit + 2:
        raise ValueError("N must be larger than all three qubits")
    if mat.shape[0] != 8:
        raise ValueError("only 3-qubit operations are supported")
    e_ops = [_ID2] * N
    e_ops = e_ops[: first_qubit - 1] + [sp.csc_matrix(mat.astype(C))] + e_ops[first_qubit - 1 :]
    for _ in range(3):
        e_ops.pop(first_qubit)
    return _foldl_kron(e_ops)


# --------------------------------------------------------------------------------------------------
# partial traces -- src/linalg.jl:83-230
# --------------------------------------------------------------------------------------------------


def partial_trace_1(state: np.ndarray, q: int) -> np.ndarray:
    """src/linalg.jl:167-192.  Julia reshape is column-major: (left, 2, right) with left = 2^(N-q)."""
    N = get_N(state)
    left = 1 << (N - q)
    right = 1 << (q - 1)
    t = state.reshape((left, 2, right), order="F")
    rho = np.zeros((2, 2), dtype=C)
    # sum_{i,j} v v'  ==  einsum over the outer indices (same arithmetic, vectorised)
    rho[:, :] = np.einsum("iaj,ibj->ab", t, t.conj())
    return rho


def partial_trace_2adj(state: np.ndarray, k1: int, k2: int) -> np.ndarray:
    """src/linalg.jl:198-230.  Index of the 4x4 = 2*b_min + b_max."""
    N = get_N(state)
    if abs(k1 - k2) > 1:
        raise ValueError("must be local")
    qi1 = N - min(k1, k2)
    left = 1 << (qi1 - 1)
    right = 1 << (N - qi1 - 1)
    t = state.reshape((left, 4, right), order="F")
    return np.einsum("iaj,ibj->ab", t, t.conj()).astype(C)


def partial_trace_general(state: np.ndarray, keep: Sequence[int]) -> np.ndarray:
    """src/linalg.jl:83-140 via ``state*state'``: subsystem k <-> qubit k (MSB first); kept qubits stay in
    ascending label order.  Restated with a tensor reshape instead of the O(4^N) dense loop (same sums)."""
    N = get_N(state)
    keep_sorted = sorted(set(keep))
    t = state.reshape([2] * N)  # C-order: axis k-1 <-> qubit k
    other = [a for a in range(N) if (a + 1) not in keep_sorted]
    perm = [k - 1 for k in keep_sorted] + other
    m = np.transpose(t, perm).reshape(1 << len(keep_sorted), -1)
    return (m @ m.conj().T).astype(C)


def partial_trace_rho(rho, dims: Sequence[int], trace_out: Sequence[int]) -> np.ndarray:
    """src/linalg.jl:88-140 (general dims; trace_out 1-based subsystem labels)."""
    rho = np.asarray(rho.todense()) if sp.issparse(rho) else np.asarray(rho)
    n = len(dims)
    t = rho.reshape(list(dims) + list(dims))
    keep = [i for i in range(n) if (i + 1) not in trace_out]
    tr = [i for i in range(n) if (i + 1) in trace_out]
    t = np.transpose(t, keep + tr + [n + i for i in keep] + [n + i for i in tr])
    dk = int(np.prod([dims[i] for i in keep])) if keep else 1
    dt = int(np.prod([dims[i] for i in tr])) if tr else 1
    t = t.reshape(dk, dt, dk, dt)
    return np.einsum("akbk->ab", t).astype(C)


# --------------------------------------------------------------------------------------------------
# op descriptors -- src/struct.jl:365-474 (Op), :182-249 (OpQC), :666-696 (ifOp), :90-173 (channels)
# --------------------------------------------------------------------------------------------------


class Draws:
    """Uniform-draw source standing in for Julia's global RNG.  ``uniform()`` <-> ``rand()``;
    ``randint(n)`` <-> ``rand(1:n)`` / ``rand([..n items..])`` (0-based result)."""

    def __init__(self, seed_or_gen=0):
        self.g = seed_or_gen if isinstance(seed_or_gen, np.random.Generator) else np.random.Generator(np.random.PCG64(seed_or_gen))
        self.log: List[float] = []

    def uniform(self) -> float:
        u = float(self.g.random())
        self.log.append(u)
        return u

    def randint(self, n: int) -> int:
        return int(self.g.integers(0, n))


class ListDraws(Draws):
    """Draws replayed from an explicit list (used to feed the device path's recorded uniforms back)."""

    def __init__(self, us: Sequence[float], ints: Sequence[int] = ()):
        self.us = list(us)
        self.ints = list(ints)
        self.log = []

    def uniform(self) -> float:
        u = self.us.pop(0)
        self.log.append(u)
        return u

    def randint(self, n: int) -> int:
        return self.ints.pop(0)


@dataclass
class Op:
    """src/struct.jl:365-453."""

    name: str
    qubit: int
    target_qubit: int = -1
    control: int = -2
    mat: Optional[np.ndarray] = None
    noisy: bool = True
    type: str = ""
    q: int = field(init=False, default=1)

    def __post_init__(self):
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
        if self.mat is None:
            self.mat = gates(self.name)
        self.mat = np.asarray(self.mat, dtype=C)
        if self.mat.shape != (1 << self.q, 1 << self.q):
            raise ValueError(f"size of matrix {self.mat.shape} not compatible with {self.q}-qubit operation")
        self.ismeasure = self.q == 1 and is_measurement(self.name)
        if self.ismeasure:
            if self.control != -2:
                raise ValueError("measurement and control operations are incompatible.")
            self.type = "🔬"
            self.noisy = False
        elif not self.type:
            self.type = "phase" if _clean_name(self.name) in GATES_WITH_PHASE else "op"

    def expand(self, N: int, draws: Optional[Draws] = None):
        """src/struct.jl:476-489."""
        if self.target_qubit == -1:
            if self.ismeasure:
                return measurement_hilbert(N, self.name, self.qubit, draws)
            return hilbert1(N, self.mat, self.qubit, self.control)
        return hilbert2(N, self.mat, self.qubit, self.target_qubit, self.control)


def Op3(name: str, qubit: int, control_qubit: int, target_qubit: int) -> Op:
    """src/struct.jl:436-449."""
    m = {"CCZ": "CZ", "CCX": "CX", "CCY": "CY", "CSWAP": "SWAP"}
    if name not in m:
        raise ValueError("Unsupported three-qubit operation")
    return Op(m[name], qubit, target_qubit, control=control_qubit)


@dataclass
class OpQC:
    """src/struct.jl:182-249."""

    name: str
    kraus: List[np.ndarray]
    qubit: int
    target_qubit: int = -1
    type: str = ""

    def __post_init__(self):
        self.kraus = [np.asarray(k, dtype=C) for k in self.kraus]
        if not is_valid_quantum_channel(self.kraus):
            raise ValueError("not valid kraus operators: not CPTP!")
        self.q = int(round(math.log2(self.kraus[0].shape[0])))
        self.name = self.name.lower()
        self.control = -2
        self.noisy = False
        if self.q == 1 and self.target_qubit > 0:
            raise ValueError("for 1-qubit quantum channel, target_qubit should be -1")
        if self.q == 2 and self.target_qubit < 0:
            raise ValueError("for 2-qubit quantum channel, you must select the target_qubit")
        if self.q == 3:
            self.target_qubit = -1
        if self.q > 3:
            raise ValueError("Noise models are available only for up to 3 qubits!")

    @staticmethod
    def model(model: str, p: float, qubit: int, target_qubit: int = -1) -> "OpQC":
        """src/struct.jl:242-247."""
        two = target_qubit > 0
        return OpQC(model.lower(), noise_model(model, p, two_qubit=two), qubit, target_qubit, type=str(p))

    def prob(self, state: np.ndarray) -> List[float]:
        if self.q == 1:
            return calc_prob(state, self.kraus, self.qubit)
        if self.q == 2:
            return calc_prob(state, self.kraus, self.qubit, self.target_qubit)
        return calc_prob3(state, self.kraus, self.qubit)

    def apply(self, x, draws: Optional[Draws] = None):
        if x.ndim == 2:
            return channel_apply_rho(x, self.kraus, self.qubit, self.target_qubit)
        if self.q == 3:
            return channel_apply3(x, self.kraus, self.qubit, draws)
        return channel_apply(x, self.kraus, self.qubit, self.target_qubit, draws)


def RES(qubit: int) -> OpQC:
    """src/struct.jl:421-423: Op("RES",q) is an OpQC of amplitude damping with gamma = 1."""
    return OpQC("RES", [np.array([[1, 0], [0, 0]], dtype=C), np.array([[0, 1], [0, 0]], dtype=C)], qubit)


@dataclass
class QuantumChannel:
    """src/struct.jl:90-140."""

    q: int
    name: str
    p: float
    kraus: List[np.ndarray] = field(default=None)

    def __post_init__(self):
        self.name = self.name.lower()
        if self.kraus is None:
            self.kraus = noise_model(self.name, self.p, two_qubit=(self.q == 2))
        if not is_valid_quantum_channel(self.kraus):
            raise ValueError("not valid kraus operators: not CPTP!")

    def apply(self, x, qubit: int, target: int = -1, draws: Optional[Draws] = None):
        if x.ndim == 2:
            return channel_apply_rho(x, self.kraus, qubit, target)
        return channel_apply(x, self.kraus, qubit, target, draws)

    def prob(self, state, qubit: int, target: int = -1):
        return calc_prob(state, self.kraus, qubit, target)


@dataclass
class NoiseModel:
    """src/struct.jl:147-173."""

    q1: object
    q2: object

    @staticmethod
    def model(model: str, p: float) -> "NoiseModel":
        return NoiseModel(QuantumChannel(1, model, p), QuantumChannel(2, model, p))


@dataclass
class ifOp:
    """src/struct.jl:666-696."""

    name: str
    qubit: int
    if0: List[Op] = field(default_factory=list)
    if1: List[Op] = field(default_factory=list)

    def __post_init__(self):
        if not is_measurement(self.name):
            raise ValueError("select MX or MY or MZ or MR basis.")
        self.q = 1
        self.type = "🔬"
        self.target_qubit = -1
        self.control = -2


# --------------------------------------------------------------------------------------------------
# measurement / noise arithmetic -- src/hilbert.jl:322-364, :669-819; src/struct.jl:9-76, :554-594
# --------------------------------------------------------------------------------------------------


class OpF:
    """src/struct.jl:703-744: f(state) -> state; a full-register matrix (mat * state); or an op list applied in order."""

    def __init__(self, name: str, data):
        self.q = 1
        self.name = name
        self.type = ""
        self.data = data
        if callable(data):
            self.apply = lambda state, **kw: data(state, **kw)
        elif isinstance(data, (list, tuple)):
            def _run(state, **kw):
                for o in data:  # src/struct.jl:733-735: state = o * state (all.jl:41-49: apply without noise)
                    state = apply(state, o)
                return state
            self.apply = _run
        else:
            self.apply = lambda state, **kw: data @ state


def _normalize(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v)


def weighted_sample(probs: Sequence[float], draws: Draws) -> Optional[int]:
    """src/hilbert.jl:810-819: first i (0-based here) with rval <= cumsum(probs)[i]; None if none."""
    rval = draws.uniform()
    for i, cw in enumerate(np.cumsum(np.asarray(probs, dtype=float))):
        if rval <= cw:
            return i
    return None


def calc_prob(state: np.ndarray, kraus: Sequence[np.ndarray], qubit: int, target: int = -1) -> List[float]:
    """src/struct.jl:9-29."""
    if target == -1:
        pA = partial_trace_1(state, qubit)
    elif abs(qubit - target) == 1:
        pA = partial_trace_2adj(state, qubit, target)
    else:
        pA = partial_trace_general(state, [qubit, target])
    return [float(np.real(np.trace(K @ pA @ K.conj().T))) for K in kraus]


def calc_prob3(state: np.ndarray, kraus: Sequence[np.ndarray], first: int) -> List[float]:
    """src/struct.jl:44-49."""
    pA = partial_trace_general(state, [first, first + 1, first + 2])
    return [float(np.real(np.trace(K @ pA @ K.conj().T))) for K in kraus]


def channel_apply(state: np.ndarray, kraus: Sequence[np.ndarray], qubit: int, target: int, draws: Draws) -> np.ndarray:
    """src/struct.jl:31-41 (state-vector trajectory step)."""
    N = get_N(state)
    ind = weighted_sample(calc_prob(state, kraus, qubit, target), draws)
    if ind is None:
        raise RuntimeError("_weighted_sample returned nothing (SURVEY App. A.5 #4)")
    E = hilbert1(N, kraus[ind], qubit) if target == -1 else hilbert2(N, kraus[ind], qubit, target)
    return _normalize(E @ state)


def channel_apply3(state: np.ndarray, kraus: Sequence[np.ndarray], first: int, draws: Draws) -> np.ndarray:
    """src/struct.jl:51-55."""
    N = get_N(state)
    ind = weighted_sample(calc_prob3(state, kraus, first), draws)
    return _normalize(hilbert3(N, kraus[ind], first) @ state)


def channel_apply_rho(rho: np.ndarray, kraus: Sequence[np.ndarray], qubit: int, target: int = -1) -> np.ndarray:
    """src/struct.jl:58-76."""
    N = get_N(rho)
    new = np.zeros_like(rho)
    for K in kraus:
        E = hilbert1(N, K, qubit) if target == -1 else hilbert2(N, K, qubit, target)
        new = new + E @ rho @ E.conj().T
    return new


def measurement_mat(name: str) -> np.ndarray:
    """src/struct.jl:554-563."""
    u = name.upper()
    if u in ("M(Z)", "MZ"):
        return GATE["I"]
    if u in ("M(X)", "MX"):
        return GATE["H"]
    if u in ("M(Y)", "MY"):
        return GATE["HSP"]
    raise ValueError(name)


def resolve_measurement_name(name: str, draws: Optional[Draws]) -> str:
    """src/struct.jl:565-571 ("MR" -> one discrete draw over ["MX","MY","MZ"])."""
    u = name.upper()
    if u in ("MR", "M(R)"):
        return ["MX", "MY", "MZ"][draws.randint(3)]
    return u


def measurement_hilbert(N: int, name: str, qubit: int, draws: Optional[Draws]):
    """src/struct.jl:573-576."""
    return hilbert1(N, measurement_mat(resolve_measurement_name(name, draws)), qubit)


def born_measure_Z(N: int, state: np.ndarray, qubit: int, draws: Draws) -> Tuple[np.ndarray, int]:
    """src/hilbert.jl:682-696."""
    prob0 = float(np.real(partial_trace_1(state, qubit))[0, 0])
    ind = 0 if draws.uniform() < prob0 else 1
    P = GATE["P0"] if ind == 0 else GATE["P1"]
    return _normalize(hilbert1(N, P, qubit) @ state), ind


def born_measure(state: np.ndarray, o, draws: Draws) -> Tuple[np.ndarray, int]:
    """src/hilbert.jl:669-679."""
    N = get_N(state)
    rot = o.expand(N, draws)
    state = rot @ state
    state, ind = born_measure_Z(N, state, o.qubit, draws)
    state = rot.conj().T @ state
    return state, ind


def reset_Z(state: np.ndarray, qubit: int, draws: Draws) -> Tuple[np.ndarray, int]:
    """src/hilbert.jl:752-759."""
    N = get_N(state)
    state, ind = born_measure_Z(N, state, qubit, draws)
    if ind == 1:
        state = hilbert1(N, GATE["X"], qubit) @ state
    return state, ind


def born_measure_Z_rho(N: int, rho: np.ndarray, qubit: int) -> np.ndarray:
    """src/hilbert.jl:784-796 (non-selective dephasing)."""
    new = np.zeros_like(rho)
    for b in (GATE["P0"], GATE["P1"]):
        eb = hilbert1(N, b, qubit)
        new = new + eb @ rho @ eb.conj().T
    return new


def ifop_apply(state: np.ndarray, op: ifOp, noise, draws: Draws) -> Tuple[np.ndarray, int]:
    """src/struct.jl:578-594."""
    N = get_N(state)
    rname = resolve_measurement_name(op.name, draws)
    rot = hilbert1(N, measurement_mat(rname), op.qubit)
    state = rot @ state
    state, ind = born_measure_Z(N, state, op.qubit, draws)
    state = rot.conj().T @ state
    for o in op.if0 if ind == 0 else op.if1:
        state = apply(state, o, noise=noise, draws=draws)
    return state, ind


def apply_noise(state: np.ndarray, op, noise: NoiseModel, draws: Optional[Draws]) -> np.ndarray:
    """src/hilbert.jl:322-364 (same branch structure for state vectors and density matrices)."""
    if not (hasattr(op, "noisy") and op.noisy is True):
        return state
    if op.q == 1:
        if op.control == -2:
            return noise.q1.apply(state, op.qubit, draws=draws)
        return noise.q2.apply(state, op.control, op.qubit, draws=draws)
    if op.q == 2:
        return noise.q2.apply(state, op.qubit, op.target_qubit, draws=draws)
    return state


def apply(state: np.ndarray, op, noise=False, draws: Optional[Draws] = None, track_measurements: bool = False):
    """src/hilbert.jl:469-515 (state vector) and :639-666 (density matrix, dispatch on ndim)."""
    if isinstance(op, (list, tuple)):
        return apply_ops(state, op, noise=noise, draws=draws, track_measurements=track_measurements)
    if state.ndim == 2:
        return _apply_rho(state, op, noise)
    N = get_N(state)
    mid: List[int] = []
    if isinstance(op, OpF):  # src/hilbert.jl:486-487
        state = op.apply(state)
    elif isinstance(op, OpQC):
        if op.name.upper() in ("RES", "RESET"):
            state, _ = reset_Z(state, op.qubit, draws)
        else:
            state = op.apply(state, draws)
    elif op.type == "🔬":
        if isinstance(op, ifOp):
            state, ind = ifop_apply(state, op, noise, draws)
        else:
            state, ind = born_measure(state, op, draws)
        if track_measurements:
            mid.append(ind)
    else:
        state = op.expand(N) @ state
    if isinstance(noise, NoiseModel):
        state = apply_noise(state, op, noise, draws)
    return (state, mid) if track_measurements else state


def apply_ops(state: np.ndarray, ops, noise=False, draws: Optional[Draws] = None, track_measurements: bool = False):
    """src/hilbert.jl:517-553."""
    mid: List[int] = []
    for o in ops:
        if track_measurements and state.ndim == 1:
            state, new = apply(state, o, noise=noise, draws=draws, track_measurements=True)
            mid.extend(new)
        else:
            state = apply(state, o, noise=noise, draws=draws)
    return (state, mid) if track_measurements else state


def _apply_rho(rho: np.ndarray, op, noise=False) -> np.ndarray:
    """src/hilbert.jl:639-666."""
    N = get_N(rho)
    if isinstance(op, OpQC):
        rho = op.apply(rho)
    elif op.type == "🔬":
        raise RuntimeError("fix this:")  # src/hilbert.jl:772 / src/struct.jl:616 -- unsupported in the reference
    else:
        e = op.expand(N)
        rho = e @ rho @ e.conj().T
    if isinstance(noise, NoiseModel):
        rho = apply_noise(rho, op, noise, None)
    return np.asarray(rho)


# --------------------------------------------------------------------------------------------------
# states -- src/hilbert.jl:835-882
# --------------------------------------------------------------------------------------------------


def zero_state(N: int) -> np.ndarray:
    s = np.zeros(1 << N, dtype=C)
    s[0] = 1
    return s


def one_state(N: int) -> np.ndarray:
    s = np.zeros(1 << N, dtype=C)
    s[-1] = 1
    return s


def plus_state(N: int) -> np.ndarray:
    return np.full(1 << N, 1.0 / math.sqrt(2 ** N), dtype=C)


def product_state(bits: Sequence[int]) -> np.ndarray:
    """src/hilbert.jl:835 (``sign.`` => any positive entry is |1>)."""
    s = np.zeros(1 << len(bits), dtype=C)
    s[bin2int([1 if b > 0 else 0 for b in bits])] = 1
    return s


def neel_state01(N: int) -> np.ndarray:
    return product_state([0 if i % 2 == 1 else 1 for i in range(1, N + 1)])


def neel_state10(N: int) -> np.ndarray:
    return product_state([1 if i % 2 == 1 else 0 for i in range(1, N + 1)])


def random_state(N: int, gen: np.random.Generator) -> np.ndarray:
    """src/hilbert.jl:882 (uniform re/im in [0,1), normalised)."""
    v = gen.random(1 << N) + 1j * gen.random(1 << N)
    return _normalize(v.astype(C))


def rho_from_state(state: np.ndarray) -> np.ndarray:
    return np.outer(state, state.conj())


# --------------------------------------------------------------------------------------------------
# sampling and observables -- src/ops.jl:46-132, :928-1019; src/func.jl:17-23, :91-182
# --------------------------------------------------------------------------------------------------


def sample(state: np.ndarray, us: Sequence[float]) -> np.ndarray:
    """src/ops.jl:46-62 with the inverse-CDF contract of SURVEY App. A.6 (StatsBase unpinned):
    t = u * sum(p); k = first i with cumsum_i >= t.  0-based basis indices."""
    probs = np.abs(state) ** 2
    cdf = np.cumsum(probs)
    t = np.asarray(us, dtype=float) * cdf[-1]
    k = np.searchsorted(cdf, t, side="left")
    return np.minimum(k, len(cdf) - 1).astype(np.int64)


def sample_exact(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """src/ops.jl:98-101 (state) / :129-132 (rho): structurally non-zero probabilities, ascending index."""
    p = np.abs(x) ** 2 if x.ndim == 1 else np.real(np.diag(x))
    nz = np.nonzero(p)[0]
    return nz.astype(np.int64), p[nz]


def get_probs_from_sample(samples: Sequence[int], N: int) -> Tuple[np.ndarray, np.ndarray]:
    """src/ops.jl:76-93."""
    vals, counts = np.unique(np.asarray(samples, dtype=np.int64), return_counts=True)
    return vals, counts / float(len(samples))


def sample_to_expectation(bitstr: Sequence[int], prob: Sequence[float], N: int, qubits: Sequence[int]) -> float:
    """src/func.jl:166-182: +p for even parity of the listed qubits, -p otherwise."""
    tot = 0.0
    for a, p in zip(bitstr, prob):
        f = int2bin(int(a), N)
        tot += p if sum(f[q - 1] for q in qubits) % 2 == 0 else -p
    return tot


def mag_moments(N: int, bitstr: Sequence[int], prob: Sequence[float], order: int) -> float:
    """src/func.jl:234-237."""
    mag = np.array([sum(mag_basis(int(a), N)) for a in bitstr], dtype=float)
    return float(np.sum(mag ** order * np.asarray(prob)))


def string_to_matrix(list_of_operators: str):
    """src/ops.jl:1016-1019."""
    return _foldl_kron([sp.csc_matrix(gates(s).astype(C)) for s in list_of_operators.split(",")])


def expand_multi_op(list_of_operators: str, qubits: Sequence[int], N: int):
    """src/ops.jl:928-944."""
    ops_str = list_of_operators.split(",")
    if len(ops_str) != len(qubits):
        raise ValueError("qubit number does not match with operators")
    result = ["I"] * N
    for o, q in zip(ops_str, qubits):
        result[q - 1] = o
    return string_to_matrix(",".join(result))


def expect(x: np.ndarray, what):
    """src/func.jl:91-101: Op -> scalar; operator-name string -> per-qubit vector; matrix -> scalar."""
    N = get_N(x)

    def one(M):
        if x.ndim == 1:
            return float(np.real(np.vdot(x, M @ x)))
        return float(np.real((M.T.multiply(x)).sum())) if sp.issparse(M) else float(np.real(np.trace(x @ M)))

    if isinstance(what, Op):
        return one(what.expand(N))
    if isinstance(what, str):
        return [one(expand_multi_op(what, [q], N)) for q in range(1, N + 1)]
    return one(what)


def correlation(x: np.ndarray, list_of_operators: str, qubits: Sequence[int]) -> float:
    """src/func.jl:139-147."""
    M = expand_multi_op(list_of_operators, list(qubits), get_N(x))
    if x.ndim == 1:
        return float(np.real(np.vdot(x, M @ x)))
    return float(np.real((M.T.multiply(x)).sum()))


def correlation_z(state: np.ndarray, qubits: Sequence[int]) -> float:
    """src/func.jl:17-23 (parity correlation from exact probabilities)."""
    a, b = sample_exact(state)
    return sample_to_expectation(a, b, get_N(state), qubits)


def inner(a: np.ndarray, b: np.ndarray) -> complex:
    """src/tensor.jl:199 analogue for state vectors (<a|b>)."""
    return complex(np.vdot(a, b))


def fidelity(a: np.ndarray, b: np.ndarray) -> float:
    """src/tensor.jl:219 (|<a|b>|^2)."""
    return float(abs(np.vdot(a, b)) ** 2)


def _sqrt_psd(m: np.ndarray) -> np.ndarray:
    """Julia's ``sqrt(::Matrix)`` for a Hermitian positive semi-definite argument: eigen-decomposition, square roots of the
    eigenvalues (tiny negative rounding values clipped)."""
    m = np.asarray(m.todense()) if sp.issparse(m) else np.asarray(m)
    w, v = np.linalg.eigh((m + m.conj().T) / 2)
    return (v * np.sqrt(np.clip(w, 0, None))) @ v.conj().T


def fidelity_rho(rho, sigma) -> float:
    """src/tensor.jl:222-229: a = sqrt(rho) * sigma * sqrt(rho); b = sqrt(a); real(tr(b)^2)."""
    sigma = np.asarray(sigma.todense()) if sp.issparse(sigma) else np.asarray(sigma)
    r = _sqrt_psd(rho)
    a = r @ sigma @ r
    b = _sqrt_psd(a)
    return float(np.real(np.trace(b) ** 2))


# --------------------------------------------------------------------------------------------------
# drivers -- src/ops.jl:599-692, :790-844 (no layout/compile: ops are applied raw)
# --------------------------------------------------------------------------------------------------


def to_state(ops, N: int, noise=False, draws: Optional[Draws] = None) -> np.ndarray:
    """src/ops.jl:790-795 without twirl/ZNE expansion."""
    return apply_ops(zero_state(N), ops, noise=noise, draws=draws)


def to_rho(ops, N: int, noise=False) -> np.ndarray:
    """src/ops.jl:806-844 without twirl/ZNE expansion."""
    rho = rho_from_state(zero_state(N))
    for o in ops:
        rho = _apply_rho(rho, o, noise)
    return rho


def run(ops, N: int, shots: int, noise=False, draws: Optional[Draws] = None) -> List[List[int]]:
    """src/ops.jl:659-692 (statevector backend): mid-circuit outcomes per shot."""
    out = []
    for _ in range(shots):
        _, mid = apply_ops(zero_state(N), ops, noise=noise, draws=draws, track_measurements=True)
        out.append(mid)
    return out


def shadow(ops, N: int, number_of_experiment: int, noise=False, draws: Optional[Draws] = None) -> np.ndarray:
    """src/ops.jl:145-185: classical shadow.  Draw order per experiment as in the reference: circuit draws (to_state, :152),
    then ``rand(1:3)`` per qubit (:157), then the single shot (:162).  Dense instead of sparse."""
    draws = draws or Draws(0)
    m_list = [GATE["H"], GATE["HSP"], GATE["I"]]
    eye = np.eye(2, dtype=C)
    rho = np.zeros((1 << N, 1 << N), dtype=C)
    for _ in range(number_of_experiment):
        state = to_state(ops, N, noise=noise, draws=draws)
        basis = []
        for q in range(1, N + 1):
            m = draws.randint(3)
            state = apply(state, Op(["Xbasis", "Ybasis", "Zbasis"][m], q, mat=m_list[m]))  # a rotation, not a measurement (:158)
            basis.append(m)
        k = int(sample(state, [draws.uniform()])[0])
        bits = int2bin(k, N)
        snap = np.ones((1, 1), dtype=C)
        for b, m in zip(bits, basis):
            bv = np.zeros((2, 1), dtype=C)
            bv[b, 0] = 1.0
            um = m_list[m]
            snap = np.kron(snap, 3.0 * (um.conj().T @ bv @ bv.conj().T @ um) - eye)  # :173
        rho += snap / number_of_experiment
    if not np.isclose(np.trace(rho), 1.0):
        raise ValueError("la.tr(rho)!≈1")
    return rho


def final_measurement(state: np.ndarray, basis: str = "Z", draws: Optional[Draws] = None) -> np.ndarray:
    """src/ops.jl:856-914 without readout noise (broken in the reference)."""
    N = get_N(state)
    for q in range(1, N + 1):
        if basis == "Z":
            m = GATE["I"]
        elif basis == "X":
            m = GATE["H"]
        elif basis == "Y":
            m = GATE["HSP"]
        elif basis == "R":
            m = [GATE["H"], GATE["HSP"], GATE["I"]][draws.randint(3)]
        else:
            raise ValueError("measurement_basis error!")
        state = hilbert1(N, m, q) @ state
    return state


def born_measure_Z2(N: int, state: np.ndarray, qubit1: int, qubit2: int, draws: Draws) -> Tuple[np.ndarray, int]:
    """src/hilbert.jl:705-721 (two-qubit Born measurement; returns the 1-based index like the reference)."""
    slist = [expand_multi_op(names, [qubit1, qubit2], N) @ state for names in ("P0,P0", "P1,P0", "P0,P1", "P1,P1")]
    probs = [abs(np.vdot(state, s)) for s in slist]
    ind = weighted_sample(probs, draws)
    return _normalize(slist[ind]), ind + 1


def entanglement_entropy(psi: np.ndarray) -> float:
    """src/func.jl:299-312 (Julia reshape is column-major: rows = low N/2 index bits)."""
    N = get_N(psi)
    part_a = N // 2
    mat = psi.reshape((1 << part_a, 1 << (N - part_a)), order="F")
    spec = np.linalg.svd(mat, compute_uv=False) ** 2
    spec = spec[spec > 0]
    return float(np.sum(-spec * np.log(spec)))


def bipartition_trace(rho) -> np.ndarray:
    """src/linalg.jl:151-161: ptr[i, j] = sum_k rho[i + k d, j + k d], d = 2^(N/2) (keeps the low half of the index = the last
    N/2 qubits); ``Int(N/2)`` throws for odd N."""
    rho = np.asarray(rho.todense()) if sp.issparse(rho) else np.asarray(rho)
    N = get_N(rho)
    if N % 2:
        raise ValueError("InexactError: Int(N/2)")
    d = 1 << (N // 2)
    out = np.zeros((d, d), dtype=C)
    for k in range(d):
        out += rho[k * d:(k + 1) * d, k * d:(k + 1) * d]
    return out


def entanglement_entropy_rho(rho):
    """src/func.jl:323-328: singular values (not squared) of the half-traced density matrix -> (entropy, -log.(spec))."""
    spec = np.linalg.svd(bipartition_trace(rho), compute_uv=False)
    spec = spec[spec > 0]
    return float(np.sum(-spec * np.log(spec))), -np.log(spec)


# ---- variational front end (callers of apply + expect): src/vqa.jl ---------------------------------------------
GATES_WITH_PHASE = ["P", "RX", "RY", "RZ", "U1", "U2", "U3", "CP", "GIVENS", "FSIM", "SWAPA", "RXX", "RYY", "RZZ", "RXY"]  # src/gates.jl:361
TWO_QUBIT_GATES = ["CX", "CNOT", "CY", "CZ", "CP", "RXX", "RYY", "RZZ", "RXY", "GIVENS", "FSIM", "SWAP", "ISWAP", "FSWAP", "SYC", "ECR"]  # src/gates.jl:65
_N_ARGS = {"U2": 2, "U3": 3, "FSIM": 2}  # _find_argument_number of the gate functions (src/gates.jl:369-420); every other phase gate takes 1


def hamiltonian(N: int, string_of_ops: Sequence, boundary: str = "open"):
    """src/vqa.jl:36-67: H as a 2^N x 2^N sparse matrix, summed term by term with expand_multi_op."""
    couplings = [o for i, o in enumerate(string_of_ops) if i % 2 == 0]
    names = [o for i, o in enumerate(string_of_ops) if i % 2 == 1]
    Hm = sp.csc_matrix((1 << N, 1 << N), dtype=C)
    for idx, op in enumerate(names):
        len_op = len(op.split(","))
        if boundary == "open":
            for site in range(1, N - (len_op - 1) + 1):
                Hm = Hm + couplings[idx] * expand_multi_op(op, list(range(site, site + len_op)), N)
        elif boundary == "periodic":
            for site in range(1, N + 1):
                Hm = Hm + couplings[idx] * expand_multi_op(op, [(i - 1) % N + 1 for i in range(site, site + len_op)], N)
    return Hm


def variational_circuit_from_string(N: int, ops: Sequence[str], deep_circuit: bool = False):
    """src/vqa.jl:340-404 -> ([(name, qubit, target), ...], args per op, dim)."""
    out, args, dim, brick_c = [], [], 0, 0
    for gate_name in ops:
        gate_name = _clean_name(gate_name)
        two = gate_name in TWO_QUBIT_GATES
        c = 1
        mb = 0 if deep_circuit else brick_c % 2
        while c <= N:
            if two:
                if mb + c >= N:
                    break
                r = (gate_name, mb + c, mb + c + 1)
                c = c + 1 if deep_circuit else c + 2
            else:
                r = (gate_name, c, -1)
                c = c + 1
            a = _N_ARGS.get(gate_name, 1) if gate_name in GATES_WITH_PHASE else 0
            dim += a
            args.append(a)
            out.append(r)
        if two:
            brick_c += 1
    return out, args, dim


def efficient_su2(N: int, reps: int = 1, gate_names: Sequence[str] = ("RY", "RZ"), deep_circuit: bool = True):
    """src/vqa.jl:302-314."""
    gl: List[str] = []
    for i in range(1, reps + 2):
        gl += list(gate_names)
        if i <= reps:
            gl.append("CX")
    return variational_circuit_from_string(N, gl, deep_circuit)


def variational_apply(pars: Sequence[float], N: int, vops, args, init: Optional[np.ndarray] = None, noise=False, draws: Optional[Draws] = None) -> np.ndarray:
    """src/vqa.jl:420-456: op by op, the next ``fn`` parameters go into the op's matrix function; noise after every op."""
    state = zero_state(N) if init is None else np.array(init, dtype=C)
    c = 0
    for (name, q, t), fn in zip(vops, args):
        full = name if fn == 0 else f"{name}({','.join(repr(float(p)) for p in pars[c:c + fn])})"
        op = Op(full, q, t) if t != -1 else Op(full, q)
        state = apply(state, op)
        c += fn
        if isinstance(noise, NoiseModel):
            state = apply_noise(state, op, noise, draws)
    return state


def loss_and_grad_paramshift(p: Sequence[float], loss, N: int, vops, args, init: Optional[np.ndarray] = None):
    """src/vqa.jl:588-611: shift pi/2, g_i = (f(p + s e_i) - f(p - s e_i)) / 2."""
    p = np.asarray(p, dtype=float)
    f = lambda v: float(loss(variational_apply(v, N, vops, args, init)))
    l0 = f(p)
    g = np.zeros(len(p))
    base = p.copy()
    for i in range(len(p)):
        base[i] = p[i] + math.pi / 2
        fp = f(base)
        base[i] = p[i] - math.pi / 2
        fm = f(base)
        g[i] = 0.5 * (fp - fm)
        base[i] = p[i]
    return l0, g
