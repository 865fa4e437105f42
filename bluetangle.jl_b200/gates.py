"""Gate and Kraus tables of the host-side front end (mirror of src/gates.jl:22-59, :369-453 and
src/noise.jl:52-131).  Matrices built here are what crosses the C ABI as column-major ComplexF64.

This is product code: it does not import anything from ``oracle/``.
"""
from __future__ import annotations

import ast
import math
import re
from typing import List

import numpy as np

C = np.complex128
_r2 = 1.0 / math.sqrt(2.0)


def _round_sigdigits(x: float, sig: int) -> float:
    if x == 0.0:
        return 0.0
    return round(x, sig - 1 - int(math.floor(math.log10(abs(x)))))


def _round10(m) -> np.ndarray:
    """Julia ``round.(m, sigdigits=10)`` -- T, TD, CT, CTD are stored rounded (src/gates.jl:32-33,47-48)."""
    m = np.array(m, dtype=C)
    f = np.vectorize(lambda z: complex(_round_sigdigits(z.real, 10), _round_sigdigits(z.imag, 10)))
    return f(m).astype(C)


def _a(rows) -> np.ndarray:
    return np.array(rows, dtype=C)


def _diag(*v) -> np.ndarray:
    return np.diag(np.array(v, dtype=C))


_w = complex(math.cos(math.pi / 4), math.sin(math.pi / 4))

gate = {
    "I": _a([[1, 0], [0, 1]]),
    "X": _a([[0, 1], [1, 0]]),
    "SX": 0.5 * _a([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]),
    "Y": _a([[0, -1j], [1j, 0]]),
    "Z": _a([[1, 0], [0, -1]]),
    "H": _r2 * _a([[1, 1], [1, -1]]),
    "S": _a([[1, 0], [0, 1j]]),
    "SD": _a([[1, 0], [0, -1j]]),
    "T": _round10([[1, 0], [0, np.exp(1j * math.pi / 4)]]),
    "TD": _round10([[1, 0], [0, np.exp(-1j * math.pi / 4)]]),
    "HSP": _r2 * _a([[1, -1j], [1, 1j]]),
    "HY": _r2 * _a([[1, 1j], [1, -1j]]),
    "H2": 0.5 * _a([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]),
    "P0": _a([[1, 0], [0, 0]]),
    "P1": _a([[0, 0], [0, 1]]),
    "SP": _a([[0, 1], [0, 0]]),
    "SM": _a([[0, 0], [1, 0]]),
    "CX": _a([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]),
    "CY": _a([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]]),
    "CZ": _diag(1, 1, 1, -1),
    "CS": _diag(1, 1, 1, 1j),
    "CT": _round10(np.diag([1, 1, 1, np.exp(1j * math.pi / 4)])),
    "CTD": _round10(np.diag([1, 1, 1, np.exp(-1j * math.pi / 4)])),
    "CSD": _diag(1, 1, 1, -1j),
    "CI": _diag(1, 1, 1, 1),
    "CH": _a([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, _r2, _r2], [0, 0, _r2, -_r2]]),
    "SWAP": _a([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
    "ISWAP": _a([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]),
    "FSWAP": _a([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, -1]]),
    "SYC": _a([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, np.exp(-1j * math.pi / 6)]]),
    "ECR": _r2 * _a([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]]),
}
gate["XSQRT"] = gate["SX"]
gate["CNOT"] = gate["CX"]
gate["CCX"] = np.eye(8, dtype=C)
gate["CCX"][6:8, 6:8] = gate["X"]
gate["CCZ"] = np.diag(np.array([1, 1, 1, 1, 1, 1, 1, -1], dtype=C))

gates_with_phase = ["P", "RX", "RY", "RZ", "U1", "U2", "U3", "CP", "GIVENS", "FSIM", "SWAPA", "RXX", "RYY", "RZZ", "RXY"]
one_qubit_gates = ["I", "X", "Y", "Z", "SX", "XSQRT", "H", "T", "S", "SD", "P", "U2", "U3"]
two_qubit_gates = ["CX", "CNOT", "CY", "CZ", "CP", "RXX", "RYY", "RZZ", "RXY", "GIVENS", "FSIM", "SWAP", "ISWAP", "FSWAP", "SYC", "ECR"]


def P_(lam):
    return _a([[1, 0], [0, np.exp(1j * lam)]])


def RX_(th):
    c, s = math.cos(th / 2), math.sin(th / 2)
    return _a([[c, -1j * s], [-1j * s, c]])


def RY_(th):
    c, s = math.cos(th / 2), math.sin(th / 2)
    return _a([[c, -s], [s, c]])


def RZ_(th):
    return _a([[np.exp(-1j * th / 2), 0], [0, np.exp(1j * th / 2)]])


def U2_(phi, lam):
    return _r2 * _a([[1, -np.exp(1j * lam)], [np.exp(1j * phi), np.exp(1j * (phi + lam))]])


def U3_(th, phi, lam):
    c, s = math.cos(th / 2), math.sin(th / 2)
    return _a([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]])


def CP_(lam):
    return np.diag(np.array([1, 1, 1, np.exp(1j * lam)], dtype=C))


def GIVENS_(th):
    c, s = math.cos(th), math.sin(th)
    return _a([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]])


def FSIM_(th, phi):
    c, s = math.cos(th), math.sin(th)
    return _a([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, np.exp(1j * phi)]])


def SWAPA_(a):
    e = np.exp(1j * math.pi * a)
    return 0.5 * _a([[2, 0, 0, 0], [0, 1 + e, 1 - e, 0], [0, 1 - e, 1 + e, 0], [0, 0, 0, 2]])


def RXX_(phi):
    # the reference uses cos(phi)/sin(phi), not the half angle (src/gates.jl:401)
    c, s = math.cos(phi), math.sin(phi)
    return _a([[c, 0, 0, -1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [-1j * s, 0, 0, c]])


def RYY_(phi):
    c, s = math.cos(phi), math.sin(phi)
    return _a([[c, 0, 0, 1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [1j * s, 0, 0, c]])


def RZZ_(phi):
    m, p = np.exp(-0.5j * phi), np.exp(0.5j * phi)
    return np.diag(np.array([m, p, p, m], dtype=C))


def RXY_(phi):
    c, s = math.cos(phi), math.sin(phi)
    return _a([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, 1]])


_functions = {
    "P": P_, "RX": RX_, "RY": RY_, "RZ": RZ_, "U1": P_, "U2": U2_, "U3": U3_, "CP": CP_, "GIVENS": GIVENS_,
    "SWAPA": SWAPA_, "FSIM": FSIM_, "RXX": RXX_, "RYY": RYY_, "RZZ": RZZ_, "RXY": RXY_,
}


def clean_name(name: str) -> str:
    return name.split("(", 1)[0].upper()


def is_measurement(name: str) -> bool:
    return name.upper() in {"MZ", "M(Z)", "MX", "M(X)", "MY", "M(Y)", "MR", "M(R)"}


_BIN = {ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b, ast.Div: lambda a, b: a / b}


def _eval_node(n) -> float:
    if isinstance(n, ast.Expression):
        return _eval_node(n.body)
    if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
        return float(n.value)
    if isinstance(n, ast.Name) and n.id == "pi":
        return math.pi
    if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.UAdd, ast.USub)):
        v = _eval_node(n.operand)
        return -v if isinstance(n.op, ast.USub) else v
    if isinstance(n, ast.BinOp) and type(n.op) in _BIN:
        return _BIN[type(n.op)](_eval_node(n.left), _eval_node(n.right))
    raise ValueError("unsupported expression")


def parse_number(text: str) -> float:
    """Numeric gate arguments as the reference writes them: ``0.37``, ``.1pi``, ``0.5π``, ``-pi/2``, ``(1+2)*pi/4``.
    Evaluated by a small AST walker (numbers, pi, + - * /, unary minus, parentheses): no ``eval``, no powers, no names -- gate
    arguments arrive from QASM files."""
    e = text.strip().replace("π", "pi")
    e = re.sub(r"(?<=[0-9\.])\s*pi", "*pi", e)
    if len(e) > 200 or re.fullmatch(r"[0-9eE\.\+\-\*/\(\) pi]+", e) is None or "**" in e:
        raise ValueError(f"cannot parse gate argument {text!r}")
    try:
        return float(_eval_node(ast.parse(e, mode="eval")))
    except (SyntaxError, ValueError, ZeroDivisionError, RecursionError) as err:
        raise ValueError(f"cannot parse gate argument {text!r}") from err


def split_args(inner: str):
    """split a gate argument list at top-level commas (arguments may contain parentheses)"""
    out, depth, cur = [], 0, []
    for ch in inner:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth < 0:
                raise ValueError(f"unbalanced parentheses in {inner!r}")
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if depth != 0:
        raise ValueError(f"unbalanced parentheses in {inner!r}")
    out.append("".join(cur))
    return out


def gates(op_name: str) -> np.ndarray:
    """Matrix for a gate name (src/gates.jl:369-453)."""
    cn = clean_name(op_name)
    un = op_name.upper()
    if cn in _functions:
        inner = op_name.split("(", 1)[1]
        inner = inner[: inner.rindex(")")]
        return _functions[cn](*[parse_number(t) for t in split_args(inner)])
    if un in ("M(Z)", "MZ", "M(R)", "MR", "RES"):
        return gate["I"]
    if un in ("M(X)", "MX"):
        return gate["H"]
    if un in ("M(Y)", "MY"):
        return gate["HSP"]
    if un in gate:
        return gate[un]
    raise KeyError(f"Gate {op_name} not found")


def noise_model(model: str, p: float, two_qubit: bool = False) -> List[np.ndarray]:
    """Kraus operators (src/noise.jl:52-131).  ``phase_flip`` produces the Y operator, as the reference does
    (its comparison of a String with a Symbol at :80 is always false)."""
    model = model.lower()
    I, X, Y, Z = gate["I"], gate["X"], gate["Y"], gate["Z"]
    if model == "amplitude_damping":
        ops = [_a([[1, 0], [0, math.sqrt(1 - p)]]), _a([[0, math.sqrt(p)], [0, 0]])]
    elif model == "phase_damping":
        ops = [_a([[1, 0], [0, math.sqrt(1 - p)]]), _a([[0, 0], [0, math.sqrt(p)]])]
    elif model in ("phase_flip", "bit_flip", "bit_phase_flip"):
        ops = [math.sqrt(1 - p) * I, math.sqrt(p) * (X if model == "bit_flip" else Y)]
    elif model == "depolarizing_amp":
        ops = [math.sqrt(1 - 3 * p / 4) * I] + [math.sqrt(p / 4) * s for s in (X, Y, Z)]
    elif model == "depolarizing":
        ops = [math.sqrt(1 - p) * I] + [math.sqrt(p / 3) * s for s in (X, Y, Z)]
    elif model == "rot_z":
        ops = [RZ_(p)]
    elif model == "rot_y":
        ops = [RY_(p)]
    elif model == "rot_x":
        ops = [RX_(p)]
    elif model == "rot_p":
        ops = [P_(p)]
    elif model == "rot_xyz":
        ops = [m / math.sqrt(3) for m in (RX_(p), RY_(p), RZ_(p))]
    elif is_measurement(model):
        ops = [gate["P0"], gate["P1"]]
    else:
        raise ValueError("Unknown quantum error model")
    if two_qubit:
        return [np.kron(a, b) for a in ops for b in ops]
    return ops


def is_valid_quantum_channel(kraus) -> bool:
    """src/struct.jl:291-309: sum K'K ~ I and Choi matrix Hermitian PSD (eigenvalues rounded to 10 digits)."""
    n = kraus[0].shape[0]
    acc = np.zeros((n, n), dtype=C)
    for K in kraus:
        acc += K.conj().T @ K
    tp = bool(np.allclose(acc, np.eye(n), rtol=1.4901161193847656e-08, atol=0.0))
    choi = np.zeros((n * n, n * n), dtype=C)
    for K in kraus:
        v = np.asarray(K, dtype=C).reshape(-1, order="F")
        choi += np.outer(v, v.conj())
    ev = np.linalg.eigvalsh(choi)
    return tp and bool(np.allclose(choi, choi.conj().T)) and bool(np.all(np.round(ev, 10) >= 0))
