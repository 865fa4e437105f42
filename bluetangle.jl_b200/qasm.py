"""OpenQASM 2.0 front door for the device path ("next" row of SURVEY 8f): statement list -> Op list, with the same
gate-name mapping as the reference importer (src/qasm.jl:285-421: cx a,b -> Op("X", b; control=a), sdg -> "SD",
sx -> "XSQRT", qubit q[i] -> label i+1, measure -> Op("MZ"), reset -> Op("RES")).  Table-driven; host-side only."""
from __future__ import annotations

import re
from typing import List

from .gates import split_args
from .host import Op

_Q = r"([a-z_][a-z0-9_]*\[\d+\])"

_ONE = {"x": "X", "y": "Y", "z": "Z", "h": "H", "s": "S", "sdg": "SD", "t": "T", "tdg": "TD", "sx": "XSQRT", "id": "I"}
_ONE_PARAM = {"rx": "RX", "ry": "RY", "rz": "RZ", "u1": "U1", "p": "P", "u2": "U2", "u3": "U3", "u": "U3"}
_CTRL_PARAM = {"cu1": "U1", "cp": "U1", "crx": "RX", "cry": "RY", "crz": "RZ", "cu3": "U3"}
_CTRL = {"cx": "X", "cnot": "X", "cy": "Y", "cz": "Z", "ch": "H"}
_TWO = {"swap": "SWAP", "iswap": "ISWAP", "fswap": "FSWAP", "syc": "SYC", "ecr": "ECR"}
_TWO_PARAM = {"rxx": "RXX", "ryy": "RYY", "rzz": "RZZ", "rxy": "RXY", "givens": "GIVENS", "fsim": "FSIM", "swapa": "SWAPA"}
_THREE = {"ccx": "CCX", "ccy": "CCY", "ccz": "CCZ", "cswap": "CSWAP"}
_SKIP = ("openqasm", "include", "qreg", "creg", "opaque", "gate", "barrier")


def _label(tok: str) -> int:
    m = re.fullmatch(r"[a-z_][a-z0-9_]*\[(\d+)\]", tok.strip())
    if not m:
        raise ValueError(f"cannot parse qubit reference {tok!r}")
    return int(m.group(1)) + 1


def from_qasm(text: str) -> List:
    """Parse OpenQASM 2.0 text into a list of ops for ``apply`` / ``run``."""
    src = re.sub(r"//.*", "", text)
    ops: List = []
    for raw in src.split(";"):
        s = " ".join(raw.split()).lower()
        if not s or s.startswith(_SKIP):
            continue
        m = re.fullmatch(rf"measure {_Q}\s*->\s*\S+", s)
        if m:
            ops.append(Op("MZ", _label(m.group(1))))
            continue
        m = re.fullmatch(rf"reset {_Q}", s)
        if m:
            ops.append(Op("RES", _label(m.group(1))))
            continue
        m = re.fullmatch(r"([a-z0-9]+)\s*(.*)", s)
        if not m:
            raise ValueError(f"Unsupported or unrecognized statement: {raw.strip()};")
        head, rest, args = m.group(1), m.group(2).strip(), None
        if rest.startswith("("):  # argument list: up to the parenthesis that balances the first one
            depth = 0
            for k, ch in enumerate(rest):
                depth += ch == "("
                depth -= ch == ")"
                if depth == 0:
                    args, rest = rest[1:k].strip(), rest[k + 1:].strip()
                    break
            else:
                raise ValueError(f"Unsupported or unrecognized statement: {raw.strip()};")
        if not rest:
            raise ValueError(f"Unsupported or unrecognized statement: {raw.strip()};")
        q = [_label(t) for t in rest.split(",")]
        arg_s = "(" + ",".join(a.strip() for a in split_args(args)) + ")" if args else ""
        if head in _ONE and len(q) == 1 and not args:
            ops.append(Op(_ONE[head], q[0]))
        elif head in _ONE_PARAM and len(q) == 1 and args:
            ops.append(Op(_ONE_PARAM[head] + arg_s, q[0]))
        elif head in _CTRL and len(q) == 2 and not args:
            ops.append(Op(_CTRL[head], q[1], control=q[0]))
        elif head in _CTRL_PARAM and len(q) == 2 and args:
            ops.append(Op(_CTRL_PARAM[head] + arg_s, q[1], control=q[0]))
        elif head in _TWO and len(q) == 2 and not args:
            ops.append(Op(_TWO[head], q[0], q[1]))
        elif head in _TWO_PARAM and len(q) == 2 and args:
            ops.append(Op(_TWO_PARAM[head] + arg_s, q[0], q[1]))
        elif head in _THREE and len(q) == 3 and not args:
            # Op(name, qubit, control_qubit, target_qubit); cswap a,b,c swaps b,c under control a (src/qasm.jl:392)
            ops.append(Op("CSWAP", q[1], q[0], q[2]) if head == "cswap" else Op(_THREE[head], q[0], q[1], q[2]))
        else:
            raise ValueError(f"Unsupported or unrecognized statement: {raw.strip()};")
    return ops
