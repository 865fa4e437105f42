"""Synthetic workloads C1..C5 of SURVEY.md 8(d) / BASELINE.json `configs`, generated host-side from one
numpy Generator(PCG64(seed)) so that the device path, the oracle and the CPU baselines see identical circuits.

Each generator returns a list of plain tuples ``(name, qubit, target, control)`` (target -1 / control -2 when
absent) so that it can be turned into op objects of either the product front end or the oracle.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

Spec = Tuple[str, int, int, int]


def _g(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def c1_brickwork(N: int = 12, depth: int = 20, seed: int = 12) -> List[Spec]:
    """C1: odd layer -> H on every qubit, even layer -> RZ(theta); then CNOT(q,q+1) brickwork."""
    g = _g(seed)
    ops: List[Spec] = []
    for l in range(1, depth + 1):
        for q in range(1, N + 1):
            if l % 2 == 1:
                ops.append(("H", q, -1, -2))
            else:
                ops.append((f"RZ({float(g.uniform(0, 2 * math.pi))!r})", q, -1, -2))
        for q in range(1 + (l % 2), N, 2):
            ops.append(("CNOT", q, q + 1, -2))
    return ops


def qft(N: int) -> List[Spec]:
    """QFT without final swaps: H(q) then CP(pi/2^(j-q)) between q and j = q+1..N (raw, non-adjacent pairs)."""
    ops: List[Spec] = []
    for q in range(1, N + 1):
        ops.append(("H", q, -1, -2))
        for j in range(q + 1, N + 1):
            ops.append((f"CP({math.pi / 2 ** (j - q)!r})", j, q, -2))
    return ops


def layered(N: int, depth: int, seed: int, two: Tuple[str, ...] = ("CNOT", "CZ", "CP")) -> List[Spec]:
    """Random layered circuit: one random gate of (H, RX, RY, RZ, T) per qubit + brickwork of CNOT / CZ / CP(theta)."""
    g = _g(seed)
    ops: List[Spec] = []
    for l in range(1, depth + 1):
        for q in range(1, N + 1):
            k = int(g.integers(5))
            th = float(g.uniform(0, 2 * math.pi))
            name = ["H", f"RX({th!r})", f"RY({th!r})", f"RZ({th!r})", "T"][k]
            ops.append((name, q, -1, -2))
        for q in range(1 + (l % 2), N, 2):
            k = int(g.integers(len(two)))
            th = float(g.uniform(0, 2 * math.pi))
            name = two[k] if two[k] != "CP" else f"CP({th!r})"
            ops.append((name, q, q + 1, -2))
    return ops


def c2_qft_layered(N: int = 28, depth: int = 100, seed: int = 28) -> List[Spec]:
    """C2: QFT(N) followed by `depth` random layers."""
    return qft(N) + layered(N, depth, seed)


def c3_noisy_dm(N: int = 14, depth: int = 20, seed: int = 14, p_dep: float = 0.01, p_ad: float = 0.02):
    """C3: C1-style circuit; after every gate, on the qubits it touched: depolarizing(p_dep) then
    amplitude_damping(p_ad) as OpQC channels (2-qubit product forms after 2-qubit gates).
    Returns a list of ('gate', spec) / ('qc', model, p, qubit, target) entries."""
    out = []
    for name, q, t, c in c1_brickwork(N, depth, seed):
        out.append(("gate", (name, q, t, c)))
        out.append(("qc", "depolarizing", p_dep, q, t))
        out.append(("qc", "amplitude_damping", p_ad, q, t))
    return out


def c4_monitored(N: int = 20, depth: int = 20, seed: int = 20, p_meas: float = 0.1):
    """C4: brickwork of H/CNOT/RZ layers; after each layer every qubit is measured (MZ) with probability p_meas
    (locations fixed at build time, as in src/gates.jl:204-209).  Returns (specs, number_of_measurements)."""
    g = _g(seed)
    ops: List[Spec] = []
    nm = 0
    for l in range(1, depth + 1):
        for q in range(1, N + 1):
            if l % 2 == 1:
                ops.append(("H", q, -1, -2))
            else:
                ops.append((f"RZ({float(g.uniform(0, 2 * math.pi))!r})", q, -1, -2))
        for q in range(1 + (l % 2), N, 2):
            ops.append(("CNOT", q, q + 1, -2))
        for q in range(1, N + 1):
            if g.random() < p_meas:
                ops.append(("MZ", q, -1, -2))
                nm += 1
    return ops, nm


def c5_random(N: int, depth: int = 20, seed: int = 31) -> List[Spec]:
    """C5: depth layers of random 1q gates on all qubits + brickwork CZ/CNOT."""
    return layered(N, depth, seed, two=("CNOT", "CZ"))


def to_ops(mod, specs):
    """Instantiate specs as ``mod.Op`` objects (mod = the product package or the oracle module)."""
    return [mod.Op(name, q, t, control=c) for (name, q, t, c) in specs]
