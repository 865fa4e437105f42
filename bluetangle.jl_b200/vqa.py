"""Host-side mirror of the reference's variational front end for device-resident states: the heaviest
*caller* of the apply + expect path.

  hamiltonian            src/vqa.jl:36-67    -> PauliSum (the term list; the reference builds the 2^N x 2^N sparse matrix)
  variational circuits   src/vqa.jl:302-404  (EfficientSU2, generate_ansatz_circuit, _variational_circuit_from_string)
  AnsatzOptions          src/vqa.jl:167-283  (fields used by the state-vector path: N, ops, args, loss, noise, init, dim)
  variational_apply      src/vqa.jl:420-456
  loss_and_grad_paramshift  src/vqa.jl:588-611
  VQA                    src/vqa.jl:488-583  gradient branch on the parameter-shift gradient ("descent", "adam")

Everything numeric runs through the C ABI: the circuit is one ``bt_sv_apply_circuit`` call (its fused passes keep their
structure from step to step, only the angles change, so the pass specialiser compiles them once and every later step of the
optimisation reuses the modules), the loss is one ``bt_sv_expect_pauli_sum`` call (one read of the state per commuting
group of terms instead of one 2^N x 2^N sparse operator).  The derivative-free optimisers the reference reaches through
PRIMA.jl (cobyla, ...) and its ForwardDiff path are host libraries outside this repo's scope: VQA() here offers the
gradient models only and raises for the others.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from . import host as H
from .gates import clean_name, gates_with_phase, two_qubit_gates

_ARGS = {"P": 1, "RX": 1, "RY": 1, "RZ": 1, "U1": 1, "U2": 2, "U3": 3, "CP": 1, "GIVENS": 1, "FSIM": 2, "SWAPA": 1, "RXX": 1, "RYY": 1, "RZZ": 1, "RXY": 1}


# Parameter-shift rules per gate argument: the derivative of loss(theta) with respect to ONE gate parameter is exact as a
# short linear combination of shifted evaluations once the frequencies of the gate in that parameter are known (differences of
# the eigenvalues of its generator).  ("two", w): frequency w alone -> w/2 * (f(x + pi/2w) - f(x - pi/2w)); "four": frequencies
# {1, 2} -> sum_mu (-1)^(mu-1) / (8 sin^2(x_mu/2)) f(x + x_mu), x_mu = (2 mu - 1) pi/4, mu = 1..4.  The reference differentiates
# with ForwardDiff (src/vqa.jl:564) and is exact for every gate; its own loss_and_grad_paramshift (:590-611) uses w = 1 throughout.
#   RX RY RZ P U1 U2 U3 CP RZZ: generator gap 1.   RXX, RYY: exp(-i phi XX) (cos phi, not the half angle: src/gates.jl:401-403) -> w = 2.
#   SWAPA(a): eigenphases 1 and exp(i pi a) -> w = pi.   GIVENS(theta), FSIM(theta, .), RXY(phi): rotation of the {01, 10} block,
#   generator eigenvalues {0, 0, +1, -1} -> frequencies {1, 2}.
_RULES = {"P": [("two", 1.0)], "U1": [("two", 1.0)], "RX": [("two", 1.0)], "RY": [("two", 1.0)], "RZ": [("two", 1.0)],
          "U2": [("two", 1.0)] * 2, "U3": [("two", 1.0)] * 3, "CP": [("two", 1.0)], "RZZ": [("two", 1.0)],
          "RXX": [("two", 2.0)], "RYY": [("two", 2.0)], "SWAPA": [("two", math.pi)],
          "GIVENS": [("four", 0.0)], "FSIM": [("four", 0.0), ("two", 1.0)], "RXY": [("four", 0.0)]}


def shift_rule(gate_name: str, arg: int) -> List[Tuple[float, float]]:
    """[(shift, coefficient)]: d loss / d parameter = sum coefficient * loss(parameter + shift), exact for this gate argument."""
    kind, w = _RULES[clean_name(gate_name)][arg]
    if kind == "two":
        s = math.pi / (2.0 * w)
        return [(s, 0.5 * w), (-s, -0.5 * w)]
    out = []
    for mu in range(1, 5):
        x = (2 * mu - 1) * math.pi / 4.0
        out.append((x, (-1.0) ** (mu - 1) / (8.0 * math.sin(x / 2.0) ** 2)))
    return out


class PauliSum:
    """H = sum_k coef_k * (operator string on qubits): what ``hamiltonian`` (src/vqa.jl:36-67) sums up, kept as terms."""

    def __init__(self, N: int, terms: Sequence[Tuple[float, str, Sequence[int]]]):
        self.N = N
        self.terms = [(float(c), str(names), [int(q) for q in qs]) for c, names, qs in terms]

    def __len__(self):
        return len(self.terms)

    def strings(self) -> Tuple[bytes, np.ndarray]:
        """terms as N-character Pauli strings (qubit 1 first) + coefficients; raises if a factor is not I/X/Y/Z"""
        out, coefs = [], []
        for c, names, qs in self.terms:
            s = ["I"] * self.N
            parts = names.split(",")
            if len(parts) != len(qs):
                raise ValueError("qubit number does not match with operators")
            for n, q in zip(parts, qs):
                n = n.strip().upper()
                if n not in ("I", "X", "Y", "Z"):
                    raise ValueError(f"{n} is not a Pauli operator")
                if not 1 <= q <= self.N:
                    raise ValueError(f"qubit {q} out of range")
                if s[q - 1] != "I" and n != "I":
                    raise ValueError(f"qubit {q} appears twice in a term")
                if n != "I":
                    s[q - 1] = n
            out.append("".join(s))
            coefs.append(c)
        return "".join(out).encode(), np.asarray(coefs, dtype=np.float64)

    def expect(self, x):
        """real(state' * H * state) (src/func.jl:91) / real(tr(rho * H)) (:92)"""
        if isinstance(x, H.CuState):
            try:
                ps, coefs = self.strings()
            except ValueError:
                return H.hamiltonian_expect(x, self.terms)
            out = np.empty(x.n_batch, dtype=np.float64)
            L.check(x.lib.bt_sv_expect_pauli_sum(x.h, len(coefs), ps, L.pdouble(coefs), L.pdouble(out)))
            return float(out[0]) if x.n_batch == 1 else out
        return H.hamiltonian_expect(x, self.terms)


def hamiltonian(N: int, string_of_ops: Sequence, boundary: str = "open") -> PauliSum:
    """src/vqa.jl:36-67: alternating [coupling, "Z,Z", coupling, "X", ...]; a k-site operator string is placed on sites
    site..site+k-1 for site = 1..N-(k-1) ("open") or on mod1(site..site+k-1, N) for site = 1..N ("periodic")."""
    couplings = [string_of_ops[i] for i in range(0, len(string_of_ops), 2)]
    ops = [string_of_ops[i] for i in range(1, len(string_of_ops), 2)]
    terms = []
    for coup, op in zip(couplings, ops):
        k = len(op.split(","))
        if boundary == "open":
            for site in range(1, N - (k - 1) + 1):
                terms.append((coup, op, list(range(site, site + k))))
        elif boundary == "periodic":
            for site in range(1, N + 1):
                terms.append((coup, op, [((i - 1) % N) + 1 for i in range(site, site + k)]))
    return PauliSum(N, terms)


class VOp:
    """one op of a variational circuit: a fixed gate (nargs = 0) or a parametrised one whose matrix is built from the next
    ``nargs`` entries of the parameter vector (the reference stores a function in Op.mat, src/vqa.jl:369-391)"""

    def __init__(self, name: str, qubit: int, target_qubit: int = -1, control: int = -2):
        self.name = clean_name(name)
        self.qubit, self.target_qubit, self.control = qubit, target_qubit, control
        self.nargs = _ARGS[self.name] if self.name in gates_with_phase else 0
        self.q = 2 if target_qubit != -1 else 1

    def bind(self, pars: Sequence[float]) -> "H.Op":
        name = self.name if self.nargs == 0 else f"{self.name}({','.join(repr(float(p)) for p in pars)})"
        if self.q == 2:
            return H.Op(name, self.qubit, self.target_qubit, control=self.control)
        return H.Op(name, self.qubit, control=self.control)

    def __repr__(self):
        return f"VOp({self.name}, {self.qubit}, {self.target_qubit}, nargs={self.nargs})"


def _variational_circuit_from_string(N: int, ops: Sequence[str], deep_circuit: bool = False):
    """src/vqa.jl:340-404: every gate name becomes a layer; 2-qubit names form a brickwork that alternates its offset
    from one 2-qubit layer to the next (or a full ladder when deep_circuit)."""
    out: List[VOp] = []
    args: List[int] = []
    dim = 0
    brick_c = 0
    for gate_name in ops:
        gate_name = clean_name(gate_name)
        two = gate_name in two_qubit_gates
        c = 1
        mb = 0 if deep_circuit else brick_c % 2
        while c <= N:
            if two:
                if mb + c >= N:
                    break
                r = VOp(gate_name, mb + c, mb + c + 1)
                c = c + 1 if deep_circuit else c + 2
            else:
                r = VOp(gate_name, c)
                c += 1
            dim += r.nargs
            args.append(r.nargs)
            out.append(r)
        if two:
            brick_c += 1
    return out, args, dim


def EfficientSU2(N: int, reps: int = 1, gate_names: Sequence[str] = ("RY", "RZ"), deep_circuit: bool = True) -> List[VOp]:
    """src/vqa.jl:302-314"""
    gl: List[str] = []
    for i in range(1, reps + 2):
        gl += list(gate_names)
        if i <= reps:
            gl.append("CX")
    return _variational_circuit_from_string(N, gl, deep_circuit)[0]


def generate_ansatz_circuit(N: int, reps: int = 1, gate_names: Sequence[str] = ("RY", "RZ", "CX"), deep_circuit: bool = True) -> List[VOp]:
    """src/vqa.jl:335-338"""
    return _variational_circuit_from_string(N, list(gate_names) * reps, deep_circuit)[0]


class AnsatzOptions:
    """src/vqa.jl:167-283, the fields the state-vector path uses.  ``loss`` is a PauliSum (the reference's "loss is a
    matrix" case, :282-283: real(expect(state, H))) or a function state -> number.  ``init``: None = zero_state(N), or a
    CuState that is copied at every evaluation."""

    def __init__(self, N: int, ops, loss, noise=False, init: Optional["H.CuState"] = None, model: str = "descent", number_of_iterations: int = 1000,
                 learning_rate: float = 0.01, pars_initial: Optional[Sequence[float]] = None, deep_circuit: bool = False, history: bool = True, rng=None):
        self.N = N
        if len(ops) and isinstance(ops[0], str):
            self.ops, self.args, self.dim = _variational_circuit_from_string(N, list(ops), deep_circuit)
        else:
            self.ops = list(ops)
            self.args = [o.nargs if isinstance(o, VOp) else 0 for o in self.ops]
            self.dim = sum(self.args)
        self.loss: Callable = (lambda st, Hm=loss: Hm.expect(st)) if isinstance(loss, PauliSum) else loss
        self.noise = noise
        self.init = init
        self.model, self.number_of_iterations, self.learning_rate, self.history = model, number_of_iterations, learning_rate, history
        if pars_initial is not None and len(pars_initial) >= self.dim:
            self.pars_initial = np.asarray(pars_initial[: self.dim], dtype=np.float64)
        else:  # src/vqa.jl:262 rand(dim) * pi
            r = H._rng(rng)
            self.pars_initial = np.array([r.uniform() for _ in range(self.dim)], dtype=np.float64) * math.pi

    def bound_ops(self, pars: Sequence[float]) -> List["H.Op"]:
        out, c = [], 0
        for op, fn in zip(self.ops, self.args):
            out.append(op.bind(pars[c:c + fn]) if isinstance(op, VOp) else op)
            c += fn
        return out


def variational_apply(pars: Sequence[float], opt: AnsatzOptions, noise_override=False, rng=None, out: Optional["H.CuState"] = None) -> "H.CuState":
    """src/vqa.jl:420-456: the ansatz with the given parameters on a copy of the initial state.  Without noise the whole
    op list goes to the device in one fused call; with a NoiseModel the per-op loop of the reference is kept (apply, then
    apply_noise, drawing from ``rng``).  ``out``: an existing device state to overwrite instead of allocating a new one
    (the optimisation loops evaluate the ansatz thousands of times; a 28-qubit state is 4 GiB)."""
    if out is None:
        state = H.zero_state(opt.N) if opt.init is None else opt.init.copy()
    else:
        state = out
        if opt.init is None:
            L.check(state.lib.bt_sv_set_basis(state.h, 0))
        else:
            L.check(state.lib.bt_sv_copy(state.h, opt.init.h))
    noise = noise_override if isinstance(noise_override, H.NoiseModel) else opt.noise
    ops = opt.bound_ops(pars)
    if isinstance(noise, H.NoiseModel):
        H.apply(ops, state, noise=noise, rng=rng)
    else:
        H.apply(ops, state)
    return state


def _loss(p, opt: AnsatzOptions) -> float:
    """loss at p on the options' work state (allocated once, overwritten by every evaluation)"""
    work = getattr(opt, "_work", None)
    if work is None:
        work = opt._work = H.zero_state(opt.N) if opt.init is None else opt.init.copy()
    return float(opt.loss(variational_apply(p, opt, out=work)))


def loss_and_grad_paramshift(p: Sequence[float], opt: AnsatzOptions) -> Tuple[float, np.ndarray]:
    """src/vqa.jl:590-611 as written there: g_i = (loss(p + pi/2 e_i) - loss(p - pi/2 e_i)) / 2 for every parameter, in order
    (exact for the frequency-1 gates only -- "for Pauli/Pauli-string rotations", :595; ``loss_and_grad`` is exact for all)."""
    p = np.asarray(p, dtype=np.float64)
    l0 = _loss(p, opt)
    shift = math.pi / 2
    g = np.zeros(len(p))
    base = p.copy()
    for i in range(len(p)):
        base[i] = p[i] + shift
        fp = _loss(base, opt)
        base[i] = p[i] - shift
        fm = _loss(base, opt)
        g[i] = 0.5 * (fp - fm)
        base[i] = p[i]
    return l0, g


def loss_and_grad(p: Sequence[float], opt: AnsatzOptions) -> Tuple[float, np.ndarray]:
    """Exact gradient of the loss for every parametrised gate of the table (the stand-in for ForwardDiff.gradient, src/vqa.jl:564):
    per parameter the shift rule of its gate (``shift_rule``).  For the frequency-1 gates this is loss_and_grad_paramshift."""
    p = np.asarray(p, dtype=np.float64)
    l0 = _loss(p, opt)
    g = np.zeros(len(p))
    base = p.copy()
    i = 0
    for op, fn in zip(opt.ops, opt.args):
        for k in range(fn):
            rule = shift_rule(op.name, k)
            if len(rule) == 2:  # same arithmetic as the reference's fixed rule: w/2 * (f+ - f-)
                base[i] = p[i] + rule[0][0]
                fp = _loss(base, opt)
                base[i] = p[i] + rule[1][0]
                fm = _loss(base, opt)
                g[i] = rule[0][1] * (fp - fm)
            else:
                acc = 0.0
                for sh, c in rule:
                    base[i] = p[i] + sh
                    acc += c * _loss(base, opt)
                g[i] = acc
            base[i] = p[i]
            i += 1
    return l0, g


def VQA(opt: AnsatzOptions):
    """src/vqa.jl:488-583, the gradient branch (:563-582): per iteration the gradient, one optimiser update, and -- when
    ``history`` -- the loss at the NEW parameters.  The reference differentiates with ForwardDiff; across the C ABI the
    per-gate parameter-shift rules (``loss_and_grad``: exact for every parametrised gate the ansatz generators accept,
    including RXX/RYY, SWAPA and the two-frequency gates GIVENS/FSIM/RXY) take its place.  Models: "descent" /
    "gradient" (Optimisers.Descent: p -= lr * g) and "adam" (Optimisers.Adam defaults beta = (0.9, 0.999), eps = 1e-8).
    Returns (energy_history, pars, pars_history) or, with history = False, (loss, pars) like the reference."""
    model = opt.model.lower()
    if model not in ("descent", "gradient", "adam"):
        raise NotImplementedError(f"optimiser model {opt.model!r}: only the gradient models (descent, adam) are mirrored; PRIMA / OptimKit are host libraries outside this backend")
    p = opt.pars_initial.copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    b1, b2, eps = 0.9, 0.999, 1e-8
    energy_history, pars_history = [], []
    for it in range(1, opt.number_of_iterations + 1):
        _, g = loss_and_grad(p, opt)
        if model == "adam":
            m = b1 * m + (1 - b1) * g
            v = b2 * v + (1 - b2) * g * g
            p = p - opt.learning_rate * (m / (1 - b1 ** it)) / (np.sqrt(v / (1 - b2 ** it)) + eps)
        else:
            p = p - opt.learning_rate * g
        if opt.history:
            energy_history.append(_loss(p, opt))
            pars_history.append(p.copy())
    if not opt.history:
        return _loss(p, opt), p
    return energy_history, p, pars_history


VQE = VQA  # the reference documents the loop under both names (src/vqa.jl:458-487 docstring, :488 definition)
