"""Host logic of the variational front end (bluetangle.jl_b200/vqa.py) and its oracle restatement (src/vqa.jl), no GPU:
term placement of ``hamiltonian``, the ansatz generators, and the parameter-shift rule the gradient loop rests on."""
import math

import numpy as np
import pytest


def dense(m):
    return m.toarray() if hasattr(m, "toarray") else np.asarray(m)


@pytest.mark.parametrize("boundary", ["open", "periodic"])
def test_hamiltonian_terms_sum_to_the_reference_matrix(bt, orc, boundary):
    """src/vqa.jl:36-67: the term list the device path evaluates, expanded with the oracle, is the reference's matrix."""
    N = 5
    spec = [-1.0, "Z,Z", -0.5, "X", 0.3, "X,Y,Z", 0.2, "Y,Y"]
    ps = bt.hamiltonian(N, spec, boundary)
    want = dense(orc.hamiltonian(N, spec, boundary))
    got = np.zeros_like(want)
    for c, names, qs in ps.terms:
        got = got + c * dense(orc.expand_multi_op(names, qs, N))
    assert np.array_equal(got, want)
    n_expected = {"open": 4 + 5 + 3 + 4, "periodic": 20}[boundary]
    assert len(ps) == n_expected
    strings, coefs = ps.strings()
    assert len(strings) == N * len(ps) and len(coefs) == len(ps)
    assert strings[:N] == b"ZZIII"
    if boundary == "periodic":
        assert strings[4 * N:5 * N] == b"ZIIIZ"  # wraps around: sites (5, 1)


def test_pauli_strings_reject_what_the_device_kernel_cannot_take(bt):
    with pytest.raises(ValueError):
        bt.PauliSum(3, [(1.0, "T", [1])]).strings()
    with pytest.raises(ValueError):
        bt.PauliSum(3, [(1.0, "Z,Z", [1, 1])]).strings()
    with pytest.raises(ValueError):
        bt.PauliSum(3, [(1.0, "Z,Z", [1, 4])]).strings()


@pytest.mark.parametrize("deep", [False, True])
def test_ansatz_generators_match_the_restated_reference(bt, orc, deep):
    """src/vqa.jl:340-404: same op sequence, argument counts and dimension as the oracle's restatement."""
    for N, names in ((4, ["RY", "RZ", "CX", "RY", "CZ", "U3", "RXX"]), (5, ["H", "CX", "RZ", "CX", "FSIM"]), (2, ["RX", "CNOT"])):
        vops, args, dim = bt._variational_circuit_from_string(N, names, deep)
        o_ops, o_args, o_dim = orc.variational_circuit_from_string(N, names, deep)
        assert [(v.name, v.qubit, v.target_qubit) for v in vops] == o_ops
        assert args == o_args and dim == o_dim
    su2 = bt.EfficientSU2(4, 2)
    assert len(su2) == 3 * 8 + 2 * 3 and sum(v.nargs for v in su2) == 24  # deep ladder of CX between rotation layers
    brick = bt._variational_circuit_from_string(5, ["CX", "CX"], False)[0]
    assert [(v.qubit, v.target_qubit) for v in brick] == [(1, 2), (3, 4), (2, 3), (4, 5)]  # offset alternates per 2q layer


def test_parameter_shift_equals_the_derivative(orc):
    """src/vqa.jl:590-611: for rotation generators with eigenvalues +-1/2 the pi/2 shift rule is the exact derivative;
    checked on the oracle against central differences (this is what stands in for ForwardDiff across the C ABI)."""
    N = 4
    vops, args, dim = orc.efficient_su2(N, 1, ("RY", "RZ"), True)
    Hm = orc.hamiltonian(N, [-1.0, "Z,Z", -0.7, "X"], "open")
    loss = lambda st: float(np.real(np.vdot(st, Hm @ st)))
    p = np.random.default_rng(3).uniform(0, math.pi, dim)
    l0, g = orc.loss_and_grad_paramshift(p, loss, N, vops, args)
    h = 1e-6
    fd = np.zeros(dim)
    for i in range(dim):
        e = np.zeros(dim); e[i] = h
        fd[i] = (loss(orc.variational_apply(p + e, N, vops, args)) - loss(orc.variational_apply(p - e, N, vops, args))) / (2 * h)
    assert np.max(np.abs(g - fd)) < 1e-7
    assert abs(l0 - loss(orc.variational_apply(p, N, vops, args))) < 1e-14


def test_vqa_golden_fixture_is_reproduced_by_the_oracle(orc):
    """tests/golden/golden_vqa_r1.npz (made by tests/golden/make_golden_vqa.py): regression pin of the restated ansatz,
    hamiltonian and parameter-shift gradient."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_vqa as M

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_vqa_r1.npz"))
    N = 6
    vops, args, dim = orc.variational_circuit_from_string(N, M.NAMES, False)
    st = orc.variational_apply(G["n6_pars"], N, vops, args)
    assert np.max(np.abs(st - G["n6_state"])) < 1e-14
    for key, (spec, bc) in M.HAMS.items():
        for n, v in ((6, st), (12, G["n12_state"])):
            e = float(np.real(np.vdot(v, orc.hamiltonian(n, spec, bc) @ v)))
            assert abs(e - float(G[f"n{n}_energy_{key}"])) < 1e-12


def test_shift_rules_are_exact_for_every_parametrised_gate(bt):
    """shift_rule(gate, arg): sum c * f(theta + s) equals d/dtheta Re<psi|U(theta)' O U(theta)|psi> for every parametrised gate
    of the table (pure numpy: the gate matrices of gates.py, a random state and observable); the reference's fixed pi/2 rule
    (src/vqa.jl:590-611) is exact only for the frequency-1 gates -- for RXX/RYY it is identically zero."""
    rng = np.random.default_rng(11)
    nargs = {"P": 1, "RX": 1, "RY": 1, "RZ": 1, "U1": 1, "U2": 2, "U3": 3, "CP": 1, "GIVENS": 1, "FSIM": 2, "SWAPA": 1, "RXX": 1, "RYY": 1, "RZZ": 1, "RXY": 1}
    for name, na in nargs.items():
        dim = bt.gates(f"{name}({','.join(['0.3'] * na)})").shape[0]
        psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        psi /= np.linalg.norm(psi)
        A = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        O = A + A.conj().T
        th0 = rng.uniform(0, 2 * np.pi, size=na)

        def f(th):
            U = bt.gates(f"{name}({','.join(repr(float(t)) for t in th)})")
            v = U @ psi
            return float(np.real(np.vdot(v, O @ v)))

        for k in range(na):
            def fk(x):
                th = th0.copy()
                th[k] = x
                return f(th)
            h = 1e-5
            num = (fk(th0[k] + h) - fk(th0[k] - h)) / (2 * h)
            got = sum(c * fk(th0[k] + s) for s, c in bt.shift_rule(name, k))
            assert abs(got - num) < 1e-7, (name, k, got, num)
            if name in ("RXX", "RYY"):
                fixed = 0.5 * (fk(th0[k] + np.pi / 2) - fk(th0[k] - np.pi / 2))
                assert abs(fixed) < 1e-12 and abs(num) > 1e-3
