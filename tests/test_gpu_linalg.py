"""Dense linear algebra on the device (csrc/bt_linalg.cu): Schmidt spectra / entanglement entropy by one-sided Jacobi,
partial_trace with any number of kept qubits, expect() for controlled and 2-qubit operators on states and density matrices,
reduced density matrices of density matrices -- each against the oracle's restatement of the reference (numpy SVD / dense
operators) at sizes it finishes in seconds, plus closed-form known answers."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rand_state(N, seed, nb=1):
    g = np.random.default_rng(seed)
    v = g.normal(size=(nb, 1 << N)) + 1j * g.normal(size=(nb, 1 << N))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v[0] if nb == 1 else v


@pytest.mark.parametrize("N", [1, 2, 3, 4, 7, 8, 11, 12, 16])
def test_entanglement_entropy_matches_the_svd_of_the_oracle(bt, orc, N):
    """src/func.jl:299-312 on random (volume-law) states and on states after a shallow brickwork (decaying spectra)."""
    v = rand_state(N, 100 + N)
    s = bt.CuState.from_numpy(v)
    e, logs = bt.entanglement_entropy(s, spectrum_bool=True)
    assert abs(e - orc.entanglement_entropy(v)) < TOL
    pa = N // 2
    ref = np.linalg.svd(v.reshape((1 << pa, 1 << (N - pa)), order="F"), compute_uv=False) ** 2
    assert np.max(np.abs(np.exp(-logs) - ref[: len(logs)])) < 1e-12
    assert np.max(np.abs(s.to_numpy() - v)) == 0.0  # the state itself is untouched (the iteration runs on a scratch copy)
    if N >= 4:
        ops_d, ops_o = [], []
        g = np.random.default_rng(N)
        for l in range(3):
            for q in range(1, N + 1):
                th = float(g.uniform(0, 2 * np.pi))
                ops_d.append(bt.Op(f"RY({th!r})", q)); ops_o.append(orc.Op(f"RY({th!r})", q))
            for q in range(1 + (l % 2), N, 2):
                ops_d.append(bt.Op("CNOT", q, q + 1)); ops_o.append(orc.Op("CNOT", q, q + 1))
        st = bt.apply(ops_d, bt.zero_state(N))
        ref_state = orc.apply_ops(orc.zero_state(N), ops_o)
        assert abs(bt.entanglement_entropy(st) - orc.entanglement_entropy(ref_state)) < TOL


def test_entropy_known_answers_and_20_qubits(bt, orc):
    """product state: 0; GHZ: ln 2; N/2 Bell pairs across the cut: (N/2) ln 2; a 20-qubit state (1024 x 1024 reshape)."""
    assert abs(bt.entanglement_entropy(bt.zero_state(10))) < 1e-12
    N = 10
    ghz = bt.apply([bt.Op("H", 1)] + [bt.Op("CNOT", q, q + 1) for q in range(1, N)], bt.zero_state(N))
    assert abs(bt.entanglement_entropy(ghz) - math.log(2)) < TOL
    bell = bt.apply([bt.Op("H", q) for q in range(1, 6)] + [bt.Op("CNOT", q, q + 5) for q in range(1, 6)], bt.zero_state(N))
    assert abs(bt.entanglement_entropy(bell) - 5 * math.log(2)) < TOL
    v = rand_state(20, 20)
    s = bt.CuState.from_numpy(v)
    assert abs(bt.entanglement_entropy(s) - orc.entanglement_entropy(v)) < TOL


def test_entropy_of_every_trajectory_of_a_batch(bt, orc):
    v = rand_state(9, 3, nb=7)
    v[2] = 0
    v[2, 5] = 1.0  # one product state among them
    s = bt.CuState.from_numpy(v)
    e = bt.entanglement_entropy(s)
    assert e.shape == (7,)
    for t in range(7):
        assert abs(e[t] - orc.entanglement_entropy(v[t])) < TOL


def test_schmidt_spectrum_at_any_cut(bt):
    """bt_sv_schmidt_spectrum for every cut position (the reference only uses N/2): squared singular values of the reshape."""
    N = 9
    v = rand_state(N, 5)
    s = bt.CuState.from_numpy(v)
    for n_low in range(0, N + 1):
        k = min(n_low, N - n_low)
        spec = np.empty(1 << k)
        sw = C.c_int()
        bt._lib.check(s.lib.bt_sv_schmidt_spectrum(s.h, n_low, bt._lib.pdouble(spec), C.byref(sw)))
        ref = np.linalg.svd(v.reshape((1 << n_low, 1 << (N - n_low)), order="F"), compute_uv=False) ** 2
        assert np.max(np.abs(spec - ref)) < 1e-12, n_low
        assert abs(spec.sum() - 1) < 1e-12
    with pytest.raises(bt._lib.BTError):
        bt._lib.check(s.lib.bt_sv_schmidt_spectrum(s.h, N + 1, bt._lib.pdouble(np.empty(4)), None))


@pytest.mark.parametrize("N,keep", [(6, [1, 2, 4, 6]), (9, [2, 3, 5, 8]), (10, [1, 3, 4, 7, 10]), (10, [10, 2, 9, 4, 6, 1]), (11, [1, 2, 3, 4, 5, 6, 7]),
                                    (12, list(range(3, 11))), (12, [12, 11, 10, 9, 8, 7, 6, 5, 4, 3]), (5, [1, 2, 3, 4, 5])])
def test_partial_trace_keeping_many_qubits(bt, orc, N, keep):
    """partial_trace(state, keep) src/linalg.jl:83-140 for 4..10 kept qubits, listed in any order."""
    v = rand_state(N, 7 * N + len(keep))
    got = bt.partial_trace(bt.CuState.from_numpy(v), keep)
    ref = orc.partial_trace_general(v, keep)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < TOL
    assert abs(np.trace(got) - 1) < 1e-12


def test_partial_trace_many_qubits_batched(bt, orc):
    v = rand_state(8, 1, nb=3)
    got = bt.partial_trace(bt.CuState.from_numpy(v), [2, 3, 5, 7, 8])
    for t in range(3):
        assert np.max(np.abs(got[t] - orc.partial_trace_general(v[t], [2, 3, 5, 7, 8]))) < TOL


def test_expect_of_controlled_and_two_qubit_ops_on_states_and_rho(bt, orc):
    """expect(x, op) src/func.jl:91-92 for every kind of Op: controlled 1-qubit, controlled 2-qubit (CCX / CCZ / CSWAP style),
    2-qubit on a density matrix, non-adjacent pairs, qubit > target."""
    N = 6
    v = rand_state(N, 42)
    s = bt.CuState.from_numpy(v)
    g = np.random.default_rng(0)
    A = g.normal(size=(8, N, N)) + 1j * g.normal(size=(8, N, N))
    mix = sum(np.outer(w, w.conj()) for w in [rand_state(N, 50 + i) for i in range(3)]) / 3
    rho = bt.CuRho.from_numpy(mix)
    cases = [("X", (2,), 5), ("RY(0.7)", (6,), 1), ("H", (3,), 4), ("CX", (1, 2), -2), ("CZ", (5, 2), -2), ("RXX(0.4)", (6, 3), -2),
             ("SWAP", (2, 5), -2), ("FSIM(0.3,0.9)", (4, 5), -2), ("CX", (2, 3), 1), ("CZ", (3, 4), 6), ("SWAP", (2, 3), 5), ("CX", (5, 4), 2)]
    for name, qs, control in cases:
        od = bt.Op(name, *qs, control=control)
        oo = orc.Op(name, qs[0], qs[1] if len(qs) == 2 else -1, control)
        assert abs(bt.expect(s, od) - orc.expect(v, oo)) < TOL, (name, qs, control)
        assert abs(bt.expect(rho, od) - orc.expect(mix, oo)) < TOL, (name, qs, control)
    # batched states: one value per trajectory
    vb = rand_state(5, 9, nb=4)
    sb = bt.CuState.from_numpy(vb)
    e = bt.expect(sb, bt.Op("X", 2, control=4))
    for t in range(4):
        assert abs(e[t] - orc.expect(vb[t], orc.Op("X", 2, -1, 4))) < TOL


def test_density_matrix_partial_trace_and_entropy(bt, orc):
    """partial_trace(rho, dims, trace_out) src/linalg.jl:88-140, bipartition_trace :151-161, entanglement_entropy(rho)
    src/func.jl:323-328 (N even; odd N throws like ``Int(N/2)``)."""
    N = 6
    ws = [rand_state(N, 70 + i) for i in range(4)]
    mix = sum(p * np.outer(w, w.conj()) for p, w in zip([0.4, 0.3, 0.2, 0.1], ws))
    rho = bt.CuRho.from_numpy(mix)
    for keep in ([1], [6], [2, 5], [6, 1, 3], [1, 2, 3, 4], [2, 3, 4, 5, 6], list(range(1, 7)), []):
        got = bt.partial_trace_rho(rho, keep)
        ref = orc.partial_trace_rho(mix, [2] * N, [q for q in range(1, N + 1) if q not in keep])
        assert np.max(np.abs(got - ref)) < TOL, keep
    half = bt.partial_trace_rho(rho, [4, 5, 6])
    assert np.max(np.abs(half - orc.bipartition_trace(mix))) < TOL
    e, logs = bt.entanglement_entropy(rho)
    eo, lo = orc.entanglement_entropy_rho(mix)
    assert abs(e - eo) < TOL and np.max(np.abs(np.sort(logs) - np.sort(lo))) < 1e-8
    # pure product state across the cut: reduced matrix is a projector, entropy 0
    st = bt.apply([bt.Op("H", 1), bt.Op("CNOT", 1, 2), bt.Op("RY(0.3)", 5)], bt.zero_state(6))
    e0, _ = bt.entanglement_entropy(bt.to_rho_from_state(st) if hasattr(bt, "to_rho_from_state") else bt.CuRho.from_numpy(np.outer(st.to_numpy(), st.to_numpy().conj())))
    assert abs(e0) < TOL
    with pytest.raises(ValueError):
        bt.entanglement_entropy(bt.CuRho(5))


def test_fidelity_of_density_matrices(bt, orc):
    """fidelity(rho, sigma) src/tensor.jl:222-229 on the device (Jacobi eigen-decomposition with accumulated eigenvectors + two GEMMs +
    Jacobi): mixed-mixed against the oracle's restatement, and closed forms -- F(rho, rho) = 1, F(|a><a|, |b><b|) = |<a|b>|^2,
    F(|a><a|, sigma) = <a|sigma|a>, symmetry, a noisy circuit against its ideal state.
    Tolerance 1e-6: the formula takes square roots of eigenvalues that are zero up to rounding (rank-deficient rho, sigma), and
    sqrt(1e-16) = 1e-8 per null direction enters tr sqrt(.) in the reference, the oracle and here alike."""
    FT = 1e-6
    rng = np.random.default_rng(12)

    def mixed(N, k, seed):
        ws = [rand_state(N, seed + i) for i in range(k)]
        p = rng.random(k)
        p /= p.sum()
        return sum(pi * np.outer(w, w.conj()) for pi, w in zip(p, ws))

    for N in (1, 2, 3, 5, 7):
        r, s_ = mixed(N, 1 << min(N, 3), 300 + N), mixed(N, 3, 400 + N)
        R, S = bt.CuRho.from_numpy(r), bt.CuRho.from_numpy(s_)
        f = bt.fidelity(R, S)
        assert abs(f - orc.fidelity_rho(r, s_)) < FT, N
        assert abs(f - bt.fidelity(S, R)) < FT
        assert abs(bt.fidelity(R, R) - 1) < FT
        a, b = rand_state(N, 1), rand_state(N, 2)
        Pa, Pb = np.outer(a, a.conj()), np.outer(b, b.conj())
        assert abs(bt.fidelity(bt.CuRho.from_numpy(Pa), bt.CuRho.from_numpy(Pb)) - abs(np.vdot(a, b)) ** 2) < FT
        assert abs(bt.fidelity(bt.CuRho.from_numpy(Pa), S) - np.real(np.vdot(a, s_ @ a))) < FT
    # a noisy 6-qubit circuit against the ideal one (the use the reference documents: how far noise moved the state)
    N = 6
    ops = [bt.Op("H", 1)] + [bt.Op("CNOT", q, q + 1) for q in range(1, N)] + [bt.Op("RY(0.4)", 3)]
    ideal = bt.apply(ops, bt.CuRho(N))
    noisy = bt.apply(ops, bt.CuRho(N), noise=bt.NoiseModel("depolarizing", 0.02))
    f = bt.fidelity(ideal, noisy)
    fo = orc.fidelity_rho(ideal.to_numpy(), noisy.to_numpy())
    assert abs(f - fo) < FT and 0.3 < f < 0.999
    with pytest.raises(TypeError):
        bt.fidelity(ideal, bt.zero_state(N))
