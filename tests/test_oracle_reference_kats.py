"""The reference's own tests (test/runtests.jl) restated against the oracle: this is what pins oracle/bt_oracle.py.
Every test cites the @test block it follows.  CPU only."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import bt_oracle as O


def rand_state(N, seed):
    return O.random_state(N, np.random.default_rng(seed))


def random_ops(N, depth, seed, measure_prob=0.0):
    """src/gates.jl:155-209 (random_ops) with a numpy generator."""
    g = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        c = 1
        while c <= N:
            pool = O.ONE_QUBIT_GATES if c == N else O.ONE_QUBIT_GATES + O.TWO_QUBIT_GATES
            s = pool[g.integers(len(pool))]
            p = [round(float(x) * math.pi, 2) for x in g.normal(size=3)]
            if s in O.ONE_QUBIT_GATES:
                name = {"P": f"P({p[0]})", "U2": f"U2({p[0]},{p[1]})", "U3": f"U3({p[0]},{p[1]},{p[2]})"}.get(s, s)
                ops.append(O.Op(name, c))
                c += 1
            else:
                if s == "FSIM":
                    name = f"FSIM({p[0]},{p[1]})"
                elif s in O.GATES_WITH_PHASE:
                    name = f"{s}({p[0]})"
                else:
                    name = s
                ops.append(O.Op(name, c, c + 1))
                c += 2
        for i in range(1, N + 1):
            if g.random() < measure_prob:
                ops.append(O.Op(["MX", "MY", "MZ"][g.integers(3)], i))
    return ops


def test_int2bin_known_answer():
    """runtests.jl:8"""
    assert O.int2bin(2, 4) == [0, 0, 1, 0]
    assert O.bin2int([0, 0, 1, 0]) == 2
    for a in range(64):
        assert O.bin2int(O.int2bin(a, 6)) == a


def test_nonlocal_hilbert_exact_identity():
    """runtests.jl:10-17: CNOT(2,5) through the Pauli-reconstruction path == X on 5 controlled by 2, with `==`."""
    N = 6
    a = O.Op("CNOT", 2, 5).expand(N)
    b = O.Op("X", 5, control=2).expand(N)
    assert (a != b).nnz == 0


def test_kraus_probabilities_three_ways():
    """runtests.jl:19-35"""
    N = 6
    s = rand_state(N, 1)
    nm = O.QuantumChannel(2, "depolarizing", 0.5)
    pA = O.partial_trace_general(s, [2, 3])
    probs = [np.real(np.trace(k @ pA @ k.conj().T)) for k in nm.kraus]
    probs2 = [np.sum(np.abs(O.hilbert2(N, k, 2, 3) @ s) ** 2) for k in nm.kraus]
    probs3 = O.OpQC("my quantum channel", nm.kraus, 2, 3).prob(s)
    assert np.allclose(probs, probs2) and np.allclose(probs2, probs3)


def test_partial_trace_rho_general_dims():
    """runtests.jl:37-52 uses partial_trace(rho,[4,4],[1]); checked here against an explicit block trace."""
    g = np.random.default_rng(0)
    b = g.normal(size=(16, 16)) + 1j * g.normal(size=(16, 16))
    c = O.partial_trace_rho(sp.csc_matrix(b), [4, 4], [1])
    ref = sum(b[4 * k:4 * k + 4, 4 * k:4 * k + 4] for k in range(4))
    assert np.allclose(c, ref)


def test_sequential_apply_equals_circuit_product():
    """runtests.jl:54-73: N=10, depth 50, apply gate by gate == product of the expanded operators, atol 1e-12."""
    N, depth = 8, 20
    ops = random_ops(N, depth, 2)
    s = rand_state(N, 3)
    s1 = s.copy()
    for o in ops:
        s1 = O.apply(s1, o)
    U = sp.identity(1 << N, dtype=complex, format="csc")
    for o in ops:
        U = o.expand(N) @ U
    assert np.allclose(s1, U @ s, atol=1e-12)


def test_state_vs_density_matrix_and_trajectory_average():
    """runtests.jl:89-133"""
    N, depth = 4, 20
    ops = random_ops(N, depth, 4)
    state = O.apply_ops(O.zero_state(N), ops)
    rho = O.to_rho(ops, N)
    assert np.allclose(np.outer(state, state.conj()), rho, atol=1e-10)
    assert abs(sum(O.expect(state, "T")) - sum(O.expect(rho, "T"))) < 1e-10
    nm = O.NoiseModel.model("depolarizing", 0.1)
    d = O.Draws(5)
    n_exp = 300
    mags = [sum(O.expect(O.apply_ops(O.zero_state(N), ops, noise=nm, draws=d), "T")) for _ in range(n_exp)]
    rho_n = O.to_rho(ops, N, noise=nm)
    assert abs(sum(O.expect(rho_n, "T")) - np.mean(mags)) < 0.25


def test_opqc_equals_noisemodel_on_density_matrix():
    """runtests.jl:136-166 (t1)"""
    N, depth, p = 1, 20, 0.01
    ops = random_ops(N, depth, 6)
    rs = rand_state(N, 7)
    rho1 = np.outer(rs, rs.conj())
    rho2 = rho1.copy()
    nm = O.NoiseModel.model("depolarizing", p)
    for o in ops:
        rho1 = O.apply(rho1, o)
        rho1 = O.apply(rho1, O.OpQC.model("depolarizing", p, o.qubit))
        rho2 = O.apply(rho2, o, noise=nm)
    assert np.linalg.norm(rho1 - rho2) < 1e-12


def test_mid_circuit_known_answers():
    """runtests.jl:195-231"""
    ops_reset = [O.Op("X", 1), O.RES(1), O.Op("MZ", 1)]
    ops_ifop = [O.Op("X", 1), O.ifOp("MZ", 1, [O.Op("I", 1)], [O.Op("X", 1)])]
    ops_mid = [O.Op("X", 1), O.Op("X", 2), O.Op("CX", 2, 3), O.Op("CX", 1, 2), O.Op("X", 2), O.Op("MZ", 2), O.RES(2), O.Op("CX", 2, 3), O.Op("CX", 1, 2),
               O.Op("CX", 2, 3), O.Op("X", 1)]
    d = O.Draws(0)
    assert O.run(ops_reset, 1, 1, draws=d)[0] == [0]
    assert O.run(ops_ifop, 1, 1, draws=d)[0] == [1]
    assert O.run(ops_mid, 3, 1, draws=d)[0] == [1]


def test_seed_locked_runs_are_reproducible():
    """runtests.jl:233-285: same seed => same outcomes and state (here: the oracle against itself and against the
    strided port, which plays the role the MPS backend plays in the reference's test)."""
    from oracle import strided as S

    ops = lambda: [O.Op("H", 1), O.Op("CX", 1, 2), O.Op("RY(0.37)", 1), O.Op("RZ(0.19)", 2), O.Op("MZ", 1), O.Op("X", 2), O.Op("MZ", 2)]
    nm = O.NoiseModel.model("depolarizing", 0.03)
    for seed in range(1, 41):
        s1, m1 = O.apply_ops(O.zero_state(2), ops(), noise=nm, draws=O.Draws(1000 + seed), track_measurements=True)
        sv, m2 = S.SV(2).apply_ops(ops(), noise=nm, draws=O.Draws(1000 + seed), track_measurements=True)
        assert m1 == m2
        assert np.linalg.norm(s1 - sv.v) < 1e-8


def test_gate_table_quirks():
    """values parity depends on."""
    assert O.GATE["T"][1, 1] == complex(0.7071067812, 0.7071067812)  # rounded to 10 significant digits
    assert np.allclose(O.gates("RXX(0.3)")[0, 0], math.cos(0.3))      # cos(phi), not cos(phi/2)
    assert np.allclose(O.noise_model("phase_flip", 0.2)[1], math.sqrt(0.2) * O.GATE["Y"])  # phase_flip yields Y
    assert np.allclose(O.gates("RZ(.1pi)"), np.diag([np.exp(-0.05j * math.pi), np.exp(0.05j * math.pi)]))
    assert np.allclose(O.gates("RX(0.5π)"), O._RX(0.5 * math.pi))
    K2 = O.noise_model("amplitude_damping", 0.3, two_qubit=True)
    K1 = O.noise_model("amplitude_damping", 0.3)
    assert np.allclose(K2[1], np.kron(K1[0], K1[1]))  # Ki outer, Kj inner (src/noise.jl:126)


def test_two_qubit_channel_quirk_when_qubit_gt_target():
    """probabilities come from the (min,max)-ordered RDM with the unswapped Kraus operators."""
    N = 4
    s = rand_state(N, 9)
    K = O.noise_model("amplitude_damping", 0.3, two_qubit=True)
    p32 = O.calc_prob(s, K, 3, 2)
    true = [float(np.sum(np.abs(O.hilbert2(N, k, 3, 2) @ s) ** 2)) for k in K]
    assert not np.allclose(p32, true)
    assert np.allclose(sorted(p32), sorted(true))  # same numbers, attached to the swapped partner
    assert np.allclose(O.calc_prob(s, K, 2, 3), [float(np.sum(np.abs(O.hilbert2(N, k, 2, 3) @ s) ** 2)) for k in K])


def test_sampling_contract():
    v = np.zeros(8, dtype=complex)
    v[[1, 4, 6]] = [0.6, 0.64j, 0.48]
    # u = 0 returns the first entry even if its weight is 0 (cumsum_1 >= 0), like a direct sampler's `cw < t` loop
    assert list(O.sample(v, [0.0, 1e-9, 0.3599, 0.3601, 0.77, 0.999])) == [0, 1, 1, 4, 6, 6]
    a, b = O.sample_exact(v)
    assert list(a) == [1, 4, 6]
    assert abs(O.sample_to_expectation(a, b, 3, [1]) - (0.36 - 0.4096 - 0.2304)) < 1e-15
