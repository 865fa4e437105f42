"""GPU parity for the noise -> measure/sample -> expect part of the path, through the C ABI, against the oracle;
includes the reference's own star tests (test/runtests.jl) restated for the device backend."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rand_state(N, seed):
    g = np.random.default_rng(seed)
    v = g.normal(size=1 << N) + 1j * g.normal(size=1 << N)
    return v / np.linalg.norm(v)


def random_ops(mod, N, depth, seed, measure_prob=0.0):
    """gates.jl:155-209 style generator (same structure, numpy RNG) producing ops for module ``mod``."""
    g = np.random.default_rng(seed)
    one = ["I", "X", "Y", "Z", "SX", "XSQRT", "H", "T", "S", "SD", "P", "U2", "U3"]
    two = ["CX", "CNOT", "CY", "CZ", "CP", "RXX", "RYY", "RZZ", "RXY", "GIVENS", "FSIM", "SWAP", "ISWAP", "FSWAP", "SYC", "ECR"]
    ops = []
    for _ in range(depth):
        c = 1
        while c <= N:
            pool = one if c == N else one + two
            s = pool[g.integers(len(pool))]
            p = [round(float(x) * np.pi, 2) for x in g.normal(size=3)]
            if s in one:
                if s == "P":
                    ops.append(mod.Op(f"P({p[0]})", c))
                elif s == "U2":
                    ops.append(mod.Op(f"U2({p[0]},{p[1]})", c))
                elif s == "U3":
                    ops.append(mod.Op(f"U3({p[0]},{p[1]},{p[2]})", c))
                else:
                    ops.append(mod.Op(s, c))
                c += 1
            else:
                if s == "FSIM":
                    ops.append(mod.Op(f"FSIM({p[0]},{p[1]})", c, c + 1))
                elif s in ("CP", "RXX", "RYY", "RZZ", "RXY", "GIVENS"):
                    ops.append(mod.Op(f"{s}({p[0]})", c, c + 1))
                else:
                    ops.append(mod.Op(s, c, c + 1))
                c += 2
        for i in range(1, N + 1):
            if g.random() < measure_prob:
                ops.append(mod.Op(["MX", "MY", "MZ"][g.integers(3)], i))
    return ops


# ---- partial traces ------------------------------------------------------------------------------------------
def test_partial_trace(bt, orc):
    N = 7
    v = rand_state(N, 3)
    s = bt.CuState.from_numpy(v)
    for q in range(1, N + 1):
        assert np.max(np.abs(bt.partial_trace(s, q) - orc.partial_trace_1(v, q))) < TOL
    for q in range(1, N):
        assert np.max(np.abs(bt.partial_trace(s, q, q + 1) - orc.partial_trace_2adj(v, q, q + 1))) < TOL
        assert np.max(np.abs(bt.partial_trace(s, q + 1, q) - orc.partial_trace_2adj(v, q + 1, q))) < TOL
    for a, b in itertools.permutations(range(1, N + 1), 2):
        assert np.max(np.abs(bt.partial_trace(s, [a, b]) - orc.partial_trace_general(v, [a, b]))) < TOL
    for q in range(1, N - 1):
        assert np.max(np.abs(bt.partial_trace(s, [q, q + 1, q + 2]) - orc.partial_trace_general(v, [q, q + 1, q + 2]))) < TOL
    for trip in [(1, 3, 6), (7, 2, 4), (5, 6, 1)]:
        assert np.max(np.abs(bt.partial_trace(s, list(trip)) - orc.partial_trace_general(v, list(trip)))) < TOL


def test_norm_inner_probs(bt, orc):
    N = 9
    a, b = rand_state(N, 1), rand_state(N, 2)
    sa, sb = bt.CuState.from_numpy(a), bt.CuState.from_numpy(b * 0.5)
    assert abs(bt.inner(sa, sb) - np.vdot(a, b * 0.5)) < TOL
    assert abs(bt.fidelity(sa, sb) - abs(np.vdot(a, 0.5 * b)) ** 2) < TOL
    assert abs(bt.norm2(sb) - 0.25) < TOL
    bt.normalize(sb)
    assert np.max(np.abs(sb.to_numpy() - b)) < TOL
    assert np.max(np.abs(bt.prob(sa) - np.abs(a) ** 2)) < 1e-15


# ---- measurement ---------------------------------------------------------------------------------------------
def test_born_measure_all_bases_same_draws(bt, orc):
    N = 6
    for seed in range(30):
        v = rand_state(N, seed)
        q = 1 + seed % N
        name = ["MZ", "MX", "MY", "MR"][seed % 4]
        s = bt.CuState.from_numpy(v)
        _, ind_d = bt.apply(s, bt.Op(name, q), rng=bt.Draws(seed), track_measurements=True)
        ref, ind_o = orc.apply(v, orc.Op(name, q), draws=orc.Draws(seed), track_measurements=True)
        assert ind_d == ind_o
        assert np.max(np.abs(s.to_numpy() - ref)) < TOL


def test_reference_mid_circuit_known_answers(bt):
    """test/runtests.jl:195-231."""
    ops_reset = [bt.Op("X", 1), bt.Op("RES", 1), bt.Op("MZ", 1)]
    ops_ifop = [bt.Op("X", 1), bt.ifOp("MZ", 1, bt.Op("I", 1), bt.Op("X", 1))]
    ops_mid = [bt.Op("X", 1), bt.Op("X", 2), bt.Op("CX", 2, 3), bt.Op("CX", 1, 2), bt.Op("X", 2), bt.Op("MZ", 2), bt.Op("RES", 2),
               bt.Op("CX", 2, 3), bt.Op("CX", 1, 2), bt.Op("CX", 2, 3), bt.Op("X", 1)]
    assert bt.run(ops_reset, 1)[0] == [0]
    assert bt.run(ops_ifop, 1)[0] == [1]
    assert bt.run(ops_mid, 1)[0] == [1]
    _, m = bt.apply(ops_reset, bt.zero_state(1), track_measurements=True)
    assert m == [0]
    _, m = bt.apply(ops_ifop, bt.zero_state(1), track_measurements=True)
    assert m == [1]
    _, m = bt.apply(ops_mid, bt.zero_state(3), track_measurements=True)
    assert m == [1]


def test_reference_seed_locked_noisy_run(bt, orc):
    """test/runtests.jl:233-259: the device backend as a third participant: identical mid-circuit outcomes and
    states under the same draws."""
    def ops(m):
        return [m.Op("H", 1), m.Op("CX", 1, 2), m.Op("RY(0.37)", 1), m.Op("RZ(0.19)", 2), m.Op("MZ", 1), m.Op("X", 2), m.Op("MZ", 2)]
    nm_d = bt.NoiseModel("depolarizing", 0.03)
    nm_o = orc.NoiseModel.model("depolarizing", 0.03)
    for seed in range(1, 41):
        sd, mid_d = bt.apply(ops(bt), bt.zero_state(2), noise=nm_d, rng=bt.Draws(1000 + seed), track_measurements=True)
        so, mid_o = orc.apply_ops(orc.zero_state(2), ops(orc), noise=nm_o, draws=orc.Draws(1000 + seed), track_measurements=True)
        assert mid_d == mid_o
        assert np.linalg.norm(sd.to_numpy() - so) < 1e-8


def test_reference_seed_locked_opqc(bt, orc):
    """test/runtests.jl:261-285."""
    def ops(m, QC):
        return [m.Op("H", 1), m.Op("CX", 1, 2), QC("depolarizing", 0.07, 1), QC("depolarizing", 0.04, 1, 2), m.Op("MZ", 1), m.Op("MZ", 2)]
    for seed in range(1, 41):
        sd, mid_d = bt.apply(ops(bt, bt.OpQC), bt.zero_state(2), rng=bt.Draws(2000 + seed), track_measurements=True)
        so, mid_o = orc.apply_ops(orc.zero_state(2), ops(orc, orc.OpQC.model), draws=orc.Draws(2000 + seed), track_measurements=True)
        assert mid_d == mid_o
        assert np.linalg.norm(sd.to_numpy() - so) < 1e-8


# ---- Kraus channels on state vectors -----------------------------------------------------------------------------
MODELS = ["amplitude_damping", "phase_damping", "phase_flip", "bit_flip", "bit_phase_flip", "depolarizing", "depolarizing_amp", "rot_x", "rot_y", "rot_z", "rot_p", "rot_xyz", "MZ"]


def test_kraus_probs_and_trajectory_step(bt, orc):
    N = 5
    k = 0
    for model in MODELS:
        if model == "rot_xyz":
            continue  # not CPTP for generic p: the reference's validity check rejects it
        for (q, t) in [(2, -1), (5, -1), (2, 3), (3, 2), (1, 4), (4, 1), (5, 3)]:
            k += 1
            v = rand_state(N, k)
            p = 0.3
            try:
                od = bt.OpQC(model, p, q, t)
            except ValueError:
                with pytest.raises(ValueError):
                    orc.OpQC.model(model, p, q, t)
                continue
            oo = orc.OpQC.model(model, p, q, t)
            s = bt.CuState.from_numpy(v)
            assert np.max(np.abs(od.prob(s) - np.array(oo.prob(v)))) < TOL, (model, q, t)
            bt.apply(s, od, rng=bt.Draws(k))
            ref = orc.apply(v, oo, draws=orc.Draws(k))
            assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (model, q, t)


def test_reference_kraus_prob_relation(bt, orc):
    """test/runtests.jl:19-35: tr(K pA K') == ||hilbert(N,K,q,t) psi||^2 == OpQC.prob."""
    N = 6
    v = rand_state(N, 11)
    s = bt.CuState.from_numpy(v)
    K = bt.noise_model("depolarizing", 0.5, two_qubit=True)
    pA = bt.partial_trace(s, [2, 3])
    probs = [np.real(np.trace(k @ pA @ k.conj().T)) for k in K]
    probs2 = []
    for k in K:
        c = s.copy()
        bt.apply(c, bt.Op("K", k, 2, 3))
        probs2.append(bt.norm2(c))
    probs3 = bt.OpQC("my quantum channel", K, 2, 3).prob(s)
    assert np.allclose(probs, probs2, atol=1e-12) and np.allclose(probs, probs3, atol=1e-12)


def test_kraus_3q(bt, orc):
    N = 5
    K1 = orc.noise_model("amplitude_damping", 0.3)
    K3 = [np.kron(np.kron(a, b), c) for a in K1 for b in K1 for c in K1]
    for first in (1, 2, 3):
        v = rand_state(N, first)
        s = bt.CuState.from_numpy(v)
        bt.apply(s, bt.OpQC("ad3", K3, first), rng=bt.Draws(first))
        ref = orc.apply(v, orc.OpQC("ad3", K3, first), draws=orc.Draws(first))
        assert np.max(np.abs(s.to_numpy() - ref)) < TOL


def test_weighted_sample_failure_is_loud(bt):
    """src/hilbert.jl:810-819 returns nothing when the draw exceeds the last cumulative sum -> error, not a silent pick."""
    import ctypes as C
    s = bt.zero_state(2)
    K = [0.5 * bt.gate["I"], 0.5 * bt.gate["X"]]  # not trace preserving on purpose: probabilities sum to 0.5
    tab = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in K])
    u = np.array([0.9])
    chosen = np.zeros(1, dtype=np.int32)
    rc = s.lib.bt_sv_kraus(s.h, 1, 1, -1, tab.ctypes.data_as(C.c_void_p), 2, u.ctypes.data_as(C.POINTER(C.c_double)), chosen.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == -4


# ---- noisy trajectories vs oracle on a random circuit with measurements -------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_noisy_monitored_circuit_same_draws(bt, orc, seed):
    N, depth = 5, 6
    od = random_ops(bt, N, depth, seed, measure_prob=0.15)
    oo = random_ops(orc, N, depth, seed, measure_prob=0.15)
    nm_d, nm_o = bt.NoiseModel("amplitude_damping", 0.05), orc.NoiseModel.model("amplitude_damping", 0.05)
    sd, mid_d = bt.apply(od, bt.zero_state(N), noise=nm_d, rng=bt.Draws(seed), track_measurements=True)
    so, mid_o = orc.apply_ops(orc.zero_state(N), oo, noise=nm_o, draws=orc.Draws(seed), track_measurements=True)
    assert mid_d == mid_o
    assert np.max(np.abs(sd.to_numpy() - so)) < TOL


# ---- circuits: gate by gate == circuit call (test/runtests.jl:54-73 model) -----------------------------------------
def test_circuit_call_equals_sequential_and_oracle(bt, orc):
    N, depth = 10, 50
    od, oo = random_ops(bt, N, depth, 7), random_ops(orc, N, depth, 7)
    v = rand_state(N, 5)
    s1 = bt.CuState.from_numpy(v)
    for o in od:
        bt.apply(s1, o)
    s2 = bt.CuState.from_numpy(v)
    bt.apply(od, s2)
    ref = orc.apply_ops(v, oo)
    assert np.max(np.abs(s1.to_numpy() - ref)) < TOL
    assert np.max(np.abs(s2.to_numpy() - ref)) < 1e-12 + TOL


# ---- expectation values, correlations, sampling ----------------------------------------------------------------------
def test_expect_and_correlation(bt, orc):
    N = 6
    v = rand_state(N, 21)
    s = bt.CuState.from_numpy(v)
    for name in ["Z", "X", "Y", "T", "H", "S", "P0", "SP", "RX(0.3)"]:
        assert np.max(np.abs(bt.expect(s, name) - np.array(orc.expect(v, name)))) < TOL, name
    for q in range(1, N + 1):
        assert abs(bt.expect(s, bt.Op("H", q)) - orc.expect(v, orc.Op("H", q))) < TOL
    assert abs(bt.expect(s, bt.Op("FSIM(0.2,0.3)", 2, 5)) - orc.expect(v, orc.Op("FSIM(0.2,0.3)", 2, 5))) < TOL
    for ops_s, qs in [("Z,Z", [2, 4]), ("X,Y", [1, 6]), ("Y,Y,Z", [3, 1, 5]), ("T,Z", [2, 3]), ("X,H,SP", [6, 2, 4]), ("Y", [4])]:
        assert abs(bt.correlation(s, ops_s, qs) - orc.correlation(v, ops_s, qs)) < TOL, ops_s
    assert abs(bt.correlation(s, [1, 3]) - orc.correlation_z(v, [1, 3])) < TOL


def test_sample_matches_inverse_cdf(bt, orc):
    for N in (3, 11, 14):
        v = rand_state(N, N)
        s = bt.CuState.from_numpy(v)
        us = np.random.default_rng(N).random(4096)
        assert np.array_equal(bt.sample(s, 4096, uniforms=us), orc.sample(v, us))
    # sparse state: only structurally non-zero entries are ever returned, sample_exact lists them ascending
    v = np.zeros(1 << 10, dtype=complex)
    v[[5, 77, 900]] = [0.6, 0.0 + 0.64j, 0.48]
    s = bt.CuState.from_numpy(v)
    us = np.concatenate([[0.0, 0.36, 0.3600001, 0.999999], np.random.default_rng(1).random(200)])
    assert np.array_equal(bt.sample(s, len(us), uniforms=us), orc.sample(v, us))
    a, b = bt.sample_exact(s)
    assert list(a) == [5, 77, 900] and np.allclose(b, [0.36, 0.4096, 0.2304])
    m = bt.measure(s)
    assert abs(m.expect[0] - orc.sample_to_expectation(a, b, 10, [1])) < 1e-15


# ---- density matrices ------------------------------------------------------------------------------------------------
def test_dm_unitaries_match_oracle_and_sv(bt, orc):
    """test/runtests.jl:89-107: state*state' == rho after the same ops."""
    N, depth = 4, 20
    od, oo = random_ops(bt, N, depth, 9), random_ops(orc, N, depth, 9)
    rho = bt.CuRho(N)
    st = bt.zero_state(N)
    for o in od:
        bt.apply(rho, o)
        bt.apply(st, o)
    ref = orc.to_rho(oo, N)
    assert np.max(np.abs(rho.to_numpy() - ref)) < TOL
    v = st.to_numpy()
    assert np.max(np.abs(np.outer(v, v.conj()) - rho.to_numpy())) < TOL
    assert np.max(np.abs(bt.CuRho.from_state(st).to_numpy() - rho.to_numpy())) < TOL
    assert np.max(np.abs(bt.expect(rho, "T") - np.array(orc.expect(ref, "T")))) < TOL
    assert np.max(np.abs(bt.expect(rho, "X") - np.array(orc.expect(ref, "X")))) < TOL


def test_dm_controlled_and_nonadjacent(bt, orc):
    N = 4
    g = np.random.default_rng(0)
    v = rand_state(N, 1)
    rho0 = np.outer(v, v.conj())
    cases = [("X", 3, -1, 1), ("H", 1, -1, 4), ("CX", 1, 3, -2), ("CZ", 4, 2, -2), ("FSIM(0.3,0.2)", 4, 1, -2), ("ECR", 2, 3, 4), ("RZZ(0.3)", 1, 4, -2), ("T", 2, -1, -2)]
    for name, q, t, c in cases:
        rho = bt.CuRho.from_numpy(rho0)
        bt.apply(rho, bt.Op(name, q, t, control=c))
        ref = orc.apply(rho0, orc.Op(name, q, t, control=c))
        assert np.max(np.abs(rho.to_numpy() - ref)) < TOL, (name, q, t, c)


def test_dm_channels_match_oracle(bt, orc):
    N = 4
    v = rand_state(N, 2)
    rho0 = np.outer(v, v.conj())
    for model in ["amplitude_damping", "phase_damping", "depolarizing", "bit_flip", "depolarizing_amp"]:
        for (q, t) in [(1, -1), (4, -1), (2, 3), (3, 2), (1, 4), (4, 2)]:
            rho = bt.CuRho.from_numpy(rho0)
            bt.apply(rho, bt.OpQC(model, 0.2, q, t))
            ref = orc.apply(rho0, orc.OpQC.model(model, 0.2, q, t))
            assert np.max(np.abs(rho.to_numpy() - ref)) < TOL, (model, q, t)
    # a non-product 2-qubit channel (correlated XX / ZZ errors) exercises the 16x16 superoperator kernel
    K = [np.sqrt(0.7) * np.eye(4), np.sqrt(0.2) * np.kron(orc.GATE["X"], orc.GATE["X"]), np.sqrt(0.1) * np.kron(orc.GATE["Z"], orc.GATE["Z"])]
    for (q, t) in [(1, 2), (4, 1), (3, 1)]:
        rho = bt.CuRho.from_numpy(rho0)
        bt.apply(rho, bt.OpQC("corr", K, q, t))
        ref = orc.apply(rho0, orc.OpQC("corr", K, q, t))
        assert np.max(np.abs(rho.to_numpy() - ref)) < TOL


def test_reference_opqc_equals_noisemodel_on_dm(bt, orc):
    """test/runtests.jl:136-166 (t1): OpQC after each op == NoiseModel on the density matrix."""
    N, depth, p = 1, 20, 0.01
    ops = random_ops(bt, N, depth, 4)
    v = rand_state(N, 8)
    rho1 = bt.CuRho.from_numpy(np.outer(v, v.conj()))
    rho2 = bt.CuRho.from_numpy(np.outer(v, v.conj()))
    nm = bt.NoiseModel("depolarizing", p)
    for o in ops:
        bt.apply(rho1, o)
        bt.apply(rho1, bt.OpQC("depolarizing", p, o.qubit))
        bt.apply(rho2, o, noise=nm)
    assert np.linalg.norm(rho1.to_numpy() - rho2.to_numpy()) < 1e-12


def test_dm_noisy_circuit_and_trajectory_average(bt, orc):
    """test/runtests.jl:109-131: noisy DM == oracle; average of noisy trajectories approaches it."""
    N, depth = 4, 8
    od, oo = random_ops(bt, N, depth, 13), random_ops(orc, N, depth, 13)
    nm_d, nm_o = bt.NoiseModel("depolarizing", 0.1), orc.NoiseModel.model("depolarizing", 0.1)
    rho = bt.CuRho(N)
    for o in od:
        bt.apply(rho, o, noise=nm_d)
    ref = orc.to_rho(oo, N, noise=nm_o)
    assert np.max(np.abs(rho.to_numpy() - ref)) < TOL
    exact = float(np.sum(bt.expect(rho, "T")))
    T = 2000
    st = bt.zero_state(N, T)
    U = np.random.default_rng(0).random((T, len(od) + 8))
    bt.apply(od, st, noise=nm_d, rng=bt.BatchDraws(U))
    avg = float(np.mean(np.sum(bt.expect(st, "T"), axis=1)))
    assert abs(avg - exact) < 0.1
    assert abs(rho.to_numpy().trace() - 1) < 1e-12


def test_dm_sample_diag_dephase_pauli(bt, orc):
    N = 5
    od, oo = random_ops(bt, N, 6, 3), random_ops(orc, N, 6, 3)
    nm_d, nm_o = bt.NoiseModel("amplitude_damping", 0.1), orc.NoiseModel.model("amplitude_damping", 0.1)
    rho = bt.CuRho(N)
    for o in od:
        bt.apply(rho, o, noise=nm_d)
    ref = orc.to_rho(oo, N, noise=nm_o)
    a, b = bt.sample_exact(rho)
    ao, bo = orc.sample_exact(ref)
    assert np.array_equal(a, ao) and np.max(np.abs(b - bo)) < TOL
    for ops_s, qs in [("Z,Z", [2, 4]), ("X,Y", [1, 5]), ("T,X", [2, 3])]:
        assert abs(bt.correlation(rho, ops_s, qs) - orc.correlation(ref, ops_s, qs)) < TOL
    bt.born_measure_Z(rho, 2)
    assert np.max(np.abs(rho.to_numpy() - orc.born_measure_Z_rho(N, ref, 2))) < TOL
    with pytest.raises(RuntimeError):
        bt.apply(rho, bt.Op("MZ", 1))  # src/hilbert.jl:772: unsupported in the reference, stays unsupported


# ---- batched trajectories ------------------------------------------------------------------------------------------
def test_batched_trajectories_equal_sequential_loop(bt, orc):
    """SURVEY C4 in miniature: every trajectory of the batch equals the oracle's sequential shot fed the same draws."""
    N, depth, T = 5, 6, 16
    od = random_ops(bt, N, depth, 5, measure_prob=0.2)
    oo = random_ops(orc, N, depth, 5, measure_prob=0.2)
    od.insert(7, bt.ifOp("MZ", 2, "I", "X"))
    oo.insert(7, orc.ifOp("MZ", 2, [orc.Op("I", 2)], [orc.Op("X", 2)]))
    M = sum(1 for o in od if o.type == "🔬")
    U = np.random.default_rng(3).random((T, M))
    outs = bt.run(od, T, batch=bt.BatchDraws(U))
    st = bt.zero_state(N, T)
    bt.apply(od, st, rng=bt.BatchDraws(U))
    got = st.to_numpy()
    for t in range(T):
        so, mid_o = orc.apply_ops(orc.zero_state(N), [o for o in oo], draws=orc.ListDraws(U[t]), track_measurements=True)
        assert outs[t] == mid_o
        assert np.max(np.abs(got[t] - so)) < TOL


def test_dm_fused_op_list_equals_op_by_op(bt, orc):
    """bt_dm_apply_ops (host-fused superoperators: gate + noise in one pass) == apply(rho, op; noise) op by op == oracle."""
    N, depth = 5, 6
    od, oo = random_ops(bt, N, depth, 21), random_ops(orc, N, depth, 21)
    od += [bt.Op("X", 2, control=4), bt.OpQC("amplitude_damping", 0.2, 3), bt.OpQC("depolarizing", 0.1, 4, 1), bt.Op("CX", 5, 1), bt.Op("FSIM(0.3,0.1)", 2, 3, control=5)]
    oo += [orc.Op("X", 2, control=4), orc.OpQC.model("amplitude_damping", 0.2, 3), orc.OpQC.model("depolarizing", 0.1, 4, 1), orc.Op("CX", 5, 1), orc.Op("FSIM(0.3,0.1)", 2, 3, control=5)]
    for nm_d, nm_o in ((False, False), (bt.NoiseModel("amplitude_damping", 0.07), orc.NoiseModel.model("amplitude_damping", 0.07))):
        fused = bt.CuRho(N)
        bt.apply(od, fused, noise=nm_d)
        seq = bt.CuRho(N)
        for o in od:
            bt.apply(seq, o, noise=nm_d)
        ref = orc.to_rho(oo, N, noise=nm_o)
        assert np.max(np.abs(seq.to_numpy() - ref)) < TOL
        assert np.max(np.abs(fused.to_numpy() - ref)) < TOL
        assert fused.launch_count() < seq.launch_count()


def test_two_qubit_born_measurement_sample_bit_entropy_hamiltonian(bt, orc):
    N = 5
    for seed, (q1, q2) in enumerate([(1, 2), (4, 2), (5, 1), (3, 4)]):
        v = rand_state(N, 40 + seed)
        s = bt.CuState.from_numpy(v)
        _, ind = bt.born_measure_Z2(s, q1, q2, rng=bt.Draws(seed))
        ref, ind_o = orc.born_measure_Z2(N, v, q1, q2, orc.Draws(seed))
        assert ind == ind_o
        assert np.max(np.abs(s.to_numpy() - ref)) < TOL
    v = rand_state(N, 50)
    s = bt.CuState.from_numpy(v)
    us = np.random.default_rng(0).random(20)
    bits = bt.sample_bit(s, 20, rng=type("R", (), {"uniform": lambda self, it=iter(us): float(next(it))})())
    assert bits == [orc.int2bin(int(a), N) for a in orc.sample(v, us)]
    for n in (4, 5, 6):
        w = rand_state(n, n)
        assert abs(bt.entanglement_entropy(bt.CuState.from_numpy(w)) - orc.entanglement_entropy(w)) < 1e-12
    terms = [(0.5, "Z,Z", [1, 2]), (-1.25, "X", [3]), (0.7, "X,Y,Z", [5, 1, 4])]
    assert abs(bt.hamiltonian_expect(s, terms) - sum(c * orc.correlation(v, o, q) for c, o, q in terms)) < TOL


def test_qasm_circuit_end_to_end(bt, orc):
    text = """OPENQASM 2.0; qreg q[5];
    h q[0]; cx q[0],q[3]; rz(0.3) q[1]; u3(0.1,0.2,0.3) q[2]; cu1(pi/4) q[4],q[0]; crz(0.5) q[0],q[2]; swap q[2],q[4];
    rzz(0.3) q[0],q[3]; fsim(0.1,0.2) q[1],q[2]; ccx q[0],q[1],q[2]; cswap q[4],q[1],q[3]; sdg q[1]; sx q[4];"""
    ops = bt.from_qasm(text)
    st = bt.apply(ops, bt.zero_state(5))
    ref = orc.zero_state(5)
    for o in ops:
        oo = orc.Op(o.name, o.qubit, o.target_qubit, control=o.control)
        if o.q == 2 and o.control != -2 and abs(o.qubit - o.target_qubit) != 1 and o.name not in ("CX", "CZ"):
            # controlled non-adjacent SWAP: the reference's expand throws (src/hilbert.jl:58-64); build it from its definition
            import scipy.sparse as sp
            P1, P0 = orc.hilbert1(5, orc.GATE["P1"], o.control), orc.hilbert1(5, orc.GATE["P0"], o.control)
            ref = (P1 @ orc.hilbert2(5, oo.mat, o.qubit, o.target_qubit) + P0) @ ref
        else:
            ref = oo.expand(5) @ ref
    assert np.max(np.abs(st.to_numpy() - ref)) < TOL


# ---- classical shadow (src/ops.jl:145-185) ---------------------------------------------------------
def test_shadow_matches_oracle_draw_for_draw(bt, orc):
    """same PCG64 stream on both sides: identical basis choices and shots, so the shadow estimates agree to rounding;
    noiseless circuit and a noisy one with a mid-circuit measurement (trajectory draws precede the basis draws)."""
    def ops(m):
        return [m.Op("H", 1), m.Op("CX", 1, 2), m.Op("RY(0.37)", 3), m.Op("CP(0.8)", 2, 3), m.Op("MZ", 2), m.Op("RX(1.1)", 1)]
    N, n_exp = 3, 60
    rd = bt.shadow(ops(bt), n_exp, rng=bt.Draws(42))
    ro = orc.shadow(ops(orc), N, n_exp, draws=orc.Draws(42))
    assert abs(np.trace(rd) - 1) < 1e-12
    assert np.max(np.abs(rd - ro)) < TOL
    nm_d, nm_o = bt.NoiseModel("amplitude_damping", 0.1), orc.NoiseModel.model("amplitude_damping", 0.1)
    circ = bt.compile(ops(bt), bt.Options(noise=nm_d))
    rd = bt.shadow(circ, n_exp, rng=bt.Draws(7))
    ro = orc.shadow(ops(orc), N, n_exp, noise=nm_o, draws=orc.Draws(7))
    assert np.max(np.abs(rd - ro)) < TOL


def test_opf_function_and_op_list_forms(bt, orc):
    """OpF (src/struct.jl:703-744, dispatched at src/hilbert.jl:486-487): an op list as one op, and a function of the
    device-resident state, alone and inside an op list with a tracked mid-circuit measurement."""
    N = 6
    v = rand_state(N, 21)
    block_d = random_ops(bt, N, 3, 77)
    block_o = random_ops(orc, N, 3, 77)
    s = bt.CuState.from_numpy(v)
    out = bt.apply(s, bt.OpF("block", block_d))
    assert out is s
    ref = orc.apply(v, orc.OpF("block", block_o))
    assert np.max(np.abs(s.to_numpy() - ref)) < TOL

    def flip_d(st):
        bt.apply(st, bt.Op("X", 2))            # in place, returns nothing

    def flip_o(st):
        return orc.apply(st, orc.Op("X", 2))

    ops_d = [bt.Op("H", 1), bt.OpF("flip", flip_d), bt.Op("CX", 1, 2), bt.Op("MZ", 2), bt.Op("RY(0.4)", 3)]
    ops_o = [orc.Op("H", 1), orc.OpF("flip", flip_o), orc.Op("CX", 1, 2), orc.Op("MZ", 2), orc.Op("RY(0.4)", 3)]
    s2 = bt.CuState.from_numpy(v)
    s2, mids_d = bt.apply(ops_d, s2, rng=bt.Draws(3), track_measurements=True)
    ref2, mids_o = orc.apply_ops(v, ops_o, draws=orc.Draws(3), track_measurements=True)
    assert mids_d == mids_o
    assert np.max(np.abs(s2.to_numpy() - ref2)) < TOL


def test_batched_ifop_branches_with_noise_and_nested_measurements(bt, orc):
    """ifOp on a batch whose trajectories measure different outcomes (src/struct.jl:578-594): the branch op lists run under a
    trajectory mask -- plain gates, a NoiseModel's Kraus draw after every noisy branch gate (apply(state, ifop; noise=noise), :587-590)
    and a measurement Op inside a branch.  Every trajectory must equal the oracle's sequential per-shot loop fed that trajectory's
    own uniform stream: a trajectory consumes draws only for the branch it took."""
    N, T = 5, 24
    noise_d, noise_o = bt.NoiseModel("amplitude_damping", 0.15), orc.NoiseModel.model("amplitude_damping", 0.15)

    def build(m):
        if0 = [m.Op("H", 3), m.Op("CX", 3, 4), m.Op("MX", 4)]
        if1 = [m.Op("X", 2), m.Op("RY(0.7)", 5), m.Op("MZ", 1), m.Op("CZ", 1, 5)]
        ops = [m.Op("H", q) for q in range(1, N + 1)] + [m.Op("CX", 1, 2), m.Op("RX(0.4)", 2), m.Op("FSIM(0.3,0.2)", 2, 3)]
        ops.append(m.ifOp("MZ", 2, if0, if1))
        ops += [m.Op("RY(1.1)", 4), m.Op("CX", 4, 5)]
        ops.append(m.ifOp("MY", 4, [m.Op("I", 4)], [m.Op("Z", 1), m.Op("H", 2)]))
        ops += [m.Op("MZ", 3), m.Op("T", 1)]
        return ops

    od, oo = build(bt), build(orc)
    U = np.random.default_rng(17).random((T, 64))  # more draws than any trajectory needs
    for noise_pair in ((False, False), (noise_d, noise_o)):
        st = bt.zero_state(N, T)
        _, mids = bt.apply(od, st, noise=noise_pair[0], rng=bt.BatchDraws(U), track_measurements=True)
        out = np.stack([np.asarray(m) for m in mids], axis=1)
        got = st.to_numpy()
        used = set()
        for t in range(T):
            dr = orc.ListDraws(U[t])
            so, mid_o = orc.apply_ops(orc.zero_state(N), oo, noise=noise_pair[1], draws=dr, track_measurements=True)
            assert list(out[t]) == mid_o, (t, out[t], mid_o)
            assert np.max(np.abs(got[t] - so)) < TOL, t
            used.add(len(dr.log))
        assert out.shape == (T, 3) and set(np.unique(out[:, 0])) == {0, 1}   # both branches were taken inside the batch
        if noise_pair[0] is not False:
            assert len(used) > 1  # the branches draw different numbers of uniforms: per-trajectory accounting matters


def test_batched_sampling_is_one_segmented_pass_per_trajectory_cdf(bt, orc):
    """bt_sv_sample_batched: every trajectory's shots follow its own inverse CDF (src/ops.jl:46-62 per trajectory, the per-shot loop
    of src/ops.jl:616-631 with shots = 1 included); three launches for the whole batch, whatever its size."""
    rng = np.random.default_rng(77)
    for N, nb, shots in ((3, 5, 7), (13, 9, 33), (14, 40, 1)):
        v = rng.normal(size=(nb, 1 << N)) + 1j * rng.normal(size=(nb, 1 << N))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v[1] = 0
        v[1, (1 << N) - 1] = 1.0  # a basis state: every draw returns the last index
        s = bt.CuState.from_numpy(v)
        us = rng.random((nb, shots))
        us[0, 0] = 0.0
        l0 = s.launch_count()
        got = bt.sample(s, shots, uniforms=us)
        assert s.launch_count() - l0 == 3
        assert got.shape == (nb, shots)
        for t in range(nb):
            assert np.array_equal(got[t], orc.sample(v[t], us[t])), t
        assert np.all(got[1] == (1 << N) - 1)


def test_multi_measurement_equals_sequential_born_measurements(bt, orc):
    """bt_sv_measure_z_multi: k measurements / resets in one read + one collapse pass give the outcomes, the collapsed state and
    the draw consumption of k sequential born_measure_Z / _reset_Z calls (src/hilbert.jl:682-696, :752-759), on single states,
    batches, under a trajectory mask, and through apply() on an op list (oracle fed the same draws)."""
    import ctypes as C
    L = bt._lib
    rng = np.random.default_rng(2024)
    for N, nb, qs, rs in ((5, 1, [3], [0]), (6, 1, [2, 5], [0, 0]), (7, 3, [7, 1, 4], [0, 1, 0]), (9, 4, [2, 9, 5, 1], [1, 0, 0, 1]), (4, 6, [4, 3, 2, 1], [0, 0, 0, 0])):
        v = rng.normal(size=(nb, 1 << N)) + 1j * rng.normal(size=(nb, 1 << N))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        k = len(qs)
        u = rng.random((nb, k))
        a = bt.CuState.from_numpy(v if nb > 1 else v[0])
        b = bt.CuState.from_numpy(v if nb > 1 else v[0])
        out = np.empty((nb, k), dtype=np.int32)
        L.check(a.lib.bt_sv_measure_z_multi(a.h, k, (C.c_int * k)(*qs), L.pdouble(np.ascontiguousarray(u)), out.ctypes.data_as(C.POINTER(C.c_int32)), (C.c_int * k)(*rs)))
        seq = np.empty((nb, k), dtype=np.int32)
        for j in range(k):
            o = np.empty(nb, dtype=np.int32)
            L.check(b.lib.bt_sv_measure_z(b.h, qs[j], L.pdouble(np.ascontiguousarray(u[:, j])), o.ctypes.data_as(C.POINTER(C.c_int32)), None, rs[j]))
            seq[:, j] = o
        assert np.array_equal(out, seq), (N, qs)
        assert np.max(np.abs(a.to_numpy() - b.to_numpy())) < 1e-13
        assert np.max(np.abs(np.atleast_1d(bt.norm2(a)) - 1)) < 1e-12
    # under a mask: untouched trajectories, outcome -1
    v = rng.normal(size=(4, 64)) + 1j * rng.normal(size=(4, 64))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    a = bt.CuState.from_numpy(v)
    a.set_mask(np.array([True, False, True, False]))
    out = np.empty((4, 2), dtype=np.int32)
    L.check(a.lib.bt_sv_measure_z_multi(a.h, 2, (C.c_int * 2)(2, 6), L.pdouble(np.ascontiguousarray(rng.random((4, 2)))), out.ctypes.data_as(C.POINTER(C.c_int32)), None))
    a.set_mask(None)
    got = a.to_numpy()
    assert np.all(out[[1, 3]] == -1) and np.all(out[[0, 2]] >= 0)
    assert np.array_equal(got[1], v[1]) and np.array_equal(got[3], v[3])
    # through apply(): a monitored layer with consecutive measurements and a reset, against the oracle's op-by-op loop
    N = 7
    names = [("H", 1), ("H", 4), ("CNOT", 1, 2), ("CNOT", 4, 5), ("RY(0.7)", 6), ("MZ", 2), ("MZ", 5), ("MZ", 6), ("MZ", 1), ("MZ", 7), ("H", 3), ("CNOT", 3, 2), ("MZ", 2), ("MZ", 3)]
    od = [bt.Op(*t) for t in names]
    oo = [orc.Op(*t) for t in names]
    od.insert(10, bt.Op("RES", 4)); oo.insert(10, orc.RES(4))
    for seed in range(6):
        st, mid = bt.apply(od, bt.zero_state(N), rng=bt.Draws(seed), track_measurements=True)
        ref, mid_o = orc.apply_ops(orc.zero_state(N), oo, draws=orc.Draws(seed), track_measurements=True)
        assert mid == mid_o, (seed, mid, mid_o)
        assert np.max(np.abs(st.to_numpy() - ref)) < 1e-12
    import bluetangle_jl_b200.host as H
    groups = H._coalesce(od, bt.zero_state(N), False)
    assert sum(isinstance(g, H._MeasureRun) for g in groups) == 3  # [2,5,6,1] + [7,RES 4] + [2,3]


def test_same_up_to_global_phase(bt):
    """src/linalg.jl:67-72 on device states."""
    g = np.random.default_rng(4)
    v = g.normal(size=256) + 1j * g.normal(size=256)
    v /= np.linalg.norm(v)
    a, b = bt.CuState.from_numpy(v), bt.CuState.from_numpy(np.exp(0.7j) * v)
    same, ph = bt.same_up_to_global_phase(b, a)
    assert same and abs(ph - 0.7) < 1e-12
    w = g.normal(size=256) + 1j * g.normal(size=256)
    assert not bt.same_up_to_global_phase(bt.CuState.from_numpy(w / np.linalg.norm(w)), a)[0]


def test_adjoint_circuit_round_trip(bt):
    """apply(adjoint(ops), apply(ops, state)) returns the state (src/linalg.jl:33-45), fused path at 16 qubits."""
    N = 16
    g = np.random.default_rng(8)
    ops = []
    for l in range(6):
        for q in range(1, N + 1):
            ops.append(bt.Op(["H", "T", f"RX({g.uniform(0, 6)!r})", f"RZ({g.uniform(0, 6)!r})", "S"][int(g.integers(5))], q))
        for q in range(1 + l % 2, N, 2):
            ops.append(bt.Op(["CNOT", "CZ", f"CP({g.uniform(0, 6)!r})", "ISWAP"][int(g.integers(4))], q, q + 1))
    s = bt.basis_state(N, 12345)
    bt.apply(ops, s)
    assert abs(bt.expect(s, "Z")).min() < 0.9  # the circuit moved the state away from the basis state
    bt.apply(bt.adjoint(ops), s)
    v = s.to_numpy()
    assert abs(abs(v[12345]) - 1) < 1e-7  # T is rounded to 10 digits in the reference's table: T T† differs from 1 by 1e-10 per pair
    assert np.max(np.abs(np.delete(v, 12345))) < 1e-9
