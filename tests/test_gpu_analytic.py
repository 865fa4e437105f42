"""Known answers with closed forms -- independent of oracle/ (no code shared with the restatement of the reference): GHZ,
QFT of a basis state, single-qubit channels on |+>, Bell-pair reduced density matrices, the shots = 1 sampling path of the
trajectory loop (src/ops.jl:623).  All through the C ABI (ctypes), tolerance 1e-12 .. 1e-10 as stated per assertion."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mods():
    import __graft_entry__ as ge
    from importlib import import_module

    return import_module(ge.PKG_NAME + ".workloads")


@pytest.mark.parametrize("N", [3, 12, 21])
def test_ghz(bt, N):
    """H(1), CX(1,2), CX(2,3), ...: (|0..0> + |1..1>)/sqrt(2); every <Z_q> = 0, <ZZ> = 1, <X..X> = 1; measuring qubit 1 fixes all."""
    ops = [bt.Op("H", 1)] + [bt.Op("CX", q, q + 1) for q in range(1, N)]
    for fused in (True, False):
        s = bt.zero_state(N)
        if fused:
            bt.apply(ops, s)
        else:
            for o in ops:
                bt.apply(s, o)
        a = s.to_numpy()
        want = np.zeros(1 << N, dtype=np.complex128)
        want[0] = want[-1] = 1 / math.sqrt(2)
        assert np.max(np.abs(a - want)) < 1e-14
        assert np.max(np.abs(bt.expect(s, "Z"))) < 1e-14
        assert abs(bt.correlation(s, "Z,Z", [1, N]) - 1) < 1e-14
        assert abs(bt.correlation(s, ",".join("X" * N), list(range(1, N + 1))) - 1) < 1e-13
        idx = bt.sample(s, 6, uniforms=np.array([0.0, 0.25, 0.4999, 0.5001, 0.75, 0.999999]))
        assert list(idx) == [0, 0, 0, (1 << N) - 1, (1 << N) - 1, (1 << N) - 1]
        s, out = bt.born_measure_Z(s, N, rng=bt.Draws(np.random.Generator(np.random.PCG64(1))))
        a = s.to_numpy()
        assert abs(abs(a[(1 << N) - 1 if out else 0]) - 1) < 1e-14 and abs(bt.norm2(s) - 1) < 1e-14


@pytest.mark.parametrize("N,x", [(4, 5), (13, 0x1234 & 0x1FFF), (22, 0x2A5A5A)])
def test_qft_of_a_basis_state_closed_form(bt, N, x):
    """The QFT of workloads.qft (H(q), then CP(pi/2^(j-q)) controlled by j > q; no final swaps) maps |x> to the product state
    prod_q (|0> + exp(2 pi i phi_q)|1>)/sqrt(2) with phi_q = sum_{j>=q} x_j 2^-(j-q+1): amplitude of y = 2^(-N/2) exp(2 pi i sum_q y_q phi_q)."""
    wl = mods()
    s = bt.basis_state(N, x)
    bt.apply(wl.to_ops(bt, wl.qft(N)), s)
    got = s.to_numpy()
    xb = [(x >> (N - q)) & 1 for q in range(1, N + 1)]
    phase = np.zeros(1 << N)
    y = np.arange(1 << N, dtype=np.int64)
    for q in range(1, N + 1):
        phi = sum(xb[j - 1] * 2.0 ** -(j - q + 1) for j in range(q, N + 1))
        phase += ((y >> (N - q)) & 1) * phi
    want = np.exp(2j * math.pi * phase) / math.sqrt(1 << N)
    assert np.max(np.abs(got - want)) < 1e-12


def test_single_qubit_channels_on_plus_closed_form(bt):
    """rho = |+><+| on one qubit of a 3-qubit register (others |0>): depolarizing(p) -> (I + (1 - 4p/3) X)/2;
    amplitude_damping(g) -> [[(1+g)/2, sqrt(1-g)/2], [sqrt(1-g)/2, (1-g)/2]]; bit_flip(p) leaves |+> alone; phase_damping(g)
    scales the coherence by sqrt(1-g).  src/noise.jl:66-118 + src/struct.jl:58-66."""
    N = 3
    for q in (1, 2, 3):
        cases = [("depolarizing", 0.3, np.array([[0.5, 0.5 * (1 - 0.4)], [0.5 * (1 - 0.4), 0.5]])),
                 ("amplitude_damping", 0.36, np.array([[0.68, 0.4], [0.4, 0.32]])),
                 ("bit_flip", 0.2, np.array([[0.5, 0.5], [0.5, 0.5]])),
                 ("phase_damping", 0.19, np.array([[0.5, 0.45], [0.45, 0.5]]))]
        for model, p, want1 in cases:
            rho = bt.CuRho(N)
            bt.apply(rho, bt.Op("H", q))
            bt.apply(rho, bt.OpQC(model, p, q))
            full = np.array([[1.0]])
            for k in range(1, N + 1):
                full = np.kron(full, want1 if k == q else np.array([[1.0, 0.0], [0.0, 0.0]]))
            assert np.max(np.abs(rho.to_numpy() - full)) < 1e-14, (model, q)
    # trajectory step on the state vector: amplitude damping on |+> keeps K0 with probability (2 - g)/2, the state becomes (1, sqrt(1-g))/sqrt(2-g)
    g = 0.36
    s = bt.plus_state(1)
    K = bt.noise_model("amplitude_damping", g)
    probs = bt.OpQC("amplitude_damping", g, 1).prob(s) if hasattr(bt.OpQC("amplitude_damping", g, 1), "prob") else None
    if probs is not None:
        assert np.max(np.abs(np.asarray(probs) - np.array([(2 - g) / 2, g / 2]))) < 1e-14

    class One:
        def __init__(self, u):
            self.u = u

        def uniform(self):
            return self.u

        def randint(self, n):
            return 0

    bt.apply(s, bt.OpQC("amplitude_damping", g, 1), rng=One((2 - g) / 2 - 1e-9))
    assert np.max(np.abs(s.to_numpy() - np.array([1.0, math.sqrt(1 - g)]) / math.sqrt(2 - g))) < 1e-14
    s = bt.plus_state(1)
    bt.apply(s, bt.OpQC("amplitude_damping", g, 1), rng=One((2 - g) / 2 + 1e-9))  # first i with u <= cumsum: K1 -> |0>
    assert np.max(np.abs(s.to_numpy() - np.array([1.0, 0.0]))) < 1e-14
    del K


def test_bell_pair_reduced_density_matrices(bt):
    """(|00> + |11>)/sqrt(2) on qubits (2, 5) of 6, (|01> - |10>)/sqrt(2) on (3, 4): partial_trace of the pair = the Bell projector,
    of one member = I/2, of an unentangled qubit = |0><0| (src/linalg.jl:167-230, :83-140)."""
    N = 6
    s = bt.zero_state(N)
    bt.apply([bt.Op("H", 2), bt.Op("CX", 2, 5), bt.Op("X", 3), bt.Op("H", 3), bt.Op("X", 4), bt.Op("CX", 3, 4)], s)
    phi = np.zeros((4, 4))
    phi[0, 0] = phi[0, 3] = phi[3, 0] = phi[3, 3] = 0.5
    psi = np.zeros((4, 4))
    psi[1, 1] = psi[2, 2] = 0.5
    psi[1, 2] = psi[2, 1] = -0.5
    with pytest.raises(ValueError):
        bt.partial_trace(s, 2, 5)  # the two-argument form is adjacent-only, like src/linalg.jl:198-201
    assert np.max(np.abs(bt.partial_trace(s, [2, 5]) - phi)) < 1e-14
    assert np.max(np.abs(bt.partial_trace(s, 3, 4) - psi)) < 1e-14
    for q in (2, 3, 4, 5):
        assert np.max(np.abs(bt.partial_trace(s, q) - np.eye(2) / 2)) < 1e-14
    for q in (1, 6):
        assert np.max(np.abs(bt.partial_trace(s, q) - np.diag([1.0, 0.0]))) < 1e-14
    # a pair that straddles the two Bell pairs is maximally mixed
    assert np.max(np.abs(bt.partial_trace(s, 2, 3) - np.eye(4) / 4)) < 1e-14


def test_shots_equal_one_sampling_path_of_the_trajectory_loop(bt):
    """The per-shot loop draws ONE sample from each final state (src/ops.jl:623: sample(state, 1)).  Contract (DESIGN.md section 2,
    "sample"): with one uniform u the index is the first i with cumsum(p)_i >= u * sum(p) -- StatsBase's direct_sample! for a single
    draw.  Checked on a state whose cumulative distribution is known in closed form: p = (1/2, 1/4, 1/8, 1/8) on indices 0, 5, 10, 15."""
    amp = np.zeros(16, dtype=np.complex128)
    amp[0], amp[5], amp[10], amp[15] = math.sqrt(0.5), 0.5j, -math.sqrt(0.125), math.sqrt(0.125) * 1j
    s = bt.CuState.from_numpy(amp)
    for u, want in ((0.0, 0), (0.499999, 0), (0.5, 0), (0.500001, 5), (0.75, 5), (0.7500001, 10), (0.875, 10), (0.8750001, 15), (0.9999999, 15)):
        assert int(bt.sample(s, 1, uniforms=np.array([u]))[0]) == want, u
    # unnormalised weights (Weights(probs) need not sum to 1): t = u * sum(p)
    s2 = bt.CuState.from_numpy(3.0 * amp)
    assert int(bt.sample(s2, 1, uniforms=np.array([0.74]))[0]) == 5 and int(bt.sample(s2, 1, uniforms=np.array([0.76]))[0]) == 10
