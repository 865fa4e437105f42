"""GPU parity: single gates through the C ABI vs the kron-chain oracle (restating src/hilbert.jl:18-159)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10  # north_star: amplitudes agree to <= 1e-10 absolute


def rand_state(N, seed):
    g = np.random.default_rng(seed)
    v = g.normal(size=1 << N) + 1j * g.normal(size=1 << N)
    return v / np.linalg.norm(v)


def rand_unitary(d, seed):
    g = np.random.default_rng(seed)
    q, r = np.linalg.qr(g.normal(size=(d, d)) + 1j * g.normal(size=(d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


ONEQ = ["I", "X", "Y", "Z", "SX", "H", "S", "SD", "T", "TD", "HSP", "HY", "P0", "P1", "SP", "SM", "RX(0.3)", "RY(1.1)", "RZ(.7pi)", "P(0.4)", "U2(0.1,0.2)", "U3(0.1,0.2,0.3)"]
TWOQ = ["CX", "CY", "CZ", "CS", "CT", "CTD", "CSD", "CI", "CH", "SWAP", "ISWAP", "FSWAP", "SYC", "ECR", "H2", "CP(0.3)", "GIVENS(0.2)", "FSIM(0.1,0.2)", "SWAPA(0.3)", "RXX(0.4)", "RYY(0.5)", "RZZ(0.6)", "RXY(0.7)"]


@pytest.mark.parametrize("N", [1, 2, 5, 9])
def test_1q_all_positions(bt, orc, N):
    for i, name in enumerate(ONEQ):
        for q in range(1, N + 1):
            v = rand_state(N, 100 * i + q)
            s = bt.CuState.from_numpy(v)
            bt.apply(s, bt.Op(name, q))
            ref = orc.Op(name, q).expand(N) @ v
            assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (name, q)


@pytest.mark.parametrize("N", [2, 6])
def test_2q_all_pairs(bt, orc, N):
    for i, name in enumerate(TWOQ):
        for q, t in itertools.permutations(range(1, N + 1), 2):
            v = rand_state(N, 1000 * i + 10 * q + t)
            s = bt.CuState.from_numpy(v)
            bt.apply(s, bt.Op(name, q, t))
            ref = orc.Op(name, q, t).expand(N) @ v
            assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (name, q, t)


def test_random_dense_2q_any_pair(bt, orc):
    N = 7
    for k, (q, t) in enumerate(itertools.permutations(range(1, N + 1), 2)):
        U = rand_unitary(4, k)
        v = rand_state(N, k)
        s = bt.CuState.from_numpy(v)
        bt.apply(s, bt.Op("U", U, q, t))
        ref = orc.Op("U", q, t, mat=U).expand(N) @ v
        assert np.max(np.abs(s.to_numpy() - ref)) < TOL


def test_controlled_1q(bt, orc):
    N = 6
    for name in ["X", "Y", "Z", "H", "T", "RX(0.3)"]:
        for q, c in itertools.permutations(range(1, N + 1), 2):
            v = rand_state(N, q * 7 + c)
            s = bt.CuState.from_numpy(v)
            bt.apply(s, bt.Op(name, q, control=c))
            ref = orc.Op(name, q, control=c).expand(N) @ v
            assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (name, q, c)


def test_controlled_2q_adjacent_and_ccx_ccz(bt, orc):
    N = 6
    # adjacent pairs with any control: src/hilbert.jl:52-54
    for name in ["CX", "CZ", "SWAP", "FSIM(0.1,0.2)"]:
        for q in range(1, N):
            for t in (q + 1,):
                for c in range(1, N + 1):
                    if c in (q, t):
                        continue
                    for (a, b) in ((q, t), (t, q)):
                        v = rand_state(N, a * 31 + b * 7 + c)
                        s = bt.CuState.from_numpy(v)
                        bt.apply(s, bt.Op(name, a, b, control=c))
                        ref = orc.Op(name, a, b, control=c).expand(N) @ v
                        assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (name, a, b, c)
    # non-adjacent: CCZX src/hilbert.jl:73-103 through the three-qubit constructor src/struct.jl:436-449
    for name in ["CCX", "CCZ"]:
        for q, c, t in itertools.permutations(range(1, N + 1), 3):
            if abs(q - t) == 1:
                continue
            v = rand_state(N, q * 100 + c * 10 + t)
            s = bt.CuState.from_numpy(v)
            bt.apply(s, bt.Op(name, q, c, t))
            ref = orc.Op3(name, q, c, t).expand(N) @ v
            assert np.max(np.abs(s.to_numpy() - ref)) < TOL, (name, q, c, t)


def test_3q_kraus_style_dense(bt, orc):
    import ctypes as C

    N = 6
    L = bt._lib
    for first in range(1, N - 1):
        U = rand_unitary(8, first)
        v = rand_state(N, first)
        s = bt.CuState.from_numpy(v)
        L.check(s.lib.bt_sv_apply_3q(s.h, first, L.ptr(L.cmat(U, 8))))
        ref = orc.hilbert3(N, U, first) @ v
        assert np.max(np.abs(s.to_numpy() - ref)) < TOL


def test_error_behaviour_mirrors_reference(bt):
    s = bt.zero_state(3)
    with pytest.raises(bt._lib.BTError):
        bt.apply(s, bt.Op("X", 4))  # N must be larger than qubit (src/hilbert.jl:145)
    with pytest.raises(ValueError):
        bt.Op("CX", 2, 2)  # src/struct.jl:465
    with pytest.raises(ValueError):
        bt.Op("X", 2, control=2)  # src/struct.jl:458
    with pytest.raises(bt._lib.BTError):
        bt.apply(s, bt.Op("CX", 1, 5))


def test_states(bt, orc):
    for N in (1, 3, 8):
        assert np.array_equal(bt.zero_state(N).to_numpy(), orc.zero_state(N))
        assert np.array_equal(bt.one_state(N).to_numpy(), orc.one_state(N))
        assert np.allclose(bt.plus_state(N).to_numpy(), orc.plus_state(N), atol=1e-16)
        assert np.array_equal(bt.neel_state01(N).to_numpy(), orc.neel_state01(N))
        assert np.array_equal(bt.neel_state10(N).to_numpy(), orc.neel_state10(N))
    assert np.array_equal(bt.product_state([1, 0, 1, 1]).to_numpy(), orc.product_state([1, 0, 1, 1]))
