"""The strided C port (oracle/strided_cpu.c + oracle/strided.py) against the kron-chain restatement.  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import bt_oracle as O
from oracle import strided as S
from test_oracle_reference_kats import rand_state, random_ops


def test_single_gates_all_positions():
    N = 6
    for name in ["H", "T", "RX(0.3)", "SP"]:
        for q in range(1, N + 1):
            v = rand_state(N, q)
            sv = S.SV(N, v)
            sv.apply(O.Op(name, q))
            assert np.allclose(sv.v, O.Op(name, q).expand(N) @ v, atol=1e-13)
    for name in ["CX", "FSIM(0.2,0.1)", "ECR", "RZZ(0.4)"]:
        for q, t in itertools.permutations(range(1, N + 1), 2):
            v = rand_state(N, 10 * q + t)
            sv = S.SV(N, v)
            sv.apply(O.Op(name, q, t))
            assert np.allclose(sv.v, O.Op(name, q, t).expand(N) @ v, atol=1e-13)
    for q, c in itertools.permutations(range(1, N + 1), 2):
        v = rand_state(N, 3)
        sv = S.SV(N, v)
        sv.apply(O.Op("RY(0.7)", q, control=c))
        assert np.allclose(sv.v, O.Op("RY(0.7)", q, control=c).expand(N) @ v, atol=1e-13)


def test_partial_traces():
    N = 7
    v = rand_state(N, 1)
    sv = S.SV(N, v)
    for q in range(1, N + 1):
        assert np.allclose(sv.partial_trace([q]), O.partial_trace_1(v, q), atol=1e-13)
    for a, b in itertools.permutations(range(1, N + 1), 2):
        assert np.allclose(sv.partial_trace([a, b]), O.partial_trace_general(v, [a, b]), atol=1e-13)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_noisy_monitored_circuits_same_draws(seed):
    N, depth = 5, 6
    ops = random_ops(N, depth, seed, measure_prob=0.2)
    ops.insert(5, O.ifOp("MX", 3, [O.Op("I", 3)], [O.Op("X", 3), O.Op("H", 1)]))
    ops.insert(9, O.RES(2))
    ops.insert(11, O.OpQC.model("amplitude_damping", 0.3, 4, 2))
    nm = O.NoiseModel.model("amplitude_damping", 0.05)
    ref, mids = O.apply_ops(O.zero_state(N), ops, noise=nm, draws=O.Draws(seed), track_measurements=True)
    sv, mids2 = S.SV(N).apply_ops(ops, noise=nm, draws=O.Draws(seed), track_measurements=True)
    assert mids == mids2
    assert np.allclose(sv.v, ref, atol=1e-12)


def test_density_matrix_port():
    N, depth = 4, 8
    ops = random_ops(N, depth, 5)
    nm = O.NoiseModel.model("depolarizing", 0.1)
    ref = O.to_rho(ops, N, noise=nm)
    dm = S.DM(N)
    for o in ops:
        dm.apply(o, noise=nm)
    assert np.allclose(dm.to_matrix(), ref, atol=1e-12)
    dm.apply(O.OpQC.model("amplitude_damping", 0.2, 3, 1))
    assert np.allclose(dm.to_matrix(), O.apply(ref, O.OpQC.model("amplitude_damping", 0.2, 3, 1)), atol=1e-12)


def test_expect_z_all():
    v = rand_state(8, 2)
    assert np.allclose(S.SV(8, v).expect_z_all(), O.expect(v, "Z"), atol=1e-13)
