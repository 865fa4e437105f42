"""The committed golden fixtures are reproduced by the oracle (regression pin for oracle/ and the workload generators)."""
import os

import numpy as np

from oracle import bt_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_r1.npz"))


def test_c1_fixture(bt):
    import __graft_entry__ as ge
    from importlib import import_module

    wl = import_module(ge.PKG_NAME + ".workloads")
    st = O.apply_ops(O.zero_state(12), wl.to_ops(O, wl.c1_brickwork(12, 20, 12)))
    assert np.max(np.abs(st - G["c1_state"])) < 1e-14
    assert np.array_equal(O.sample(st, G["c1_uniforms"]), G["c1_samples"])
    assert abs(np.vdot(st, st).real - 1) < 1e-12


def test_monitored_fixture_from_logged_draws():
    from test_oracle_reference_kats import random_ops

    for seed in range(8):
        ops = random_ops(6, 5, 100 + seed, measure_prob=0.2)
        nm = O.NoiseModel.model(["depolarizing", "amplitude_damping"][seed % 2], 0.05)
        s, mids = O.apply_ops(O.zero_state(6), ops, noise=nm, draws=O.ListDraws(G[f"mon{seed}_draws"]), track_measurements=True)
        assert list(mids) == list(G[f"mon{seed}_mids"])
        assert np.max(np.abs(s - G[f"mon{seed}_state"])) < 1e-14


def test_oracle_shadow_converges_and_follows_the_draw_order():
    """src/ops.jl:145-185 restated: tr(rho) = 1 exactly; the estimate converges to |psi><psi|; the bases are exactly the
    integers drawn between the circuit and the shot (ListDraws replays them)."""
    import numpy as np
    from oracle import bt_oracle as O
    ops = [O.Op("H", 1), O.Op("CNOT", 1, 2), O.Op("RY(0.7)", 3)]
    psi = O.apply_ops(O.zero_state(3), ops)
    rho = O.shadow(ops, 3, 3000, draws=O.Draws(5))
    assert abs(np.trace(rho) - 1) < 1e-9
    assert abs(np.real(psi.conj() @ rho @ psi) - 1) < 0.1
    # one experiment, all qubits in the Z basis (index 2), shot u = 0.0 -> first basis state with support (|000>)
    one = O.shadow(ops, 3, 1, draws=O.ListDraws([0.0], ints=[2, 2, 2]))
    snap = np.ones((1, 1))
    for _ in range(3):
        snap = np.kron(snap, 3 * np.diag([1.0, 0.0]) - np.eye(2))
    assert np.max(np.abs(one - snap)) < 1e-12
