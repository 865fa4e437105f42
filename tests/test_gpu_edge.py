"""Edge cases through the C ABI: tiny states, every tile size below the default, ragged batches, empty circuits,
fused clusters on all qubit arrangements -- always against the kron-chain oracle."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rand_state(N, seed, batch=None):
    g = np.random.default_rng(seed)
    shape = (1 << N,) if batch is None else (batch, 1 << N)
    v = g.normal(size=shape) + 1j * g.normal(size=shape)
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def rand_unitary(d, g):
    q, r = np.linalg.qr(g.normal(size=(d, d)) + 1j * g.normal(size=(d, d)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_dense_circuit(bt, orc, N, n_ops, seed):
    """dense random 1q/2q unitaries, diagonals and controlled gates on random (also non-adjacent) qubits"""
    g = np.random.default_rng(seed)
    od, oo = [], []
    for _ in range(n_ops):
        kind = g.integers(5) if N >= 2 else 0
        if kind == 0:
            q = int(g.integers(1, N + 1)); U = rand_unitary(2, g)
            od.append(bt.Op("U", U, q)); oo.append(orc.Op("U", q, mat=U))
        elif kind in (1, 2):
            q, t = (int(x) + 1 for x in g.choice(N, 2, replace=False)); U = rand_unitary(4, g)
            od.append(bt.Op("U", U, q, t)); oo.append(orc.Op("U", q, t, mat=U))
        elif kind == 3:
            q, t = (int(x) + 1 for x in g.choice(N, 2, replace=False)); th = float(g.uniform(0, 6))
            name = ["CZ", f"CP({th})", f"RZZ({th})", "CX"][int(g.integers(4))]
            od.append(bt.Op(name, q, t)); oo.append(orc.Op(name, q, t))
        else:
            q, c = (int(x) + 1 for x in g.choice(N, 2, replace=False)); th = float(g.uniform(0, 6))
            name = ["X", "H", f"RY({th})", "T", f"RZ({th})"][int(g.integers(5))]
            od.append(bt.Op(name, q, control=c)); oo.append(orc.Op(name, q, control=c))
    return od, oo


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 9, 11, 12, 13])
def test_fused_circuits_every_size(bt, orc, N):
    od, oo = random_dense_circuit(bt, orc, N, 40, N)
    v = rand_state(N, N)
    s = bt.CuState.from_numpy(v)
    bt.apply(od, s)                       # one bt_sv_apply_circuit call: fusion + tile kernel (+ clusters when T >= 4)
    ref = orc.apply_ops(v, oo)
    assert np.max(np.abs(s.to_numpy() - ref)) < TOL


@pytest.mark.parametrize("batch", [2, 3, 5, 17])
def test_ragged_batches(bt, orc, batch):
    N = 6
    od, oo = random_dense_circuit(bt, orc, N, 25, 100 + batch)
    v = rand_state(N, batch, batch)
    s = bt.CuState.from_numpy(v)
    bt.apply(od, s)
    for o in od[:5]:
        bt.apply(s, o)                    # single-gate kernels on a ragged batch too
    got = s.to_numpy()
    for t in range(batch):
        ref = orc.apply_ops(v[t], oo + oo[:5])
        assert np.max(np.abs(got[t] - ref)) < TOL
    nrm = bt.norm2(s)
    assert nrm.shape == (batch,) and np.max(np.abs(nrm - 1)) < 1e-9  # T is stored rounded to 10 digits (src/gates.jl:32)
    ez = bt.expect(s, "Z")
    assert ez.shape == (batch, N)
    assert np.max(np.abs(ez[batch - 1] - np.array(orc.expect(orc.apply_ops(v[batch - 1], oo + oo[:5]), "Z")))) < TOL


def test_empty_and_identity_circuits(bt):
    s = bt.plus_state(5)
    before = s.to_numpy().copy()
    n0 = s.launch_count()
    bt.apply([], s)
    bt.apply([bt.Op("I", 2), bt.Op("CI", 1, 4)], s)
    assert np.array_equal(s.to_numpy(), before)
    assert s.launch_count() == n0        # identities never touch HBM
    import ctypes as C
    assert s.lib.bt_sv_apply_circuit(s.h, None, 0, 1) == 0


def test_brickwork_clusters_all_alignments(bt, orc):
    """the register-cluster path: brickwork with every offset / window so that V, Lambda and chain shapes all occur"""
    N = 10
    g = np.random.default_rng(3)
    for offset in range(3):
        od, oo = [], []
        for layer in range(6):
            for q in range(1 + (layer + offset) % 2, N, 2):
                U = rand_unitary(4, g)
                a, b = (q, q + 1) if (layer + q) % 3 else (q + 1, q)
                od.append(bt.Op("U", U, a, b)); oo.append(orc.Op("U", a, b, mat=U))
        v = rand_state(N, offset)
        s = bt.CuState.from_numpy(v)
        bt.apply(od, s)
        assert np.max(np.abs(s.to_numpy() - orc.apply_ops(v, oo))) < TOL


def test_three_qubit_and_kraus_ops_inside_op_lists(bt, orc):
    N = 6
    K1 = orc.noise_model("amplitude_damping", 0.3)
    K3 = [np.kron(np.kron(a, b), c) for a in K1 for b in K1 for c in K1]
    od = [bt.Op("H", 1), bt.Op("CCX", 1, 3, 5), bt.Op("CX", 2, 3), bt.OpQC("ad3", K3, 2), bt.Op("CCZ", 6, 2, 4), bt.Op("FSIM(0.1,0.3)", 4, 1, control=6), bt.Op("MY", 3)]
    oo = [orc.Op("H", 1), orc.Op3("CCX", 1, 3, 5), orc.Op("CX", 2, 3), orc.OpQC("ad3", K3, 2), orc.Op3("CCZ", 6, 2, 4), orc.Op("FSIM(0.1,0.3)", 4, 1, control=6), orc.Op("MY", 3)]
    # the reference refuses a controlled NON-adjacent 2q gate other than CX/CZ (src/hilbert.jl:58-64); the oracle follows it
    v = rand_state(N, 1)
    s = bt.CuState.from_numpy(v)
    _, mids = bt.apply(od[:5] + od[6:], s, rng=bt.Draws(4), track_measurements=True)
    ref, mo = orc.apply_ops(v, oo[:5] + oo[6:], draws=orc.Draws(4), track_measurements=True)
    assert mids == mo and np.max(np.abs(s.to_numpy() - ref)) < TOL
    with pytest.raises(ValueError):
        oo[5].expand(N)
    s2 = bt.CuState.from_numpy(v)
    bt.apply(s2, od[5])                   # the device supports it (documented extension) ...
    bt._lib.check(s2.lib.bt_set_strict(1))
    try:
        with pytest.raises(bt._lib.BTError):
            bt.apply(s2, od[5])           # ... unless strict mode asks for the reference's behaviour
    finally:
        bt._lib.check(s2.lib.bt_set_strict(0))


def test_handle_pool_reuse_is_clean(bt):
    """destroyed small handles are parked and re-used (per-shot loops create a zero_state per shot, src/ops.jl:619-630)"""
    s = bt.plus_state(7)
    bt.apply(s, bt.Op("MZ", 2), rng=bt.Draws(1))
    del s
    t = bt.zero_state(7)
    want = np.zeros(128, dtype=complex); want[0] = 1
    assert np.array_equal(t.to_numpy(), want) and t.launch_count() <= 1
    u = bt.zero_state(7, 3)  # other shape: not the parked handle
    assert u.to_numpy().shape == (3, 128)
    del t, u
    bt._lib.check(bt._lib.load().bt_pool_release())
    assert np.array_equal(bt.zero_state(7).to_numpy(), want)
