"""Sharded state vectors on ONE GPU (all shards in this process): layout bookkeeping, rank-conditional controls and
diagonals, the peer-memory remap kernel and the circuit scheduler, against the unsharded device state and the oracle."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def dist_mod():
    import __graft_entry__ as ge
    from importlib import import_module

    return import_module(ge.PKG_NAME + ".dist"), import_module(ge.PKG_NAME + ".workloads")


def rand_state(N, seed):
    g = np.random.default_rng(seed)
    v = g.normal(size=1 << N) + 1j * g.normal(size=1 << N)
    return v / np.linalg.norm(v)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_upload_gather_roundtrip_and_explicit_remap(bt, world):
    D, _ = dist_mod()
    N = 13
    v = rand_state(N, world)
    sh = D.LocalShards(N, world)
    sh.upload_logical(v)
    assert np.array_equal(sh.gather_logical(), v)
    # arbitrary permutation that keeps the 5 lowest physical bits in place
    g = np.random.default_rng(world)
    perm = list(range(5)) + list(5 + g.permutation(N - 5))
    import ctypes as C
    arr = (C.c_int * N)(*perm)
    sh.each(lambda h: sh.lib.bt_sv_remap(h, arr))
    assert sh.layout() == perm
    assert np.array_equal(sh.gather_logical(), v)  # pure data movement: bit-exact


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gates_on_global_qubits(bt, orc, world):
    D, _ = dist_mod()
    N = 12
    g = world.bit_length() - 1
    v = rand_state(N, 7)
    ops = []
    for q in range(1, g + 2):  # the lowest-numbered qubits are the global ones
        ops += [bt.Op("RZ(0.3)", q), bt.Op("T", q), bt.Op("H", q), bt.Op("CZ", q, N), bt.Op("CX", q, N - 1), bt.Op("CX", N - 2, q), bt.Op("CP(0.4)", q, q + 1),
                bt.Op("FSIM(0.2,0.1)", q, N - 3), bt.Op("X", N, control=q), bt.Op("RY(0.2)", q, control=N)]
    sh = D.LocalShards(N, world)
    sh.upload_logical(v)
    ref = bt.CuState.from_numpy(v)
    vo = v.copy()
    for o in ops:
        sh.apply(o)
        bt.apply(ref, o)
        vo = orc.Op(o.name, o.qubit, o.target_qubit, control=o.control).expand(N) @ vo
    got = sh.gather_logical()
    assert np.max(np.abs(got - ref.to_numpy())) < 1e-13
    assert np.max(np.abs(got - vo)) < TOL
    assert sh.remap_stats()[0] >= 1


@pytest.mark.parametrize("world,fuse", [(2, 0), (4, 1), (8, 1), (2, 1)])
def test_c5_circuit_sharded_equals_single(bt, world, fuse):
    D, wl = dist_mod()
    N = 14
    specs = wl.c5_random(N, 6, 31)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    sh = D.LocalShards(N, world)
    sh.apply_circuit(arr, fuse)
    ref = bt.zero_state(N)
    bt._lib.check(ref.lib.bt_sv_apply_circuit(ref.h, bt._lib.ptr(arr), len(arr), 0))
    assert np.max(np.abs(sh.gather_logical() - ref.to_numpy())) < 1e-12
    # cross-P checksum of SURVEY C5: norm and <Z_q> from per-shard partial sums
    import ctypes as C
    nrm = sh.partial("bt_sv_norm2", n_out=1)
    assert abs(nrm[0] - 1) < 1e-9


def test_qft_sharded(bt):
    D, wl = dist_mod()
    N = 13
    arr = bt.pack_gates(wl.to_ops(bt, wl.qft(N)))
    for world in (2, 8):
        sh = D.LocalShards(N, world)
        sh.set_basis(1234)
        sh.apply_circuit(arr, 1)
        ref = bt.basis_state(N, 1234)
        bt._lib.check(ref.lib.bt_sv_apply_circuit(ref.h, bt._lib.ptr(arr), len(arr), 0))
        assert np.max(np.abs(sh.gather_logical() - ref.to_numpy())) < 1e-12
