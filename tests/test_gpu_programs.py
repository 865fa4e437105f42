"""GPU parity of the register programs of the fused tile kernel (bt_tile.cu: structured micro-ops kept from the named
gates -- real / RX-like / diagonal / phase 1-qubit ops, CX, CPHASE -- plus the conditional forms for control and phase
bits outside a program) against the kron-chain oracle, through bt_sv_apply_circuit.

Every micro-op kind, every program position and the three homes of a control bit (program position, other tile bit,
outside the tile) are reached by shrinking the tile (BT_TILE_BITS) and the always-in-tile low bits (BT_TILE_LOWB), which
the library reads at every call.  Tolerance 1e-10 absolute on amplitudes (north star)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rand_state(N, seed):
    g = np.random.default_rng(seed)
    v = g.normal(size=1 << N) + 1j * g.normal(size=1 << N)
    return v / np.linalg.norm(v)


ONE = ["H", "X", "Y", "Z", "S", "SD", "T", "SX", "RX", "RY", "RZ", "P", "U3"]
TWO = ["CX", "CNOT", "CZ", "CP", "RZZ"]


def structured_circuit(bt, orc, N, n_ops, seed, far=True):
    """named gates only (so that blocks keep a structured form), on random -- also distant -- qubits; controlled 1-qubit
    gates with the control anywhere"""
    g = np.random.default_rng(seed)
    od, oo = [], []

    def both(name, *a, **k):
        od.append(bt.Op(name, *a, **k)); oo.append(orc.Op(name, *a, **k))

    for _ in range(n_ops):
        r = g.random()
        th = round(float(g.uniform(-3, 3)), 3)
        if r < 0.5 or N == 1:
            n = ONE[int(g.integers(len(ONE)))]
            q = int(g.integers(1, N + 1))
            if n in ("RX", "RY", "RZ", "P"):
                n = f"{n}({th})"
            elif n == "U3":
                n = f"U3({th},{round(th * 0.7, 3)},{round(-th * 1.3, 3)})"
            both(n, q)
        elif r < 0.85:
            n = TWO[int(g.integers(len(TWO)))]
            if far:
                q, t = (int(x) + 1 for x in g.choice(N, 2, replace=False))
            else:
                q = int(g.integers(1, N)); t = q + 1
                if g.random() < 0.5:
                    q, t = t, q
            if n in ("CP", "RZZ"):
                n = f"{n}({th})"
            both(n, q, t)
        else:
            q, c = (int(x) + 1 for x in g.choice(N, 2, replace=False))
            n = ["X", "Z", "T", "S", f"RZ({th})", f"P({th})"][int(g.integers(6))]
            both(n, q, control=c)
    return od, oo


class tile_env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def fused_vs_oracle(bt, orc, N, od, oo, seed):
    v = rand_state(N, seed)
    s = bt.CuState.from_numpy(v)
    bt.apply(od, s)
    ref = orc.apply_ops(v, oo)
    return float(np.max(np.abs(s.to_numpy() - ref)))


@pytest.mark.parametrize("N", [4, 5, 6, 8, 10, 12, 13, 14])
def test_structured_circuits_default_tile(bt, orc, N):
    od, oo = structured_circuit(bt, orc, N, 120, 1000 + N)
    assert fused_vs_oracle(bt, orc, N, od, oo, N) < TOL


@pytest.mark.parametrize("T,lowb", [(8, 3), (8, 5), (9, 4), (10, 3), (11, 5), (12, 3)])
def test_structured_circuits_small_tiles(bt, orc, T, lowb):
    """N = 14 with tiles of 8..12 bits: controls / phase bits fall outside the tile (tile-base predicate), inside the tile
    but outside the program (thread predicate) and on program positions; tile bit sets with gaps exercise the tensor map
    with a gap dimension and the multi-box copies."""
    N = 14
    with tile_env(BT_TILE_BITS=T, BT_TILE_LOWB=lowb):
        for seed in range(3):
            od, oo = structured_circuit(bt, orc, N, 150, 7 * T + lowb + 100 * seed)
            assert fused_vs_oracle(bt, orc, N, od, oo, seed) < TOL


def test_structured_brickwork_and_qft(bt, orc):
    """the two shapes of the headline workload at a size the oracle finishes: QFT (chains of controlled phases around one
    H per qubit) and H/RX/RY/RZ/T + CNOT/CZ/CP brickwork"""
    N = 13
    from importlib import import_module
    wl = import_module(bt.__name__ + ".workloads")
    for specs in (wl.qft(N), wl.layered(N, 12, 5), wl.c2_qft_layered(N, 6, 9)):
        od = wl.to_ops(bt, specs)
        oo = [orc.Op(n, q, t, control=c) for (n, q, t, c) in specs]
        for env in ({}, {"BT_TILE_BITS": 9, "BT_TILE_LOWB": 3}, {"BT_FUSE_MAX_GATES": 3}):
            with tile_env(**env):
                assert fused_vs_oracle(bt, orc, N, od, oo, 1) < TOL


def test_programs_equal_dense_path(bt):
    """same circuit with the structured forms switched off (BT_TILE_PROGS=0: dense 4x4 blocks) and through the single-gate
    kernels: three independent implementations, one answer"""
    N = 16
    from importlib import import_module
    wl = import_module(bt.__name__ + ".workloads")
    ops = wl.to_ops(bt, wl.c2_qft_layered(N, 10, 3))
    arr = bt.pack_gates(ops)
    L = bt._lib
    outs = []
    for env, fuse in (({}, 1), ({"BT_TILE_PROGS": 0}, 1), ({}, 0)):
        with tile_env(**env):
            s = bt.zero_state(N)
            L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), fuse))
            outs.append(s.to_numpy())
    assert np.max(np.abs(outs[0] - outs[2])) < TOL
    assert np.max(np.abs(outs[1] - outs[2])) < TOL


def test_batched_states_through_programs(bt, orc):
    """batch index bits are ordinary high index bits for the tile kernel: programs on a batch of 4 trajectories"""
    N, B = 9, 4
    od, oo = structured_circuit(bt, orc, N, 80, 77)
    g = np.random.default_rng(5)
    v = g.normal(size=(B, 1 << N)) + 1j * g.normal(size=(B, 1 << N))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    s = bt.CuState.from_numpy(v)
    bt.apply(od, s)
    got = s.to_numpy()
    for t in range(B):
        assert np.max(np.abs(got[t] - orc.apply_ops(v[t], oo))) < TOL


def test_specialised_passes_equal_the_interpreter(bt, orc):
    """csrc/bt_jit.cu: the straight-line kernel compiled for a pass applies the same ops in the same order as the interpreter
    (with the 2x2 blocks rescaled by a pivot and the pass scalar applied once), so the amplitudes agree to rounding (1e-13)
    and with the oracle within 1e-10.  BT_TILE_JIT=2 compiles every pass at first sight; the statistics must show
    specialised launches, and a second run of the same structure must reuse the modules."""
    import ctypes as C
    L = bt._lib
    lib = L.load()
    N = 13
    od, oo = structured_circuit(bt, orc, N, 60, 4242)
    v = rand_state(N, 5)
    outs = {}
    c0, l0, f0 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.bt_jit_stats(C.byref(c0), C.byref(l0), C.byref(f0), None)

    def disk_hits():
        h = C.c_uint64()
        lib.bt_jit_cache_info(C.byref(h), None, 0)
        return h.value

    d0 = disk_hits()
    for mode in (0, 2):
        with tile_env(BT_TILE_JIT=mode, BT_TILE_BITS=10, BT_TILE_LOWB=3):
            s = bt.CuState.from_numpy(v)
            bt.apply(od, s)
            outs[mode] = s.to_numpy()
    c1, l1, f1 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.bt_jit_stats(C.byref(c1), C.byref(l1), C.byref(f1), None)
    if c1.value == c0.value and f1.value > f0.value:
        pytest.skip("NVRTC not available on this box: passes stay on the interpreter")
    # specialised launches must have happened; their modules are new (NVRTC or the on-disk cubin cache) unless an earlier test of
    # this process already loaded the same pass structures
    assert l1.value > l0.value and c1.value + disk_hits() >= c0.value + d0
    assert np.max(np.abs(outs[0] - outs[2])) < 1e-13
    assert np.max(np.abs(outs[2] - orc.apply_ops(v, oo))) < TOL
    # second run of the same structure with other angles: cached modules, new coefficients
    with tile_env(BT_TILE_JIT=2, BT_TILE_BITS=10, BT_TILE_LOWB=3):
        s = bt.CuState.from_numpy(v)
        bt.apply(od, s)
        c2 = C.c_uint64()
        lib.bt_jit_stats(C.byref(c2), None, None, None)
        assert c2.value == c1.value          # nothing recompiled
        assert np.array_equal(s.to_numpy(), outs[2])


@pytest.mark.parametrize("free_bits", [(6, 9, 11, 13, 16, 17, 19), (7, 14, 16, 18, 21, 22, 23), (5, 8, 10, 12, 14, 16, 18)])
def test_tiles_with_many_runs_move_as_up_to_16_tensor_boxes(bt, orc, free_bits):
    """bt_tile.cu build_tensor_map: a tile whose bits form more than four runs enumerates the tile bits of the runs above
    the fourth -- up to 4 bits = 16 boxes of 4 KB (the shapes the scheduler produces on the 28-qubit headline circuit, e.g.
    bits {0-4, 6, 9, 13, 20, 23, 24, 26}).  One pass acting on 7 scattered qubits at 24 qubits, interpreter and specialised
    kernel, against the strided CPU oracle; the pass must be a tensor-copy pass (it is specialised when NVRTC is present)."""
    import ctypes as C
    from oracle import strided as S

    N = 24
    lib = bt._lib.load()
    names = ["H", "RY(0.3)", "RX(1.1)", "H", "RY(2.2)", "RX(0.7)", "H"]
    qubits = [N - b for b in free_bits]
    specs = [(n, q, -1, -2) for n, q in zip(names, qubits)]
    specs += [("CX", qubits[0], qubits[3], -2), ("CZ", qubits[1], qubits[5], -2), ("RZ(0.4)", qubits[6], -1, -2), ("CX", qubits[2], qubits[4], -2),
              ("CP(0.9)", qubits[6], N - 20, -2), ("RY(0.5)", qubits[3], -1, -2)]
    od = [bt.Op(n, q, t, control=c) for n, q, t, c in specs]
    oo = [orc.Op(n, q, t, control=c) for n, q, t, c in specs]
    v = rand_state(N, 77)
    ref = S.SV(N, v)
    ref.apply_ops(oo)
    for mode in (0, 2):
        c0, l0 = C.c_uint64(), C.c_uint64()
        lib.bt_jit_stats(C.byref(c0), C.byref(l0), None, None)
        with tile_env(BT_TILE_JIT=mode):
            s = bt.CuState.from_numpy(v)
            n0 = s.launch_count()
            bt.apply(od, s)
            assert s.launch_count() - n0 <= 2          # one fused pass (two if the scheduler splits)
            assert np.max(np.abs(s.to_numpy() - ref.v)) < TOL
        c1, l1, f1 = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.bt_jit_stats(C.byref(c1), C.byref(l1), C.byref(f1), None)
        if mode == 2 and c1.value > c0.value:
            assert l1.value >= l0.value + 1            # the specialiser only takes tensor-copy passes


def test_specialised_qft_passes_are_cross_checked_on_the_device(bt, orc):
    """QFT passes carry dozens of conditional phases per register program -- the shape for which NVRTC 12.9 miscompiled the branchy
    form of the generated code (profiles/r2_jit_nvrtc129.txt).  With the default code shape for the run-time compiler found on this
    box: (a) BT_JIT_VERIFY cross-checks every specialised launch against the interpreter kernel on the device: no disagreement;
    (b) the final amplitudes equal the interpreter's to rounding and the strided oracle's within 1e-10."""
    import ctypes as C
    from importlib import import_module

    import __graft_entry__ as ge
    from oracle import strided as S

    wl = import_module(ge.PKG_NAME + ".workloads")
    L = bt._lib
    lib = L.load()
    N = 20
    specs = wl.qft(N)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    chk0, bad0 = C.c_uint64(), C.c_uint64()
    lib.bt_jit_verify_stats(C.byref(chk0), C.byref(bad0))
    outs = {}
    for mode, verify in ((0, 0), (2, 1)):
        with tile_env(BT_TILE_JIT=mode, BT_JIT_VERIFY=verify):
            s = bt.zero_state(N)
            L.check(lib.bt_sv_set_basis(s.h, 5))
            L.check(lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1))
            outs[mode] = s.to_numpy()
    chk1, bad1 = C.c_uint64(), C.c_uint64()
    lib.bt_jit_verify_stats(C.byref(chk1), C.byref(bad1))
    c, l, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.bt_jit_stats(C.byref(c), C.byref(l), C.byref(f), None)
    if l.value == 0 and f.value > 0:
        pytest.skip("NVRTC not available on this box: passes stay on the interpreter")
    assert chk1.value > chk0.value and bad1.value == bad0.value
    assert np.max(np.abs(outs[0] - outs[2])) < 1e-13
    ref0 = np.zeros(1 << N, dtype=np.complex128)
    ref0[5] = 1
    sv = S.SV(N, ref0)
    sv.apply_ops(wl.to_ops(orc, specs))
    assert np.max(np.abs(outs[2] - sv.v)) < TOL
