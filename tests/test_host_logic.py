"""Host-side logic of the front-end mirror that needs no device: tables, op constructors and their errors, packing,
draw bookkeeping, workload generators.  CPU only."""
import math

import numpy as np
import pytest

from oracle import bt_oracle as O


def test_gate_tables_match_the_oracle_restatement(bt):
    for name, m in O.GATE.items():
        assert np.array_equal(bt.gate[name], m), name
    for expr in ["P(0.3)", "RX(.1pi)", "RY(0.5π)", "RZ(-pi/4)", "U1(0.2)", "U2(0.1,0.2)", "U3(0.1,0.2,0.3)", "CP(0.3)", "GIVENS(0.2)", "FSIM(0.1,0.2)", "SWAPA(0.3)",
                 "RXX(0.4)", "RYY(0.5)", "RZZ(0.6)", "RXY(0.7)", "MZ", "MX", "MY", "RES"]:
        assert np.allclose(bt.gates(expr), O.gates(expr), atol=0, rtol=0), expr
    with pytest.raises(KeyError):
        bt.gates("NOPE")


def test_kraus_tables_match(bt):
    for model in ["amplitude_damping", "phase_damping", "phase_flip", "bit_flip", "bit_phase_flip", "depolarizing", "depolarizing_amp", "rot_x", "rot_y", "rot_z", "rot_p", "rot_xyz", "MZ"]:
        for two in (False, True):
            a, b = bt.noise_model(model, 0.13, two), O.noise_model(model, 0.13, two)
            assert len(a) == len(b)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (model, two)
    assert bt.is_valid_quantum_channel(bt.noise_model("depolarizing", 0.2))
    assert not bt.is_valid_quantum_channel([0.5 * bt.gate["I"]])


def test_op_constructors_mirror_reference_errors(bt):
    o = bt.Op("CX", 2, 3)
    assert (o.q, o.qubit, o.target_qubit, o.control, o.noisy, o.type) == (2, 2, 3, -2, True, "op")
    assert bt.Op("RZ(0.3)", 1).type == "phase"
    m = bt.Op("MX", 2)
    assert m.type == "🔬" and m.noisy is False
    with pytest.raises(ValueError):
        bt.Op("MZ", 1, control=2)          # src/struct.jl:400-402
    with pytest.raises(ValueError):
        bt.Op("X", np.eye(4), 1)           # src/struct.jl:393-396
    with pytest.raises(ValueError):
        bt.Op("CX", 1, 1)                  # src/struct.jl:465
    ccx = bt.Op("CCX", 1, 2, 3)            # src/struct.jl:436-449
    assert (ccx.name, ccx.qubit, ccx.target_qubit, ccx.control) == ("CX", 1, 3, 2)
    with pytest.raises(ValueError):
        bt.Op("CCH", 1, 2, 3)
    res = bt.Op("RES", 2)
    assert isinstance(res, bt.OpQC) and res.name == "res" and res.q == 1
    assert bt.Op(["RZ", 0.25], 3).name.startswith("RZ(")
    with pytest.raises(ValueError):
        bt.OpQC("bad", [0.5 * np.eye(2)], 1)
    with pytest.raises(ValueError):
        bt.OpQC("depolarizing", 0.1, 1, 2).__class__("x", bt.noise_model("depolarizing", 0.1), 1, 2)  # 1q Kraus with a target
    with pytest.raises(ValueError):
        bt.ifOp("H", 1)
    nm = bt.NoiseModel("depolarizing", 0.1)
    assert nm.q1.q == 1 and nm.q2.q == 2 and len(nm.q2.kraus) == 16


def test_pack_gates_layout(bt):
    ops = [bt.Op("H", 3), bt.Op("FSIM(0.1,0.2)", 2, 5), bt.Op("X", 4, control=1)]
    arr = bt.pack_gates(ops)
    assert arr.dtype.itemsize == 272
    assert list(arr["nq"]) == [1, 2, 1] and list(arr["qubit"]) == [3, 2, 4] and list(arr["target"]) == [-1, 5, -1] and list(arr["control"]) == [-2, -2, 1]
    assert np.allclose(arr[1]["m"].reshape(4, 4, order="F"), bt.gates("FSIM(0.1,0.2)"))
    assert np.allclose(arr[0]["m"][:4].reshape(2, 2, order="F"), bt.gate["H"])


def test_bit_helpers_and_postprocessing(bt):
    assert bt.int2bin(2, 4) == [0, 0, 1, 0] and bt.bin2int([0, 0, 1, 0]) == 2
    bitstr, probs = np.array([1, 4, 6]), np.array([0.36, 0.4096, 0.2304])
    for qs in ([1], [2], [3], [1, 3]):
        assert abs(bt._sample_to_expectation(bitstr, probs, 3, qs) - O.sample_to_expectation(bitstr, probs, 3, qs)) < 1e-15
    for k in range(1, 5):
        assert abs(bt.mag_moments(3, bitstr, probs, k) - O.mag_moments(3, bitstr, probs, k)) < 1e-12
    v, p = bt.get_probs_from_sample([3, 1, 3, 3], 2)
    assert list(v) == [1, 3] and np.allclose(p, [0.25, 0.75])


def test_batch_draws_consume_per_trajectory(bt):
    U = np.arange(12, dtype=float).reshape(3, 4)
    d = bt.BatchDraws(U)
    assert list(d.take()) == [0, 4, 8]
    assert list(d.take(np.array([True, False, True]))) == [1, 5, 9]
    assert list(d.cnt) == [2, 1, 2]
    assert list(d.take()) == [2, 5, 10]


def test_workload_generators_are_deterministic(bt):
    import __graft_entry__ as ge
    from importlib import import_module

    wl = import_module(ge.PKG_NAME + ".workloads")
    c2 = wl.c2_qft_layered(28, 100, 28)
    assert len(c2) == 4556 and c2 == wl.c2_qft_layered(28, 100, 28)
    assert len(wl.qft(28)) == 406
    c1 = wl.c1_brickwork()
    assert len(c1) == 20 * 12 + 10 * 6 + 10 * 5
    ops_d, ops_o = wl.to_ops(bt, c1), wl.to_ops(O, c1)
    for a, b in zip(ops_d, ops_o):
        assert np.array_equal(a.mat, b.mat) and (a.qubit, a.target_qubit, a.control) == (b.qubit, b.target_qubit, b.control)
    c4, nm = wl.c4_monitored(8, 6, 20)
    assert nm == sum(1 for s in c4 if s[0] == "MZ")


def test_qasm_front_door(bt):
    text = """
    OPENQASM 2.0; include "qelib1.inc";
    qreg q[4]; creg c[4];
    h q[0]; cx q[0],q[1]; // bell
    rz(0.25*pi) q[2]; u3(0.1,0.2,0.3) q[3]; sdg q[1]; tdg q[2]; sx q[0];
    cu1(pi/4) q[1],q[3]; crz(0.5) q[0],q[2]; swap q[2],q[3]; rzz(0.3) q[0],q[3]; fsim(0.1,0.2) q[1],q[2];
    ccx q[0],q[1],q[2]; cswap q[0],q[1],q[2];
    barrier q; reset q[3]; measure q[1] -> c[1];
    """
    ops = bt.from_qasm(text)
    names = [(o.name, o.qubit, getattr(o, "target_qubit", -1), getattr(o, "control", -2)) for o in ops]
    assert names[0] == ("H", 1, -1, -2)
    assert names[1] == ("X", 2, -1, 1)                     # cx a,b -> Op("X", b; control=a)  (src/qasm.jl:373)
    assert names[2][0].startswith("RZ(") and names[2][1] == 3
    assert names[4][0] == "SD" and names[5][0] == "TD" and names[6][0] == "XSQRT"
    assert names[7] == ("U1(pi/4)", 4, -1, 2)
    assert names[9] == ("SWAP", 3, 4, -2)
    assert names[12] == ("CX", 1, 3, 2)                    # ccx a,b,c -> Op("CCX", a, b, c) -> CX(a -> c) controlled by b
    assert names[13] == ("SWAP", 2, 3, 1)                  # cswap a,b,c -> Op("CSWAP", b, a, c)
    assert isinstance(ops[14], bt.OpQC) and ops[14].name == "res"
    assert ops[15].type == "🔬"
    assert np.allclose(ops[2].mat, bt.gates("RZ(0.25*pi)"))
    with pytest.raises(ValueError):
        bt.from_qasm("foo q[0];")
    # arguments with nested parentheses split at top-level commas only; arithmetic is evaluated by an AST walker, not eval()
    ops = bt.from_qasm("qreg q[2]; rz((1+2)*pi/4) q[0]; u3(pi/2, (0.1), -pi) q[1]; cu1(pi/(2)) q[0],q[1];")
    assert np.allclose(ops[0].mat, bt.gates(f"RZ({3 * np.pi / 4!r})")) and np.allclose(ops[1].mat, bt.gates(f"U3({np.pi / 2!r},0.1,{-np.pi!r})"))
    assert ops[2].control == 1 and np.allclose(ops[2].mat, bt.gates(f"U1({np.pi / 2!r})"))
    for hostile in ("rz(9**9**9**9) q[0];", "rz(__import__('os').getpid()) q[0];", "rz(1/0) q[0];", "rz((1) q[0];"):
        with pytest.raises(ValueError):
            bt.from_qasm("qreg q[1]; " + hostile)


def test_opf_dispatch_and_forms(bt, orc):
    """OpF (src/struct.jl:703-744): apply() hands the state to the op's function (src/hilbert.jl:486-487); op lists and
    functions are accepted, a full-register matrix is refused by the device mirror and applied by the oracle."""
    import numpy as np

    seen = []
    fake = object.__new__(bt.CuState)          # no device needed: the function form never touches the library here
    op = bt.OpF("probe", lambda st: seen.append(st))
    assert bt.apply(fake, op) is fake and seen == [fake]
    assert bt.apply(op, fake) is fake          # (op, state) argument order, src/hilbert.jl:555-557
    other = object.__new__(bt.CuState)
    assert bt.apply(fake, bt.OpF("swap-in", lambda st: other)) is other   # a function may return another device state
    bt.OpF("block", [bt.Op("H", 1), bt.Op("CX", 1, 2)])
    with pytest.raises(NotImplementedError):
        bt.OpF("dense", np.eye(4))
    # oracle: the three forms agree with each other
    v = np.random.default_rng(0).normal(size=8) + 1j * np.random.default_rng(1).normal(size=8)
    v /= np.linalg.norm(v)
    ops = [orc.Op("H", 1), orc.Op("CX", 1, 3), orc.Op("RZ(0.3)", 2)]
    seq = orc.apply_ops(v, ops)
    M = ops[2].expand(3) @ ops[1].expand(3) @ ops[0].expand(3)
    assert np.max(np.abs(orc.apply(v, orc.OpF("block", ops)) - seq)) < 1e-14
    assert np.max(np.abs(orc.apply(v, orc.OpF("dense", M)) - seq)) < 1e-14
    assert np.max(np.abs(orc.apply(v, orc.OpF("fn", lambda st: M @ st)) - seq)) < 1e-14


def test_remap_walk_is_a_bijection_and_reads_sigma_of_dest(bt, monkeypatch):
    """bt_remap_walk_host = the index arithmetic of k_remap_pull on the host: for every iteration order the walk must
    write every local destination index exactly once and read physical source index sigma(rank|dest)."""
    import ctypes as C

    L = bt._lib
    lib = L.load()
    rng = np.random.default_rng(3)
    N = 14
    for g in (1, 2, 3):
        nl = N - g
        world = 1 << g
        for trial in range(6):
            cur = list(range(N))
            if trial % 2:  # start from a permuted layout (low 5 bits fixed)
                hi = list(rng.permutation(np.arange(5, N)))
                cur = list(range(5)) + [int(x) for x in hi]
            new = cur[:]
            movable = [lb for lb in range(N) if cur[lb] >= 5]
            perm = rng.permutation(len(movable))
            for a, b in zip(movable, perm):
                new[a] = cur[movable[int(b)]]
            sigma = [0] * N  # sigma[dst phys] = src phys
            for lb in range(N):
                sigma[new[lb]] = cur[lb]
            for order, sel_lo, rot in ((0, 8, 1), (1, 5, 1), (1, 8, 1), (1, 8, 0), (1, 11, 1)):
                monkeypatch.setenv("BT_REMAP_ORDER", str(order))
                monkeypatch.setenv("BT_REMAP_SEL_LO", str(sel_lo))
                monkeypatch.setenv("BT_REMAP_ROT", str(rot))
                for rank in range(world):
                    loop = np.arange(1 << nl, dtype=np.uint64)
                    dest = np.empty_like(loop)
                    src = np.empty_like(loop)
                    ci = (C.c_int * N)(*cur)
                    ni = (C.c_int * N)(*new)
                    pu = C.POINTER(C.c_uint64)
                    L.check(lib.bt_remap_walk_host(N, nl, rank, ci, ni, loop.size, loop.ctypes.data_as(pu), dest.ctypes.data_as(pu), src.ctypes.data_as(pu)))
                    assert np.array_equal(np.sort(dest), loop)  # every destination exactly once
                    full = (np.uint64(rank) << np.uint64(nl)) | dest
                    want = np.zeros_like(full)
                    for d in range(N):
                        want |= ((full >> np.uint64(d)) & np.uint64(1)) << np.uint64(sigma[d])
                    assert np.array_equal(src, want)
                    assert np.array_equal(dest & np.uint64(31), loop & np.uint64(31))  # lanes keep their 512-byte run
                    if order == 1 and g > 1 and sel_lo == 8:
                        # consecutive 2^sel_lo chunks of the walk come from different ranks (if the remap moves rank bits at all)
                        srcrank = (src >> np.uint64(nl)).reshape(-1, 1 << sel_lo)
                        assert np.all(srcrank == srcrank[:, :1])
                        nsel = sum(1 for d in range(5, nl) if sigma[d] >= nl)
                        if nsel:
                            assert len(set(srcrank[: 1 << nsel, 0].tolist())) == 1 << nsel


def test_jacobi_tournament_schedule_covers_every_pair_once(bt):
    """bt_jacobi_pairs_host: the round-robin schedule the Schmidt-spectrum kernel (csrc/bt_linalg.cu) walks -- the pairs of a
    round are disjoint (one CTA per pair may rotate in place) and a sweep meets every unordered pair exactly once."""
    import ctypes as C

    lib = bt._lib.load()
    for n in (2, 4, 8, 32, 256):
        seen = set()
        for r in range(n - 1):
            pairs = (C.c_int * n)()
            bt._lib.check(lib.bt_jacobi_pairs_host(n, r, pairs))
            flat = list(pairs)
            assert sorted(flat) == list(range(n))
            for k in range(n // 2):
                p, q = flat[2 * k], flat[2 * k + 1]
                seen.add((min(p, q), max(p, q)))
        assert len(seen) == n * (n - 1) // 2
    assert lib.bt_jacobi_pairs_host(3, 0, (C.c_int * 4)()) != 0


def test_fusion_scheduler_look_ahead_plans_fewer_passes(bt):
    """bt_fusion_plan_host (pure host): the look-ahead scheduler (BT_FUSE_SCHED=1, default) chooses the tile bits of a pass by dry
    runs and must (a) carry every gate exactly once, (b) need no more passes than first fit on layered / QFT circuits, and
    (c) cut the passes of the 28-qubit C2 circuit by a third or more (146 -> 91 when this was written)."""
    import ctypes as C
    import os
    from importlib import import_module

    import __graft_entry__ as ge

    wl = import_module(ge.PKG_NAME + ".workloads")
    lib = bt._lib.load()

    def plan(N, specs, sched):
        os.environ["BT_FUSE_SCHED"] = str(sched)
        try:
            arr = bt.pack_gates(wl.to_ops(bt, specs))
            cap = 4096
            npass, nblk = C.c_int(), C.c_int()
            gip = (C.c_int * cap)()
            lc = (C.c_int * 4)()
            bt._lib.check(lib.bt_fusion_plan_host(N, bt._lib.ptr(arr), len(arr), C.byref(npass), C.byref(nblk), gip, None, None, None, cap, lc))
            assert sum(list(gip)[: npass.value]) == len(arr)
            assert lc[0] >= npass.value  # a pass may split into several launches, never the reverse
            return npass.value, lc[0]
        finally:
            os.environ.pop("BT_FUSE_SCHED", None)

    for N, specs in ((12, wl.layered(12, 10, 1)), (16, wl.qft(16) + wl.layered(16, 12, 2)), (20, wl.c1_brickwork(20, 10, 3) if hasattr(wl, "c1_brickwork") else wl.layered(20, 10, 3))):
        p0, l0 = plan(N, specs, 0)
        p1, l1 = plan(N, specs, 1)
        assert p1 <= p0 and l1 <= l0 + 1, (N, p0, p1, l0, l1)  # a deep pass may overflow the 8 program slots of one launch and split once
    p0, l0 = plan(28, wl.qft(28) + wl.layered(28, 100, 28), 0)
    p1, l1 = plan(28, wl.qft(28) + wl.layered(28, 100, 28), 1)
    assert p1 <= 0.67 * p0 and l1 <= 0.67 * l0, (p0, p1, l0, l1)


def test_cumulants_from_moments_recursion(bt):
    """src/func.jl:241-253 against the closed forms k1 = m1, k2 = m2 - m1^2, k3 = m3 - 3 m2 m1 + 2 m1^3, k4 = m4 - 4 m3 m1 - 3 m2^2 + 12 m2 m1^2 - 6 m1^4."""
    m = [0.3, 1.7, -0.4, 5.2, 0.9]
    k = bt.cumulants_from_moments(m)
    assert abs(k[0] - m[0]) < 1e-14 and abs(k[1] - (m[1] - m[0] ** 2)) < 1e-14
    assert abs(k[2] - (m[2] - 3 * m[1] * m[0] + 2 * m[0] ** 3)) < 1e-13
    assert abs(k[3] - (m[3] - 4 * m[2] * m[0] - 3 * m[1] ** 2 + 12 * m[1] * m[0] ** 2 - 6 * m[0] ** 4)) < 1e-13
    assert bt.cumulants_from_moments(m, 3) == k[2]


def test_adjoint_of_ops_lists_and_circuits(bt):
    """src/linalg.jl:1-51: Hermitian ops come back as they are, others get the conjugate transpose and a dagger in the name (a second
    adjoint removes it), controls / flags are kept, lists and circuits are reversed."""
    o = bt.Op("RZ(0.3)", 2)
    a = bt.adjoint(o)
    assert a.name == "RZ(0.3)†" and np.allclose(a.mat, o.mat.conj().T) and bt.adjoint(a).name == "RZ(0.3)"
    h = bt.Op("H", 1)
    assert bt.adjoint(h) is h and bt.ishermitian(h) and not bt.ishermitian(o)
    x = bt.Op("RY(0.2)", 3, control=1, noisy=False)
    ax = bt.adjoint(x)
    assert ax.control == 1 and ax.noisy is False and ax.qubit == 3
    c = bt.Op("CP(0.4)", 2, 3)
    names = [t.name for t in bt.adjoint([o, h, c])]
    assert names == ["CP(0.4)†", "H", "RZ(0.3)†"]
    circ = bt.compile([o, h, c])
    assert [t.name for t in bt.adjoint(circ).ops] == names and bt.adjoint(circ).N == circ.N
    assert bt.isunitary(o.mat) and not bt.isunitary(np.array([[1, 1], [0, 1]]))
