"""N > 1 path on CPU: world_size-2 (and 4) gloo processes execute the library's segment plan (bt_plan_circuit_host,
pure host logic) on numpy shards -- gates through the strided oracle, remaps through a gloo all-gather -- and must
reproduce the unsharded oracle state.  Checks the planner (dependency order, locality of non-diagonal targets),
the logical->physical bit bookkeeping and the rank-conditional treatment of controls / diagonal factors."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _block_diagonal_in(m, nbits, t):
    D = 1 << nbits
    for r in range(D):
        for c in range(D):
            if ((r >> t) & 1) != ((c >> t) & 1) and m[r, c] != 0:
                return False
    return True


def _worker(rank, world, port, N, specs, ret):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from importlib import import_module

    bt = ge.load_package()
    D = import_module(ge.PKG_NAME + ".dist")
    wl = import_module(ge.PKG_NAME + ".workloads")
    from oracle import bt_oracle as O
    from oracle import strided as S

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import torch

    g = world.bit_length() - 1
    nl = N - g
    ops = wl.to_ops(O, specs)
    arr = bt.pack_gates(wl.to_ops(bt, specs))
    plan = D.plan_circuit(N, world, arr)
    assert sorted(i for seg in plan for i in seg["gates"]) == list(range(len(ops)))  # every gate exactly once
    layout = list(range(N))
    local = np.zeros(1 << nl, dtype=np.complex128)
    if rank == 0:
        local[0] = 1.0

    def gather_phys():
        t = torch.from_numpy(np.stack([local.real, local.imag]))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return np.concatenate([(o[0] + 1j * o[1]).numpy() for o in outs])

    def to_logical(phys, lay):
        idx = np.arange(1 << N, dtype=np.uint64)
        pidx = np.zeros_like(idx)
        for lb in range(N):
            pidx |= ((idx >> np.uint64(lb)) & np.uint64(1)) << np.uint64(lay[lb])
        return phys[pidx]

    n_remaps = 0
    for seg in plan:
        if seg["remap"]:
            logical = to_logical(gather_phys(), layout)
            layout = seg["layout"]
            # physical vector under the new layout, then this rank's slice
            idx = np.arange(1 << N, dtype=np.uint64)
            lidx = np.zeros_like(idx)
            for lb in range(N):
                lidx |= ((idx >> np.uint64(layout[lb])) & np.uint64(1)) << np.uint64(lb)
            phys_new = logical[lidx]
            local = np.ascontiguousarray(phys_new[rank << nl:(rank + 1) << nl])
            n_remaps += 1
        else:
            assert seg["layout"] == layout
        for gi in seg["gates"]:
            op = ops[gi]
            if op.target_qubit == -1:
                qubits = [op.qubit]            # matrix bit 0
            else:
                qubits = [op.target_qubit, op.qubit]  # matrix bit 0 <-> target, bit 1 <-> qubit
            m = np.array(op.mat)
            if op.control != -2:               # P1 (x) m + P0 (x) I, control as the top matrix bit
                D0 = m.shape[0]
                full = np.eye(2 * D0, dtype=complex)
                full[D0:, D0:] = m
                m = full
                qubits = qubits + [op.control]
            nb = len(qubits)
            # resolve bits that live on the rank index
            keep_bits, tb = [], []
            for t, q in enumerate(qubits):
                pb = layout[N - q]
                if pb >= nl:
                    assert _block_diagonal_in(m, nb, t), f"gate {gi} has a non-diagonal target on a global qubit in its segment"
                else:
                    keep_bits.append(t)
                    tb.append(pb)
            sel = [j for j in range(1 << nb) if all((((j >> t) & 1) == ((rank >> (layout[N - qubits[t]] - nl)) & 1)) for t in range(nb) if t not in keep_bits)]
            sub = m[np.ix_(sel, sel)]
            if len(tb) == 0:
                local *= sub[0, 0]
            else:
                S.apply_bits(local, nl, tb, sub)
    final = to_logical(gather_phys(), layout)
    if rank == 0:
        ref = O.apply_ops(O.zero_state(N), ops) if N <= 10 else S.SV(N).apply_ops(ops).v
        ret["err"] = float(np.max(np.abs(final - ref)))
        ret["remaps"] = n_remaps
        ret["segments"] = len(plan)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, N, specs, port):
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, N, specs, ret), nprocs=world, join=True)
    return dict(ret)


@pytest.mark.parametrize("world,port", [(2, 29611), (4, 29612)])
def test_c5_plan_executes_correctly_on_gloo_shards(world, port):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from importlib import import_module

    ge.load_package()
    wl = import_module(ge.PKG_NAME + ".workloads")
    N = 12
    r = _run(world, N, wl.c5_random(N, 8, 31), port)
    assert r["err"] < 1e-12, r
    assert r["remaps"] >= 1


def test_qft_and_controlled_gates_plan_on_gloo_shards():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from importlib import import_module

    ge.load_package()
    wl = import_module(ge.PKG_NAME + ".workloads")
    N = 11
    specs = wl.qft(N) + [("X", 11, -1, 1), ("RY(0.3)", 1, -1, 9), ("CX", 2, 8, -2), ("FSIM(0.2,0.1)", 1, 7, -2), ("CZ", 1, 2, -2), ("H", 1, -1, -2), ("H", 2, -1, -2)]
    r = _run(2, N, specs, 29613)
    assert r["err"] < 1e-12, r
