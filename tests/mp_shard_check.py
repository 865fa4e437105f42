"""Multi-process shard check (run under torchrun on >= 2 GPUs; tests/test_gpu_multi.py launches it):
CUDA-IPC peer mapping, remap over NVLink, planner, reductions through the all-reduce callback."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

bt = ge.load_package()
L = bt._lib
from importlib import import_module  # noqa: E402

wl = import_module(ge.PKG_NAME + ".workloads")
D = import_module(ge.PKG_NAME + ".dist")


def main():
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    ok = True
    for name, specs, basis in (("c5", wl.c5_random(N, 8, 31), 0), ("qft", wl.qft(N), 12345 % (1 << N))):
        arr = bt.pack_gates(wl.to_ops(bt, specs))
        st = D.ShardedState(N)
        L.check(st.lib.bt_sv_set_basis(st.h, basis))
        L.check(st.lib.bt_sv_apply_circuit(st.h, L.ptr(arr), len(arr), 1))
        full = st.gather_logical()
        ez = np.empty(N)
        L.check(st.lib.bt_sv_expect_1q_all(st.h, L.ptr(L.cmat(bt.gate["Z"], 2)), L.pdouble(ez)))
        nrm = np.empty(1)
        L.check(st.lib.bt_sv_norm2(st.h, L.pdouble(nrm)))
        ex = np.empty(1)
        pauli = ("XY" + "Z" * (N - 4) + "YX").encode()
        L.check(st.lib.bt_sv_expect_pauli(st.h, pauli, L.pdouble(ex)))
        us = np.random.default_rng(5).random(257)
        smp = np.empty(257, dtype=np.int64)
        L.check(st.lib.bt_sv_sample(st.h, L.pdouble(us), 257, smp.ctypes.data_as(C.POINTER(C.c_int64))))
        u = np.array([0.37])
        out = np.zeros(1, dtype=np.int32)
        p0 = np.zeros(1)
        L.check(st.lib.bt_sv_measure_z(st.h, 2, L.pdouble(u), out.ctypes.data_as(C.POINTER(C.c_int32)), L.pdouble(p0), 0))
        after = st.gather_logical()
        # three more measurements in ONE call (joint distribution reduced over the shards through the all-reduce callback), the first
        # of them on qubit 1 = the top index bit, which lives on the rank id until the library remaps it
        qs3 = (C.c_int * 3)(1, N, 5)
        u3 = np.array([[0.81, 0.12, 0.55]])
        out3 = np.zeros((1, 3), dtype=np.int32)
        L.check(st.lib.bt_sv_measure_z_multi(st.h, 3, qs3, L.pdouble(u3), out3.ctypes.data_as(C.POINTER(C.c_int32)), None))
        after3 = st.gather_logical()
        nrem = st.remap_stats()
        if rank == 0:
            ref = bt.basis_state(N, basis)
            L.check(ref.lib.bt_sv_apply_circuit(ref.h, L.ptr(arr), len(arr), 0))
            r = ref.to_numpy()
            e1 = float(np.max(np.abs(full - r)))
            e2 = float(np.max(np.abs(ez - bt.expect(ref, "Z"))))
            e3 = abs(nrm[0] - bt.norm2(ref))
            exr = np.empty(1)
            L.check(ref.lib.bt_sv_expect_pauli(ref.h, pauli, L.pdouble(exr)))
            e4 = abs(ex[0] - exr[0])
            same_samples = bool(np.array_equal(smp, bt.sample(ref, 257, uniforms=us)))
            o2 = np.zeros(1, dtype=np.int32)
            p2 = np.zeros(1)
            L.check(ref.lib.bt_sv_measure_z(ref.h, 2, L.pdouble(u), o2.ctypes.data_as(C.POINTER(C.c_int32)), L.pdouble(p2), 0))
            e5 = float(np.max(np.abs(after - ref.to_numpy())))
            o3 = np.zeros(3, dtype=np.int32)
            for j in range(3):  # the reference: three sequential single-GPU measurements with the same draws
                oj = np.zeros(1, dtype=np.int32)
                L.check(ref.lib.bt_sv_measure_z(ref.h, int(qs3[j]), L.pdouble(np.array([u3[0, j]])), oj.ctypes.data_as(C.POINTER(C.c_int32)), None, 0))
                o3[j] = oj[0]
            e6 = float(np.max(np.abs(after3 - ref.to_numpy())))
            good = e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12 and e4 < 1e-12 and same_samples and out[0] == o2[0] and e5 < 1e-12 and np.array_equal(out3[0], o3) and e6 < 1e-12
            ok = ok and good
            print(f"[mp_shard_check] {name} N={N} world={world}: amp={e1:.1e} expZ={e2:.1e} norm={e3:.1e} pauli={e4:.1e} samples={same_samples} "
                  f"outcome={out[0]}=={o2[0]} p0={p0[0]:.6f} post={e5:.1e} multi={[int(v) for v in out3[0]]}=={[int(v) for v in o3]} post3={e6:.1e} remaps={nrem[0]} {'OK' if good else 'FAIL'}", flush=True)
        del st
        dist.barrier()
    # back-to-back steps without any host synchronisation in between (what bench.py's timed loop does): the remaps of
    # consecutive steps are ordered by the device-side flags alone.  Every step starts from a different basis state, so a
    # stale or torn pull cannot cancel out; amplitudes after the last step are compared with the single-GPU state.
    Nb = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    for order in ("1", "0"):
        os.environ["BT_REMAP_ORDER"] = order
        arr = bt.pack_gates(wl.to_ops(bt, wl.c5_random(Nb, 6, 31)))
        st = D.ShardedState(Nb)
        basis = 0
        for k in range(steps):
            basis = (k * 977 + 5) % (1 << Nb)
            L.check(st.lib.bt_sv_set_basis(st.h, basis))
            L.check(st.lib.bt_sv_apply_circuit(st.h, L.ptr(arr), len(arr), 1))
        full = st.gather_logical()
        nrem = st.remap_stats()
        if rank == 0:
            ref = bt.basis_state(Nb, basis)
            L.check(ref.lib.bt_sv_apply_circuit(ref.h, L.ptr(arr), len(arr), 0))
            e1 = float(np.max(np.abs(full - ref.to_numpy())))
            good = e1 < 1e-12 and nrem[0] >= steps
            ok = ok and good
            print(f"[mp_shard_check] back-to-back N={Nb} world={world} steps={steps} order={order}: amp={e1:.1e} remaps={nrem[0]} {'OK' if good else 'FAIL'}", flush=True)
            del ref
        del st
        dist.barrier()
    os.environ.pop("BT_REMAP_ORDER", None)
    if rank == 0:
        print("MP_SHARD_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
