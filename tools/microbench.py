"""Per-kernel HBM roofline microbenchmark (28q SV / 14q DM by default): one launch per timed iteration, inputs
larger than L2.  Prints achieved GB/s = algorithmic bytes / CUDA-event time."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

bt = ge.load_package()
L = bt._lib


def timed(s, fn, reps=5, warm=2):
    lib = s.lib
    for _ in range(warm):
        fn()
    ms = C.c_float()
    L.check(lib.bt_sv_timer_start(s.h))
    for _ in range(reps):
        fn()
    L.check(lib.bt_sv_timer_stop(s.h, C.byref(ms)))
    return ms.value / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    peak = 6544.7
    try:
        peak = json.load(open(os.path.join(ge.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    s = bt.plus_state(N)
    lib = s.lib
    full = 32.0 * (1 << N)
    rows = []
    H = L.cmat(bt.gate["H"], 2)
    for q in [N, N - 1, N - 2, N - 3, N - 4, N - 5, N - 6, N - 8, N - 12, N // 2, 2, 1]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_1q(s.h, q, L.ptr(H), -2)))
        rows.append((f"1q dense H bit {N-q}", full, t))
    U4 = L.cmat(np.linalg.qr(np.random.default_rng(0).normal(size=(4, 4)) + 1j * np.random.default_rng(1).normal(size=(4, 4)))[0], 4)
    for (q, t_) in [(N, N - 1), (N - 1, N), (N - 2, N - 6), (N, 1), (3, 9), (1, 2), (N // 2, N // 2 + 1)]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_2q(s.h, q, t_, L.ptr(U4), -2)))
        rows.append((f"2q dense bits ({N-q},{N-t_})", full, t))
    RZ = L.cmat(bt.gates("RZ(0.3)"), 2)
    for q in [N, N - 3, 1]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_1q(s.h, q, L.ptr(RZ), -2)))
        rows.append((f"diag RZ bit {N-q}", full, t))
    Tg = L.cmat(bt.gate["T"], 2)
    for q in [N, N - 3, 1]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_1q(s.h, q, L.ptr(Tg), -2)))
        rows.append((f"diag T (half) bit {N-q}", full / 2, t))
    CX = L.cmat(bt.gate["CX"], 4)
    for (q, t_) in [(N - 1, N), (N, N - 1), (5, 9), (1, N)]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_2q(s.h, q, t_, L.ptr(CX), -2)))
        rows.append((f"CX ctrl bit {N-q} tgt bit {N-t_} (half)", full / 2, t))
    CP = L.cmat(bt.gates("CP(0.3)"), 4)
    for (q, t_) in [(N - 1, N), (N - 4, N - 2), (5, 9), (1, N)]:
        t = timed(s, lambda: L.check(lib.bt_sv_apply_2q(s.h, q, t_, L.ptr(CP), -2)))
        rows.append((f"CP bits ({N-q},{N-t_}) (quarter)", full / 4, t))
    out = np.empty(4, dtype=np.complex128)
    for q in [N, N - 6, 1]:
        t = timed(s, lambda: L.check(lib.bt_sv_rdm1(s.h, q, L.ptr(out))))
        rows.append((f"rdm1 bit {N-q} (read only)", full / 2, t))
    ez = np.empty(N)
    Z = L.cmat(bt.gate["Z"], 2)
    t = timed(s, lambda: L.check(lib.bt_sv_expect_1q_all(s.h, L.ptr(Z), L.pdouble(ez))))
    rows.append(("expect Z all qubits (read only)", full / 2, t))
    u = np.array([0.3])
    o = np.zeros(1, dtype=np.int32)
    p0 = np.zeros(1)
    t = timed(s, lambda: (L.check(lib.bt_sv_set_plus(s.h)), L.check(lib.bt_sv_measure_z(s.h, 3, L.pdouble(u), o.ctypes.data_as(C.POINTER(C.c_int32)), L.pdouble(p0), 0))))
    rows.append(("set_plus + measure_z (fill 16 + rdm 16 + collapse 24 B/amp)", (16 + 16 + 24) * (1 << N), t))
    L.check(lib.bt_sv_set_plus(s.h))
    us = np.random.default_rng(0).random(4096)
    so = np.empty(4096, dtype=np.int64)
    t = timed(s, lambda: L.check(lib.bt_sv_sample(s.h, L.pdouble(us), 4096, so.ctypes.data_as(C.POINTER(C.c_int64)))))
    rows.append(("sample 4096 shots (read only)", full / 2, t))
    # state-vector Kraus trajectory step: RDM read (16 B/amp) + decision + scaled-Kraus apply (32 B/amp)
    K1s = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01)])
    K2s = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01, True)])
    ua = np.array([0.5])
    for q in [N, N // 2, 1]:
        t = timed(s, lambda: L.check(lib.bt_sv_kraus(s.h, 1, q, -1, L.ptr(K1s), 4, L.pdouble(ua), None)))
        rows.append((f"SV 1q depolarizing Kraus step qubit {q} (rdm 16 + apply 32 B/amp)", 48.0 * (1 << N), t))
    for (q, t_) in [(N - 1, N), (3, 17)]:
        t = timed(s, lambda: L.check(lib.bt_sv_kraus(s.h, 2, q, t_, L.ptr(K2s), 16, L.pdouble(ua), None)))
        rows.append((f"SV 2q depolarizing Kraus step ({q},{t_}) (rdm 16 + apply 32 B/amp)", 48.0 * (1 << N), t))
    del s
    # density matrix
    n = N // 2
    d = bt.CuRho(n)
    dl = d.lib

    def timed_dm(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ms = C.c_float()
        L.check(dl.bt_dm_timer_start(d.h))
        for _ in range(reps):
            fn()
        L.check(dl.bt_dm_timer_stop(d.h, C.byref(ms)))
        return ms.value / reps

    fulld = 32.0 * (1 << (2 * n))
    for q in [n, n // 2, 1]:
        t = timed_dm(lambda: L.check(dl.bt_dm_apply_1q(d.h, q, L.ptr(H), -2)))
        rows.append((f"DM 1q unitary qubit {q}", fulld, t))
    for (q, t_) in [(n - 1, n), (2, 9), (1, 2)]:
        t = timed_dm(lambda: L.check(dl.bt_dm_apply_2q(d.h, q, t_, L.ptr(U4), -2)))
        rows.append((f"DM 2q unitary ({q},{t_}) (one 16x16 superoperator pass)", fulld, t))
    K1 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01)])
    K2 = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in bt.noise_model("depolarizing", 0.01, True)])
    Kc = np.stack([np.asfortranarray(k).reshape(-1, order="F") for k in [np.sqrt(0.7) * np.eye(4), np.sqrt(0.2) * np.kron(bt.gate["X"], bt.gate["X"]), np.sqrt(0.1) * np.kron(bt.gate["Z"], bt.gate["Z"])]])
    for q in [n, 1]:
        t = timed_dm(lambda: L.check(dl.bt_dm_kraus(d.h, 1, q, -1, L.ptr(K1), 4)))
        rows.append((f"DM 1q depolarizing Kraus qubit {q}", fulld, t))
    for (q, t_) in [(n - 1, n), (2, 9)]:
        t = timed_dm(lambda: L.check(dl.bt_dm_kraus(d.h, 2, q, t_, L.ptr(K2), 16)))
        rows.append((f"DM 2q depolarizing (16 Kraus, product) ({q},{t_})", fulld, t))
        t = timed_dm(lambda: L.check(dl.bt_dm_kraus(d.h, 2, q, t_, L.ptr(Kc), 3)))
        rows.append((f"DM 2q correlated channel (16x16) ({q},{t_})", fulld, t))
    print(f"{'kernel':72s} {'ms':>9s} {'GB/s':>9s} {'of measured':>11s} {'of 8TB/s':>9s}")
    for name, byts, ms in rows:
        gbs = byts / ms / 1e6
        print(f"{name:72s} {ms:9.4f} {gbs:9.1f} {gbs/peak:11.3f} {gbs/8000:9.3f}")


if __name__ == "__main__":
    main()
