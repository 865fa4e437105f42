import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
N = 28
s = bt.plus_state(N)
out = np.empty(4, dtype=np.complex128)
for q in (N, 5):
    L.check(s.lib.bt_sv_rdm1(s.h, q, L.ptr(out)))
ez = np.empty(N)
L.check(s.lib.bt_sv_expect_1q_all(s.h, L.ptr(L.cmat(bt.gate["Z"], 2)), L.pdouble(ez)))
us = np.random.default_rng(0).random(4096); so = np.empty(4096, dtype=np.int64)
L.check(s.lib.bt_sv_sample(s.h, L.pdouble(us), 4096, so.ctypes.data_as(C.POINTER(C.c_int64))))
n = np.empty(1); L.check(s.lib.bt_sv_norm2(s.h, L.pdouble(n)))
pz = np.empty(1); L.check(s.lib.bt_sv_expect_pauli(s.h, ("XZ" * 14).encode(), L.pdouble(pz)))
print(out, ez[:3], so[:4], n, pz)
