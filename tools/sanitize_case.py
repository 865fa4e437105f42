"""Small fused circuit for compute-sanitizer (memcheck / racecheck): 14 qubits, structured + dense blocks, two tile shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
bt = ge.load_package()
from importlib import import_module
wl = import_module(ge.PKG_NAME + ".workloads")
N = 14
for env in ({}, {"BT_TILE_BITS": "10", "BT_TILE_LOWB": "3"}):
    os.environ.update(env)
    s = bt.zero_state(N)
    bt.apply(wl.to_ops(bt, wl.c2_qft_layered(N, 4, 3)), s)
    bt.apply(wl.to_ops(bt, wl.c5_random(N, 3, 5)), s)
    print(env, "norm", float(bt.norm2(s)))
