"""Print the fused passes (items / register programs per pass) of a short layered circuit at full size (BT_TILE_DEBUG=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BT_TILE_DEBUG"] = "1"
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module("bluetangle_jl_b200.workloads")
N = int(os.environ.get("DBG_N", "28")); depth = int(os.environ.get("DBG_DEPTH", "6"))
specs = wl.qft(N) if os.environ.get("DBG_CIRC", "layered") == "qft" else wl.layered(N, depth, 28)
arr = bt.pack_gates(wl.to_ops(bt, specs))
s = bt.zero_state(N)
L.check(s.lib.bt_sv_apply_circuit(s.h, L.ptr(arr), len(arr), 1)); s.sync()
print("gates", len(arr), "norm", bt.norm2(s))
