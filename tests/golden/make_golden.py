"""Generates tests/golden/golden_r1.npz with the oracle (oracle/bt_oracle.py).  The reference holds no numeric
vectors and Julia is not available here, so these fixtures are produced by the oracle that tests/
test_oracle_reference_kats.py pins against the reference's own known answers.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from oracle import bt_oracle as O

bt = ge.load_package()
from importlib import import_module

wl = import_module(ge.PKG_NAME + ".workloads")
from test_oracle_reference_kats import random_ops  # noqa: E402


def main():
    out = {}
    # C1 (BASELINE.json configs[0]): 12-qubit brickwork H/CNOT/RZ depth 20 + 4096 shots
    specs = wl.c1_brickwork(12, 20, 12)
    st = O.apply_ops(O.zero_state(12), wl.to_ops(O, specs))
    us = np.random.Generator(np.random.PCG64(12)).random(4096)
    out["c1_state"] = st
    out["c1_uniforms"] = us
    out["c1_samples"] = O.sample(st, us)
    out["c1_expect_z"] = np.array(O.expect(st, "Z"))
    out["c1_zz_1_12"] = np.array(O.correlation(st, "Z,Z", [1, 12]))
    out["c1_xy_3_7"] = np.array(O.correlation(st, "X,Y", [3, 7]))
    # noisy monitored circuits, N = 6: outcomes + states for 8 seeds
    for seed in range(8):
        ops = random_ops(6, 5, 100 + seed, measure_prob=0.2)
        nm = O.NoiseModel.model(["depolarizing", "amplitude_damping"][seed % 2], 0.05)
        d = O.Draws(seed)
        s, mids = O.apply_ops(O.zero_state(6), ops, noise=nm, draws=d, track_measurements=True)
        out[f"mon{seed}_state"] = s
        out[f"mon{seed}_mids"] = np.array(mids, dtype=np.int64)
        out[f"mon{seed}_draws"] = np.array(d.log)
    # noisy density matrix, N = 5 (C3 in miniature: depolarizing + amplitude damping after every gate)
    rho = O.rho_from_state(O.zero_state(5))
    for e in wl.c3_noisy_dm(5, 4, 14):
        if e[0] == "gate":
            name, q, t, c = e[1]
            rho = O.apply(rho, O.Op(name, q, t, control=c))
        else:
            _, model, p, q, t = e
            rho = O.apply(rho, O.OpQC.model(model, p, q, t))
    out["c3_rho"] = rho
    out["c3_expect_z"] = np.array(O.expect(rho, "Z"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_r1.npz"), **out)
    print("written", sum(v.nbytes for v in out.values()), "bytes uncompressed")


if __name__ == "__main__":
    main()
