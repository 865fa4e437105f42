"""Generates tests/golden/golden_vqa_r1.npz with the oracle's restatement of the variational front end (src/vqa.jl):
ansatz state, Pauli-sum energies and the parameter-shift gradient at fixed parameters.  Run from the repo root:
    python tests/golden/make_golden_vqa.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bt_oracle as O

NAMES = ["RY", "RZ", "CX", "RX", "RZZ", "CZ", "RY"]
HAMS = {"tfim_open": ([-1.0, "Z,Z", -0.7, "X"], "open"), "heis_periodic": ([0.5, "X,X", 0.25, "Y,Y", -1.5, "Z,Z"], "periodic"),
        "mixed_open": ([0.3, "X,Y,Z", 1.1, "Y", -0.2, "Z,X"], "open")}


def main():
    out = {}
    N = 6
    vops, args, dim = O.variational_circuit_from_string(N, NAMES, False)
    p = np.random.Generator(np.random.PCG64(2026)).uniform(0, np.pi, dim)
    st = O.variational_apply(p, N, vops, args)
    out["n6_pars"] = p
    out["n6_state"] = st
    for key, (spec, bc) in HAMS.items():
        Hm = O.hamiltonian(N, spec, bc)
        out[f"n6_energy_{key}"] = np.array(float(np.real(np.vdot(st, Hm @ st))))
    Hm = O.hamiltonian(N, *HAMS["tfim_open"])
    loss = lambda x: float(np.real(np.vdot(x, Hm @ x)))
    l0, g = O.loss_and_grad_paramshift(p, loss, N, vops, args)
    out["n6_grad_tfim_open"] = g
    # a random 12-qubit state (>= 2^12 amplitudes: the register Walsh-Hadamard kernel on the device)
    gen = np.random.Generator(np.random.PCG64(12))
    v = gen.normal(size=1 << 12) + 1j * gen.normal(size=1 << 12)
    v /= np.linalg.norm(v)
    out["n12_state"] = v
    for key, (spec, bc) in HAMS.items():
        Hm = O.hamiltonian(12, spec, bc)
        out[f"n12_energy_{key}"] = np.array(float(np.real(np.vdot(v, Hm @ v))))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_vqa_r1.npz"), **out)
    print("written", sum(x.nbytes for x in out.values()), "bytes uncompressed")


if __name__ == "__main__":
    main()
