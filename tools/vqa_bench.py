"""Variational caller of the hot path at full size on one GPU: energy of a chain Hamiltonian through bt_sv_expect_pauli_sum
(one read of the state per basis group) against the term-by-term route (one pass per term), and one ansatz evaluation
(EfficientSU2, reps 2) -- first call (passes compiled by the specialiser) and a later call with other angles (modules reused)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
bt = ge.load_package()

N = int(sys.argv[1]) if len(sys.argv) > 1 else 28
os.environ.setdefault("BT_TILE_JIT_AFTER", "1")


def wall(fn, reps=3):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r


g = np.random.default_rng(1)
opt = bt.AnsatzOptions(N=N, ops=["RY", "RZ", "CX", "RY", "RZ", "CX", "RY", "RZ"], loss=None, deep_circuit=True, pars_initial=g.uniform(0, np.pi, 6 * N))
t0 = time.perf_counter(); st = bt.variational_apply(opt.pars_initial, opt); st.sync(); t_first = time.perf_counter() - t0
p2 = g.uniform(0, np.pi, 6 * N)
t_apply, st = wall(lambda: (lambda s: (s.sync(), s)[1])(bt.variational_apply(p2, opt)))
t_reuse, _ = wall(lambda: bt.variational_apply(p2, opt, out=st).sync())
print(f"  same evaluation into an existing state (out=..., what the gradient loop does): {t_reuse*1e3:.1f} ms")
print(f"ansatz EfficientSU2-like, {N} qubits, {len(opt.ops)} ops, {opt.dim} parameters: first evaluation {t_first*1e3:.0f} ms (passes compiled), later evaluations {t_apply*1e3:.1f} ms")
for label, spec, bc in (("TFIM open", [-1.0, "Z,Z", -0.7, "X"], "open"), ("Heisenberg periodic", [0.5, "X,X", 0.5, "Y,Y", 0.5, "Z,Z"], "periodic")):
    ps = bt.hamiltonian(N, spec, bc)
    t_sum, e1 = wall(lambda: ps.expect(st))
    t_terms, e2 = wall(lambda: bt.hamiltonian_expect(st, ps.terms), reps=1)
    print(f"{label}: {len(ps)} terms at {N} qubits: grouped Pauli-sum {t_sum*1e3:.1f} ms, term by term {t_terms*1e3:.1f} ms ({t_terms/t_sum:.1f}x), |difference| = {abs(e1-e2):.2e}, E = {e1:.9f}")
