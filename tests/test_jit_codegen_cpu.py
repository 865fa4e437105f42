"""The pass specialiser's generated CUDA text, executed on the HOST: the planner's dry run (bt_fusion_plan_host with BT_JIT_DUMP=1)
prints the straight-line kernel of every fused pass; the program bodies are wrapped into plain C++ (threads become a loop, the tensor
copy a gather with the 128-byte swizzle), compiled with g++ and applied tile by tile to a state vector, which must then equal the
oracle's result of the same circuit.  This pins the code generator (op order, pivots, pass scalar, condition masks, CX renaming,
shared-memory slots) without a GPU and independently of the run-time compiler -- a mismatch on the device that this test does not
see is the compiler's (NVRTC 12.9 and BT_JIT_VARIANT)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DUMP = '''
import sys, os, ctypes as C
sys.path.insert(0, {root!r})
os.environ["BT_JIT_DUMP"] = "1"
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module(ge.PKG_NAME + ".workloads")
specs = {specs}
arr = bt.pack_gates(wl.to_ops(bt, specs))
npass, nblk, lc = C.c_int(), C.c_int(), (C.c_int * 4)()
L.check(L.load().bt_fusion_plan_host({N}, L.ptr(arr), len(arr), C.byref(npass), C.byref(nblk), None, None, None, None, 0, lc))
print(npass.value, lc[0], lc[1], lc[2])
'''


def emulate(N, specs_expr, variant, tmp_path, opt=None, shapes=None):
    env = dict(os.environ, BT_JIT_VARIANT=str(variant))
    if opt is not None:
        env["BT_JIT_OPT"] = str(opt)
    r = subprocess.run([sys.executable, "-c", DUMP.format(root=ROOT, specs=specs_expr, N=N)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    npass, launches, progs, other = (int(x) for x in r.stdout.split())
    if shapes is not None:
        shapes.update({k: r.stderr.count(k) for k in shapes})
    passes = re.split(r"// ===== pass \d+ =====\n", r.stderr)[1:]
    assert other == 0 and len(passes) == launches, "the circuit must plan into program-only passes for this test"
    cpp = ["#include <stdint.h>\n#include <math.h>\nstruct double2 { double x, y; };\nstatic inline double2 make_double2(double a, double b) { double2 r = {a, b}; return r; }\n#define PROG_AMPS 16\n"
           "#include <string.h>\nstatic inline double bt_xsign(double x, uint32_t m) { uint64_t u; memcpy(&u, &x, 8); u ^= (uint64_t)m << 32; memcpy(&x, &u, 8); return x; }\n"
           "static inline double bt_neg(double x) { return bt_xsign(x, 0x80000000u); }\n"]
    meta = []
    for pi, src in enumerate(passes):
        ncoef = int(re.search(r"struct BtCoefs \{ double c\[(\d+)\]; \};", src).group(1))
        coefs = [float(x) for x in re.findall(r"\[\d+\]=(\S+)", src.split("// coefficients:")[1])]
        assert len(coefs) == ncoef
        lowb = int(re.search(r"uint64_t base = \(uint64_t\)blockIdx.x << (\d+);", src).group(1))
        tbits = list(range(lowb)) + [int(x) for x in re.findall(r"base = \(\(base >> (\d+)\) <<", src)]
        body = src[src.index("  if (tid <"):src.index('  asm volatile("fence.proxy.async.shared::cta;"')]
        body = body.replace("__syncthreads();", "").replace("#pragma unroll 1\n", "").replace("#pragma unroll\n", "")
        body = re.sub(r'asm volatile\("mov\.b64 %0, %0;" : "\+l"\(gl\)\);', "", body)
        body = body.replace('uint32_t zoff; asm volatile("mov.u32 %0, 0;" : "=r"(zoff));', "uint32_t zoff = 0;")
        nthreads = int(re.search(r"__launch_bounds__\((\d+),", src).group(1))
        body = re.sub(r"  if \(tid < (\d+)u\) \{", rf"  for (uint32_t tid = 0; tid < {nthreads}u; ++tid) if (tid < \1u) {{", body)
        cpp.append(f'struct Coefs{pi} {{ double c[{ncoef}]; }};\nextern "C" void pass{pi}(double2* sm, uint64_t base, const Coefs{pi}* Cp) {{ const Coefs{pi}& C = *Cp;\n{body}\n}}\n')
        meta.append((tbits, coefs))
    src_path, so_path = tmp_path / "emul.cpp", tmp_path / "emul.so"
    src_path.write_text("".join(cpp))
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", str(so_path), str(src_path)])
    lib = C.CDLL(str(so_path))
    state = np.zeros(1 << N, dtype=np.complex128)
    state[5] = 1
    for pi, (tbits, coefs) in enumerate(meta):
        T = len(tbits)
        loc = np.arange(1 << T, dtype=np.int64)
        off = np.zeros(1 << T, dtype=np.int64)
        for j, b in enumerate(tbits):
            off |= ((loc >> j) & 1) << b
        slot = loc ^ ((loc >> 3) & 7)  # CU_TENSOR_MAP_SWIZZLE_128B: 16-byte chunk index XOR (128-byte row index mod 8)
        other_bits = [b for b in range(N) if b not in tbits]
        cf = (C.c_double * len(coefs))(*coefs)
        f = getattr(lib, f"pass{pi}")
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        for t in range(1 << len(other_bits)):
            base = 0
            for j, b in enumerate(other_bits):
                base |= ((t >> j) & 1) << b
            tile = np.empty(1 << T, dtype=np.complex128)
            tile[slot] = state[base + off]
            f(tile.ctypes.data, base, C.addressof(cf))
            state[base + off] = tile[slot]
    return state, launches, progs


@pytest.mark.parametrize("variant", [0, 2, 34, 128])
@pytest.mark.parametrize("N,specs_expr", [(13, "wl.qft(13)"), (13, "wl.layered(13, 5, 28)"), (12, "wl.qft(12) + wl.layered(12, 4, 3)")])
def test_generated_pass_source_equals_the_oracle_on_the_host(bt, orc, tmp_path, N, specs_expr, variant):
    from importlib import import_module

    import __graft_entry__ as ge
    from oracle import strided as S

    got, launches, progs = emulate(N, specs_expr, variant, tmp_path)
    wl = import_module(ge.PKG_NAME + ".workloads")
    ops = wl.to_ops(orc, eval(specs_expr, {"wl": wl}))
    ref0 = np.zeros(1 << N, dtype=np.complex128)
    ref0[5] = 1
    sv = S.SV(N, ref0)
    sv.apply_ops(ops)
    assert launches >= 1 and progs >= launches
    assert np.max(np.abs(got - sv.v)) < 1e-13


def test_fp64_saving_code_shapes_are_generated_and_exact(bt, orc, tmp_path):
    """BT_JIT_OPT (default 7): deferred per-thread scalars carrying the pass scalar, sign-bit flips for -1 factors / negations, the
    pi-reduced shear under a selected sign mask.  The circuit is chosen so that every one of these shapes occurs; the same circuit
    with BT_JIT_OPT=0 (the earlier generator text) must give the same state."""
    from importlib import import_module

    import __graft_entry__ as ge
    from oracle import strided as S

    N, expr = 13, "wl.layered(13, 5, 28)"
    shapes = {"bt_xsign(fma(st": 0, "const uint32_t sm_ = ": 0, "double fr = on0": 0, "double fr = C.c": 0, "bt_neg(": 0, "i1n": 0}
    (tmp_path / "new").mkdir()
    (tmp_path / "old").mkdir()  # separate directories: dlopen caches a library by its path
    got, _, _ = emulate(N, expr, 2, tmp_path / "new", opt=7, shapes=shapes)
    assert all(v > 0 for v in shapes.values()), shapes
    old_shapes = dict.fromkeys(shapes, 0)
    old, _, _ = emulate(N, expr, 2, tmp_path / "old", opt=0, shapes=old_shapes)
    assert not any(old_shapes.values()), old_shapes
    wl = import_module(ge.PKG_NAME + ".workloads")
    ref0 = np.zeros(1 << N, dtype=np.complex128)
    ref0[5] = 1
    sv = S.SV(N, ref0)
    sv.apply_ops(wl.to_ops(orc, eval(expr, {"wl": wl})))
    assert np.max(np.abs(got - sv.v)) < 1e-13 and np.max(np.abs(old - sv.v)) < 1e-13


def test_fp64_instruction_estimate_of_the_generated_text():
    """tools/jit_fp64_count.source_estimate (what bench.py's roofline.fp64.issued block calls): FP64 instructions per amplitude from
    the generated text of a workload's passes.  The reshaped generator (BT_JIT_OPT=7) must issue fewer than the earlier text in both
    code shapes, the executed count can only be below the static one, and an RX on every qubit costs exactly 2 FMAs per amplitude."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import jit_fp64_count as J

    work = "13, wl.layered(13, 6, 28)"
    for variant in ("0", "2"):
        old = J.source_estimate(work, {"BT_JIT_VARIANT": variant, "BT_JIT_OPT": "0"})
        new = J.source_estimate(work, {"BT_JIT_VARIANT": variant, "BT_JIT_OPT": "7"})
        assert old["specialised_launches"] == new["specialised_launches"] >= 1 and new["other_items"] == 0
        assert 0 < new["executed_per_amplitude"] <= new["static_per_amplitude"] < old["static_per_amplitude"]
        assert new["executed_per_amplitude"] < old["executed_per_amplitude"]
    rx = J.source_estimate("12, [('RX(0.3)', q, -1, -2) for q in range(1, 13)]", {"BT_JIT_VARIANT": "2"})
    # one pass of three programs; a rotation [[1, -i t], [-i t, 1]] (after the pivot) = 2 FMAs per amplitude, + the real pass scalar (2 multiplications)
    assert rx["specialised_launches"] == 1 and rx["programs"] == 3 and rx["static_per_amplitude"] == 12 * 2 + 2, rx


RANDOM_PAIRS = """(lambda g: [s for l in range({layers}) for s in (
    [(["H", "RX(%r)" % g.uniform(0, 6.28), "RY(%r)" % g.uniform(0, 6.28), "RZ(%r)" % g.uniform(0, 6.28), "T", "S", "Z", "X"][int(g.integers(8))], q, -1, -2) for q in range(1, {N} + 1)]
    + [((["CNOT", "CZ", "CP(%r)" % g.uniform(0, 6.28)][int(g.integers(3))],) + tuple(int(x) for x in g.choice({N}, 2, replace=False) + 1) + (-2,)) for _ in range({N} // 2)])])(
    __import__("numpy").random.Generator(__import__("numpy").random.PCG64({seed})))"""


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("N,layers,seed", [(16, 6, 1), (15, 8, 2)])
def test_generated_source_with_random_qubit_pairs_and_bits_outside_the_tile(bt, orc, tmp_path, N, layers, seed, variant):
    """Two-qubit gates on random (non-adjacent) pairs of a register wider than the tile: controls and phase bits fall on tile bits
    that are not program positions (thread-dependent conditions) and on bits outside the tile (conditions on the tile's base
    index), in both code shapes; S / Z / X join the one-qubit mix.  Host execution of the generated text against the oracle."""
    from importlib import import_module

    import __graft_entry__ as ge
    from oracle import strided as S

    expr = RANDOM_PAIRS.format(N=N, layers=layers, seed=seed).replace("\n", " ")
    shapes = {"(base & 0x0ull)": 0, "if ((base &": 0, "double fr = ": 0}
    got, launches, progs = emulate(N, expr, variant, tmp_path, shapes=shapes)
    wl = import_module(ge.PKG_NAME + ".workloads")
    ops = wl.to_ops(orc, eval(expr, {"wl": wl}))
    ref0 = np.zeros(1 << N, dtype=np.complex128)
    ref0[5] = 1
    sv = S.SV(N, ref0)
    sv.apply_ops(ops)
    assert shapes["double fr = "] > 0 and shapes["if ((base &"] > 0
    assert np.max(np.abs(got - sv.v)) < 1e-13
