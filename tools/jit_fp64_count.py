#!/usr/bin/env python
"""FP64 instructions of the specialised passes, counted WITHOUT a GPU: the planner's dry run (bt_fusion_plan_host, BT_JIT_DUMP=1)
prints the CUDA text of every fused pass of a workload, nvcc compiles each for sm_100a, cuobjdump -sass gives DFMA + DADD + DMUL per
kernel.  Every FP64 instruction sits inside the two-iteration group loop of a program (128 threads x 2 iterations x 16 amplitudes =
one 2^12 tile), so static count / 16 = FP64 instructions per amplitude per pass.  The passes are bound by the FP64 pipe (DESIGN.md
section 4), which makes this count the quantity to minimise.  A second, source-level estimate weights the blocks behind a branch
(BT_JIT_VARIANT=0) by how often a warp executes them (a condition on lane bits diverges: always; every warp-uniform bit: 1/2).

Usage: python tools/jit_fp64_count.py [c2|c5|layers] [ENV=VALUE ...]   e.g.  BT_JIT_OPT=0, BT_JIT_VARIANT=0"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DUMP = '''
import sys, os, ctypes as C
sys.path.insert(0, {root!r})
os.environ["BT_JIT_DUMP"] = "1"
import __graft_entry__ as ge
bt = ge.load_package(); L = bt._lib
from importlib import import_module
wl = import_module(ge.PKG_NAME + ".workloads")
N, specs = {work}
arr = bt.pack_gates(wl.to_ops(bt, specs))
npass, nblk, lc = C.c_int(), C.c_int(), (C.c_int * 4)()
L.check(L.load().bt_fusion_plan_host(N, L.ptr(arr), len(arr), C.byref(npass), C.byref(nblk), None, None, None, None, 0, lc))
print(len(arr), npass.value, lc[0], lc[1], lc[2])
'''
WORK = {"c2": "28, wl.c2_qft_layered()", "c5": "31, wl.c5_random(31)", "layers": "28, wl.layered(28, 100, 28)"}


def expr_cost(e):
    e, cost = e.strip(), 0
    while "fma(" in e:
        i = e.rindex("fma(")
        d = 0
        for j in range(i + 3, len(e)):
            d += (e[j] == "(") - (e[j] == ")")
            if d == 0:
                break
        cost += 1 + e[i + 4:j].count(" * ")
        e = e[:i] + "T" + e[j + 1:]
    if "?" in e:
        return cost
    k = e.count(" * ")
    terms = len(re.split(r" [+-] ", e))
    if k == 0 and terms <= 1:
        return cost
    return cost + (k + max(0, terms - k - 1) if k else terms - 1)


def line_cost(s):
    if "sm[" in s or s.startswith("uint") or s.startswith("if (it ==") or "asm" in s:
        return 0
    c = 0
    for m in re.finditer(r"(?:const double |double |, |; |\{ |^)([A-Za-z_][A-Za-z_0-9\[\]]*) (\*?=) ([^;,{}]*(?:\([^;{}]*\))?[^;,{}]*)", s):
        name, op, e = m.group(1), m.group(2), m.group(3)
        if name in ("on", "on0", "sm_", "b", "gl", "s0", "g0", "zoff"):
            continue
        if op == "*=":
            c += 1
        elif "bt_xsign" in e or "bt_neg" in e:
            c += expr_cost(re.sub(r", sm_\)", ")", re.sub(r"bt_(xsign|neg)\(", "(", e))) if "fma(" in e else 0
        elif e.strip().startswith("-") and " " not in e.strip():
            c += 1  # a plain negation is a DADD
        else:
            c += expr_cost(e)
    return c


def dynamic_estimate(src, branchy):
    total = 0.0
    for p in re.split(r"  if \(tid < \d+u\) \{  // program", src)[1:]:
        lane = 0
        for m in re.finditer(r"if \(tid & (\d+)u\) g0 \^= (\d+)u;", p):
            if int(m.group(1)) < 32:
                lane |= int(m.group(2))
        depth, stack = 0, []
        for ln in p.split("\n"):
            s = ln.strip()
            w = 1.0
            for _, ww in stack:
                w *= ww
            m = re.match(r"if \(\(base & (0x[0-9a-f]+)ull\) == \S+ && \(gl & (0x[0-9a-f]+)ull\) == \S+\) \{", s)
            if m:
                em, lm = int(m.group(1), 16), int(m.group(2), 16)
                stack.append((depth, 0.5 ** (bin(em).count("1") + bin(lm & ~lane).count("1")) if branchy else 1.0))
            else:
                total += w * line_cost(s)
            depth += s.count("{") - s.count("}")
            while stack and depth <= stack[-1][0]:
                stack.pop()
    return total


def sass_count(path):
    cubin = path[:-3] + ".cubin"
    subprocess.check_call(["nvcc", "-cubin", "-arch=sm_100a", "-std=c++17", "-o", cubin, path], stderr=subprocess.DEVNULL)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
    reg = re.search(r"REG:(\d+)", res)
    return {k: len(re.findall(r"\b" + k + r"\b", sass)) for k in ("DFMA", "DADD", "DMUL")}, int(reg.group(1)) if reg else -1, len(re.findall(r"\b(LDL|STL)\b", sass))


def dump_passes(work_expr, env=None):
    """CUDA text of every specialised pass of a workload (planner dry run in a subprocess: the dump goes to the C stderr).
    work_expr: Python expression giving (N, specs) with `wl` = the workloads module.  Returns (ngates, launches, programs, other, [text])."""
    r = subprocess.run([sys.executable, "-c", DUMP.format(root=ROOT, work=work_expr)], capture_output=True, text=True, env=dict(os.environ, **(env or {})), timeout=600)
    if r.returncode:
        raise RuntimeError(r.stderr[-2000:])
    ngates, npass, launches, progs, other = (int(x) for x in r.stdout.split())
    return ngates, launches, progs, other, re.split(r"// ===== pass \d+ =====\n", r.stderr)[1:]


def source_estimate(work_expr, env=None):
    """FP64 instructions per amplitude per step from the generated text alone (no compiler): `static` counts every instruction
    once (agrees with the SASS count to 0.1 - 3 %, profiles/r2_jit_fp64_counts.txt), `executed` weights blocks behind a branch."""
    ngates, launches, progs, other, passes = dump_passes(work_expr, env)
    branchy = any("if ((base &" in p for p in passes)
    return {"gates": ngates, "specialised_launches": launches, "programs": progs, "other_items": other,
            "static_per_amplitude": sum(dynamic_estimate(p, False) for p in passes) / 16.0,
            "executed_per_amplitude": sum(dynamic_estimate(p, branchy) for p in passes) / 16.0, "branches_weighted": branchy}


def main():
    args = [a for a in sys.argv[1:] if "=" not in a and not a.startswith("-")]
    env = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    work = args[0] if args else "c2"
    ngates, launches, progs, other, passes = dump_passes(WORK[work], env)
    branchy = any("if ((base &" in p for p in passes)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, src in enumerate(passes):
            paths.append(os.path.join(d, f"p{i:03d}.cu"))
            open(paths[-1], "w").write(src)
        with ThreadPoolExecutor(max(1, (os.cpu_count() or 2) - 1)) as ex:
            counts = list(ex.map(sass_count, paths))
    tot = {k: sum(c[0][k] for c in counts) for k in ("DFMA", "DADD", "DMUL")}
    fp64 = sum(tot.values())
    dyn = sum(dynamic_estimate(s, branchy) for s in passes)
    stat = sum(dynamic_estimate(s, False) for s in passes)
    knobs = " ".join(a for a in sys.argv[1:] if "=" in a) or "defaults"
    print(f"{work} [{knobs}]: {ngates} gates, {launches} specialised launches ({progs} programs, {other} other items)")
    print(f"  SASS static  DFMA {tot['DFMA']}  DADD {tot['DADD']}  DMUL {tot['DMUL']}  FP64 {fp64}  = {fp64 / 16:.1f} per amplitude per step; "
          f"max registers {max(c[1] for c in counts)}, kernels with local memory {sum(1 for c in counts if c[2])}")
    print(f"  source-level estimate: static {stat / 16:.1f}, executed {dyn / 16:.1f} per amplitude ({'branches weighted' if branchy else 'branch-free form'})")
    if "-v" in sys.argv:
        for i, c in enumerate(counts):
            print(f"  pass {i:3d}  FP64 {sum(c[0].values()):5d}  registers {c[1]}")


if __name__ == "__main__":
    main()
