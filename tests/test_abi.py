"""The C-ABI library loads and exports every symbol include/bluetangle_cuda.h declares; no compute calls (CPU box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bluetangle_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", txt))
    names -= {"bt_barrier_fn", "bt_allreduce_fn"}
    return sorted(names)


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 60
    for must in ["bt_sv_apply_1q", "bt_sv_apply_2q", "bt_sv_apply_circuit", "bt_sv_measure_z", "bt_sv_kraus", "bt_sv_sample", "bt_sv_expect_pauli",
                 "bt_dm_apply_1q", "bt_dm_kraus", "bt_sv_create_shard", "bt_sv_remap", "bt_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol(bt):
    lib = bt._lib.load()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_prototypes_cover_the_header(bt):
    protos = set(bt._lib.PROTOTYPES) | {"bt_last_error"}
    assert set(declared_symbols()) <= protos, sorted(set(declared_symbols()) - protos)


def test_no_device_is_a_loud_error_not_a_fallback(bt):
    """On a box without a GPU the product path must fail loudly (no CPU fallback)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bt._lib.BTError):
        bt.zero_state(4)
    assert b"no CPU fallback" in bt._lib.load().bt_last_error() or b"CUDA" in bt._lib.load().bt_last_error()


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline / reference legs may touch oracle/ (checker, never product)."""
    pkg = os.path.join(ROOT, "bluetangle.jl_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libbt_oracle|oracle/_build|oracle/_ref|bt_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), f


def test_julia_shim_binds_only_declared_entry_points():
    """julia/BlueTangleCUDA.jl cannot be executed here (no Julia); at least every symbol it ccalls must be one the header
    declares and the library exports."""
    src = open(os.path.join(ROOT, "julia", "BlueTangleCUDA.jl")).read()
    called = set(re.findall(r"ccall\(\(:(bt_[a-z0-9_]+)", src))
    assert len(called) >= 25
    assert called <= set(declared_symbols()), sorted(called - set(declared_symbols()))


def _abi_gen():
    import importlib.util

    spec = importlib.util.spec_from_file_location("abi_gen", os.path.join(ROOT, "tools", "abi_gen.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_python_prototypes_match_the_header_signatures(bt):
    """Argument by argument: the ctypes table of _lib.py against the declarations parsed out of the header (tools/abi_gen.py);
    a changed C signature that is not followed in the binding fails here, not at run time with a corrupted stack."""
    G = _abi_gen()
    sig = {n: c for n, _, c in G.parse_header()}
    assert set(sig) - {"bt_last_error"} == set(bt._lib.PROTOTYPES), set(sig) ^ set(bt._lib.PROTOTYPES)
    for name, args in bt._lib.PROTOTYPES.items():
        assert [G.ctypes_category(a) for a in args] == sig[name], name


def test_julia_ccall_argument_tuples_match_the_header_signatures():
    G = _abi_gen()
    sig = {n: (r, c) for n, r, c in G.parse_header()}
    calls = G.julia_ccalls(os.path.join(ROOT, "julia", "BlueTangleCUDA.jl"))
    assert len(calls) >= 25
    for name, ret, cats in calls:
        assert cats == sig[name][1], (name, cats, sig[name][1])
        assert ret == ("Cstring" if sig[name][0] == "cstr" else "Cint"), name


def _build_abi_smoke(tmp_path):
    import subprocess

    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "bluetangle.jl_b200", "lib")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe,
           "-L" + libdir, "-l:libbluetangle_cuda.so", "-lm", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_caller_compiles_and_links_against_the_header(tmp_path):
    """tests/abi_smoke.c with -Wall -Wextra -Werror: a plain C caller sees exactly the header's prototypes; without a device it
    exits 2 after bt_device_count (no compute on the CPU box)."""
    import subprocess
    import torch

    exe = _build_abi_smoke(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_caller_runs_the_minimal_path_on_the_device(tmp_path):
    """create -> apply_1q -> apply_2q -> rdm1 -> measure_z -> sample -> expect -> destroy from C, numbers checked inside."""
    import subprocess

    r = subprocess.run([_build_abi_smoke(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_julia_shim_blocks_and_brackets_balance():
    """No Julia in this image: the least a text check can do for julia/BlueTangleCUDA.jl is to balance its block keywords against
    `end` (keywords inside brackets are comprehensions / filters and open nothing) and its brackets."""
    src = open(os.path.join(ROOT, "julia", "BlueTangleCUDA.jl")).read()
    openers = {"function", "if", "for", "while", "let", "begin", "struct", "module", "try", "do", "quote", "macro"}
    stack, bdepth = [], 0
    for n, line in enumerate(src.split("\n"), 1):
        i, tok, in_str, line = 0, "", False, line + " "
        while i < len(line):
            c = line[i]
            if in_str:
                i += 2 if c == "\\" else 1
                in_str = c != '"'
                continue
            if c == '"':
                in_str = True
            elif c == "#":
                break
            elif c.isalnum() or c in "_!":
                tok += c
                i += 1
                continue
            if tok:
                prev = line[:i - len(tok)].rstrip()
                if tok == "end" and bdepth == 0:
                    assert stack, f"line {n}: `end` without an open block"
                    stack.pop()
                elif tok in openers and bdepth == 0 and not prev.endswith((":", ".")):
                    stack.append((tok, n))
                tok = ""
            bdepth += (c in "([{") - (c in ")]}")
            assert bdepth >= 0, f"line {n}: closing bracket without an opening one"
            i += 1
        assert not in_str, f"line {n}: unterminated string"
    assert not stack and bdepth == 0, (stack[-3:], bdepth)
