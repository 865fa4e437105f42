#!/usr/bin/env python
"""Signatures of include/bluetangle_cuda.h as data: the one source the Python prototype table (bluetangle.jl_b200/_lib.py)
and the Julia ccall argument tuples (julia/BlueTangleCUDA.jl) are checked against (tests/test_abi.py), so the three cannot
drift apart silently.  `python tools/abi_gen.py --python` prints a PROTOTYPES table generated from the header,
`--julia` the ccall argument tuples.

An argument is reduced to a category: i32 i64 u64 f32 f64 (scalars), p_f64 p_f32 p_i32 p_i64 p_u64 (typed pointers),
cstr, ptr (handles, bt_c64 / struct / void pointers), pp (pointer to a handle pointer), fn (callback)."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bluetangle_cuda.h")

_SCALAR = {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "float": "f32"}
_PTR = {"double": "p_f64", "float": "p_f32", "int": "p_i32", "int32_t": "p_i32", "int64_t": "p_i64", "uint64_t": "p_u64", "char": "cstr"}


def category(decl: str) -> str:
    d = decl.strip()
    d = re.sub(r"\bconst\b", "", d).strip()
    arr = "[" in d
    d = re.sub(r"\[.*?\]", "", d)
    stars = d.count("*")
    d = d.replace("*", " ")
    toks = d.split()
    base = toks[0]
    if base in ("bt_barrier_fn", "bt_allreduce_fn"):
        return "fn"
    nptr = stars + (1 if arr else 0)
    if nptr == 0:
        return _SCALAR[base]
    if nptr == 2:
        return "pp"
    return _PTR.get(base, "ptr")


def parse_header(path: str = HEADER):
    """[(name, return category, [argument categories])] in declaration order"""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    txt = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", txt, flags=re.S)
    txt = re.sub(r"typedef[^;]*;", "", txt)
    out = []
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(bt_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        cats = [] if args in ("", "void") else [category(a) for a in args.split(",")]
        out.append((name, "cstr" if "char" in ret else "i32", cats))
    return out


_PY = {"i32": "_i", "i64": "_i64", "u64": "_u64", "f64": "C.c_double", "f32": "C.c_float", "p_f64": "_pd", "p_f32": "C.POINTER(C.c_float)", "p_i32": "_pi32",
       "p_i64": "_pi64", "p_u64": "C.POINTER(_u64)", "cstr": "C.c_char_p", "ptr": "_vp", "pp": "C.POINTER(_vp)", "fn": "<callback>"}
_JL = {"i32": "Cint", "i64": "Int64", "u64": "UInt64", "f64": "Float64", "f32": "Float32", "p_f64": "Ptr{Float64}", "p_f32": "Ptr{Float32}", "p_i32": "Ptr{Int32}",
       "p_i64": "Ptr{Int64}", "p_u64": "Ptr{UInt64}", "cstr": "Cstring", "ptr": "Ptr{Cvoid}", "pp": "Ref{Ptr{Cvoid}}", "fn": "Ptr{Cvoid}"}


def ctypes_category(t) -> str:
    """category of a ctypes argument type from _lib.PROTOTYPES"""
    import ctypes as C

    table = {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64", C.c_float: "f32", C.c_void_p: "ptr", C.c_char_p: "cstr",
             C.POINTER(C.c_double): "p_f64", C.POINTER(C.c_float): "p_f32", C.POINTER(C.c_int): "p_i32", C.POINTER(C.c_int32): "p_i32",
             C.POINTER(C.c_int64): "p_i64", C.POINTER(C.c_uint64): "p_u64", C.POINTER(C.c_void_p): "pp"}
    if t in table:
        return table[t]
    if hasattr(t, "_flags_") and hasattr(t, "_argtypes_"):
        return "fn"
    raise KeyError(t)


def julia_category(t: str) -> str:
    t = t.strip()
    table = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cdouble": "f64", "Cfloat": "f32", "Float32": "f32", "Cstring": "cstr",
             "Ptr{Cvoid}": "ptr", "Ptr{ComplexF64}": "ptr", "Ptr{BtGate}": "ptr", "Ref{ComplexF64}": "ptr", "Ref{Ptr{Cvoid}}": "pp", "Ptr{Ptr{Cvoid}}": "pp",
             "Ptr{Float64}": "p_f64", "Ref{Float64}": "p_f64", "Ptr{Cfloat}": "p_f32", "Ref{Cfloat}": "p_f32", "Ptr{Float32}": "p_f32", "Ptr{Int32}": "p_i32",
             "Ptr{Cint}": "p_i32", "Ref{Cint}": "p_i32", "Ref{Int32}": "p_i32", "Ptr{Int64}": "p_i64", "Ptr{UInt64}": "p_u64", "Ref{UInt64}": "p_u64", "Ptr{UInt8}": "cstr"}
    return table[t]


def julia_ccalls(path: str):
    """[(name, return type, [argument categories])] of every ccall((:bt_..., LIB), ret, (types...), ...) in the shim"""
    src = open(path).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(bt_[a-z0-9_]+),\s*LIB\),\s*(\w+),\s*\(", src):
        i = m.end()
        depth, j = 1, i
        while depth:
            c = src[j]
            depth += c == "("
            depth -= c == ")"
            j += 1
        inner = src[i:j - 1]
        parts, cur, d = [], "", 0
        for c in inner:
            if c == "{":
                d += 1
            if c == "}":
                d -= 1
            if c == "," and d == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += c
        if cur.strip():
            parts.append(cur)
        out.append((m.group(1), m.group(2), [julia_category(p) for p in parts if p.strip()]))
    return out


if __name__ == "__main__":
    sigs = parse_header()
    if "--julia" in sys.argv:
        for name, ret, cats in sigs:
            print(f"# {name}: ccall((:{name}, LIB), {'Cstring' if ret == 'cstr' else 'Cint'}, ({', '.join(_JL[c] for c in cats)}{',' if len(cats) == 1 else ''}), ...)")
    else:
        print("PROTOTYPES = {")
        for name, ret, cats in sigs:
            if ret == "cstr":
                continue
            print(f'    "{name}": [{", ".join(_PY[c] for c in cats)}],')
        print("}")
