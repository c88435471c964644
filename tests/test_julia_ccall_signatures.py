"""Mechanical check of the Julia binding's `ccall` signatures against include/rcw_b200.h.

Julia is absent from this image, so the binding cannot run; what can be proved on the CPU is that every
`ccall((:rcw_x, LIB), Ret, (T1, T2, ...), ...)` in BatchedRayCastWorlds.jl has the arity of the C prototype and
that each argument (and the return value) has the same machine class: a pointer where C has a pointer, a 32-bit
integer where C has int32_t / uint32_t, a 64-bit one for int64_t / uint64_t, Csize_t for size_t, Cfloat / Cdouble for
float / double.  A wrong Int32 / Int64 in the binding — which the name / field-order test would wave through —
fails here (the last test proves that on a doctored copy)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcw_b200.h")
JULIA = os.path.join(ROOT, "raycastworlds.jl_b200", "julia", "BatchedRayCastWorlds.jl")

C_SCALARS = {"int32_t": "i32", "uint32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "i64",
             "size_t": "size", "float": "f32", "double": "f64", "uint8_t": "i8", "int8_t": "i8", "void": "void"}
JL_SCALARS = {"Int32": "i32", "UInt32": "i32", "Cint": "i32", "Cuint": "i32", "Int64": "i64", "UInt64": "i64",
              "Clonglong": "i64", "Csize_t": "size", "Float32": "f32", "Cfloat": "f32", "Float64": "f64",
              "Cdouble": "f64", "UInt8": "i8", "Int8": "i8", "Cvoid": "void", "Nothing": "void"}


def c_class(decl: str) -> str:
    """Machine class of a C parameter / return declaration such as `const int32_t* goal_ij` or `size_t bytes`."""
    decl = decl.strip()
    if "*" in decl:
        return "ptr"
    words = [w for w in re.findall(r"[A-Za-z_][A-Za-z_0-9]*", decl) if w not in ("const", "struct", "enum")]
    assert words, decl
    base = words[0]
    assert base in C_SCALARS, f"unknown C type in header: {decl!r}"
    return C_SCALARS[base]


def jl_class(t: str) -> str:
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr", "Ref"):
        return "ptr"
    assert t in JL_SCALARS, f"unknown Julia type in a ccall: {t!r}"
    return JL_SCALARS[t]


def header_prototypes(text: str) -> dict:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 ]*?[\s\*]+)(rcw_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("", "void") else [c_class(a) for a in args.split(",")]
        protos[name] = (c_class(ret), params)
    return protos


def split_top_level(s: str) -> list:
    """Split `A, Ref{B}, (C, D)` on the commas that are not inside braces or parentheses."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out if x.strip()]


def julia_ccalls(text: str) -> list:
    """[(name, return class, [argument classes], line)] for every ccall of an rcw_ symbol."""
    calls = []
    for m in re.finditer(r"ccall\(\(:(rcw_[a-z_0-9]+),\s*LIB\),\s*", text):
        i = m.end()
        j = text.index(",", i)                      # return type: a bare identifier or Ptr{...}
        ret = text[i:j].strip()
        k = text.index("(", j)                      # the argument-type tuple
        depth, e = 0, k
        while True:
            if text[e] in "({":
                depth += 1
            elif text[e] in ")}":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        types = split_top_level(text[k + 1:e])
        calls.append((m.group(1), jl_class(ret), [jl_class(t) for t in types], text.count("\n", 0, m.start()) + 1))
    return calls


def mismatches(header_text: str, julia_text: str) -> list:
    protos = header_prototypes(header_text)
    bad = []
    for name, ret, args, line in julia_ccalls(julia_text):
        if name not in protos:
            bad.append(f"line {line}: {name} is not declared in the header")
            continue
        c_ret, c_args = protos[name]
        if ret != c_ret:
            bad.append(f"line {line}: {name} returns {ret}, header says {c_ret}")
        if len(args) != len(c_args):
            bad.append(f"line {line}: {name} takes {len(args)} arguments, header says {len(c_args)}")
            continue
        for k, (a, c) in enumerate(zip(args, c_args)):
            if a != c:
                bad.append(f"line {line}: {name} argument {k + 1} is {a}, header says {c}")
    return bad


def test_header_prototypes_parse():
    protos = header_prototypes(open(HEADER).read())
    assert protos["rcw_version"] == ("i32", [])
    assert protos["rcw_last_error"] == ("ptr", [])
    assert protos["rcw_step_range"] == ("i32", ["ptr", "ptr", "i64", "i64"])
    assert protos["rcw_expand_columns"] == ("i32", ["ptr", "ptr", "size", "i64", "i32", "ptr"])
    assert protos["rcw_episode_stats"] == ("i32", ["ptr", "ptr", "ptr", "ptr", "i32"])
    assert len(protos) == len(set(protos)) >= 30


def test_every_ccall_matches_its_prototype():
    calls = julia_ccalls(open(JULIA).read())
    assert len(calls) >= 30
    assert mismatches(open(HEADER).read(), open(JULIA).read()) == []


@pytest.mark.parametrize("old, new, expect", [
    ("(Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64)", "(Ptr{Cvoid}, Ptr{UInt8}, Int32, Int64)", "rcw_step_range argument 3 is i32"),
    ("(Ptr{Cvoid}, Int32), env.handle, Int32(n_steps)", "(Ptr{Cvoid}, Int64), env.handle, Int32(n_steps)", "rcw_step_random argument 2 is i64"),
    ("(Ptr{Cvoid}, Ref{Int64}), env.handle, n", "(Ptr{Cvoid},), env.handle", "rcw_launch_count takes 1 arguments"),
])
def test_a_wrong_width_is_caught(old, new, expect):
    jl = open(JULIA).read()
    assert old in jl, "the doctored snippet must exist in the binding"
    bad = mismatches(open(HEADER).read(), jl.replace(old, new, 1))
    assert any(expect in b for b in bad), bad
