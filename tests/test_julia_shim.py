"""Julia is not installed in the build image, so the shim (particles.jl_b200/julia/device.jl) cannot be executed here.
What can be checked without Julia is checked: every `ccall` names a symbol the header declares and the library exports,
passes exactly as many arguments as the C prototype takes, and the two structs that cross the ABI by value have the
field count and byte size of their C definitions."""
import ctypes as C
import os
import re

from particles_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = open(os.path.join(ROOT, "particles.jl_b200", "julia", "device.jl")).read()
HEADER = open(os.path.join(ROOT, "include", "particles_b200.h")).read()

JL_SIZE = {"Int32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Ptr{Float64}": 8, "NTuple{4,Float64}": 32}


def _balanced(text, start):
    depth, i = 0, start
    while True:
        depth += {"(": 1, ")": -1}.get(text[i], 0)
        i += 1
        if depth == 0:
            return text[start + 1:i - 1]


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls():
    for m in re.finditer(r"ccall\(", SHIM):
        args = _split_top(_balanced(SHIM, m.end() - 1))
        name = re.match(r"\(:(\w+), PF_LIB\)", args[0]).group(1)
        argtypes = _split_top(args[2].strip()[1:-1])
        yield name, argtypes, args[3:]


def test_every_ccall_matches_the_header_and_the_library():
    L = _lib.load()
    protos = {name: argtypes for name, _, argtypes in _lib.SYMBOLS}
    seen = set()
    for name, argtypes, values in _ccalls():
        assert re.search(r"\b%s\(" % name, HEADER), f"{name} is not declared in include/particles_b200.h"
        assert hasattr(L, name), f"{name} is not exported by the library"
        assert len(argtypes) == len(protos[name]), f"{name}: {len(argtypes)} argument types in the shim, {len(protos[name])} in the C prototype"
        assert len(values) == len(argtypes), f"{name}: argument types and values differ in number"
        seen.add(name)
    # the path the reference's generic functions need is bound
    for need in ("pf_create", "pf_destroy", "pf_step", "pf_filter", "pf_filter_labeled", "pf_get_logweights", "pf_get_ncomponents",
                 "pf_get_particle", "pf_get_assignments", "pf_get_last_step", "pf_ncomponents_dist", "pf_assignment_similarity",
                 "pf_randindex", "pf_posterior_predictive", "pf_weighted_mean", "pf_save", "pf_load", "pf_mg_create_local",
                 "pf_mg_filter", "pf_mg_get_assignments"):
        assert need in seen, f"the shim does not bind {need}"


def _jl_struct(name):
    body = re.search(r"struct %s\n(.*?)\nend" % name, SHIM, re.S).group(1)
    fields = []
    for line in body.splitlines():
        for part in line.split(";"):
            part = part.split("#")[0].strip()
            if part:
                fname, ftype = part.split("::")
                fields.append((fname.strip(), ftype.strip()))
    return fields


def test_struct_layouts_match_the_c_definitions():
    for jl, ct in (("PfConfig", _lib.pf_config), ("PfStepStats", _lib.pf_step_stats)):
        fields = _jl_struct(jl)
        assert [f for f, _ in fields] == [f for f, _ in ct._fields_], jl
        # natural alignment of isbits fields (what Julia and C both do): running offset with padding
        off = 0
        for (fname, ftype), (_, ctype) in zip(fields, ct._fields_):
            size = JL_SIZE[ftype]
            align = 8 if size >= 8 else 4
            off = (off + align - 1) // align * align
            assert off == getattr(ct, fname).offset, f"{jl}.{fname}: offset {off} in Julia, {getattr(ct, fname).offset} in C"
            assert size == C.sizeof(ctype), f"{jl}.{fname}"
            off += size
        assert (off + 7) // 8 * 8 == C.sizeof(ct), jl


def test_exception_mapping_and_defaults():
    assert "rc == 1 && throw(ArgumentError(msg))" in SHIM                  # PF_ERR_ARG -> ArgumentError
    assert "DimensionMismatch(\"Ground truth and particles have different number of obs!\")" in SHIM
    # the same default traversal order as the Python mirror (the fast, sort-free one)
    assert "order::Int32=PF_ORDER_INDEX" in SHIM
    assert not re.search(r"^\s*module\s", SHIM, re.M)                       # an include, not a second module
