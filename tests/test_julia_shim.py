"""The Julia shim (julia/NetworkSolversB200.jl) cannot run here (no Julia in the image); what can be checked on the CPU is
that it is written against the ABI that exists: every symbol it `ccall`s is exported by libnsb200.so with the same number of
arguments as the ctypes binding, and its POD structs list the fields of the C structs in the same order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "julia", "NetworkSolversB200.jl")).read()


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_every_ccall_matches_the_binding():
    from networksolvers_b200 import _lib
    lib = _lib.load()
    calls = re.findall(r"ccall\(\(:(nsb_\w+), lib\), (\w+), \(([^)]*(?:\([^)]*\)[^)]*)*)\)", SRC)
    direct = {c[0] for c in calls}
    assert {"nsb_extract", "nsb_update_eigsolve", "nsb_update_exp", "nsb_insert", "nsb_site_download", "nsb_multi_create",
            "nsb_multi_extract", "nsb_multi_insert", "nsb_fit_target_upload"} <= direct
    assert ":nsb_site_upload" in SRC and ":nsb_mpo_upload" in SRC      # chosen at run time (f = operator ? ... : ...)
    for name, ret, argt in calls:
        assert name in _lib.SIGNATURES, name
        assert hasattr(lib, name)
        nargs = len(_split_args(argt))
        assert nargs == len(_lib.SIGNATURES[name][1]), (name, nargs, len(_lib.SIGNATURES[name][1]))
    # symbols chosen at run time through `f = cond ? :a : :b`
    for name in re.findall(r":(nsb_\w+)", SRC):
        assert name in _lib.SIGNATURES, name


def test_struct_layouts_match():
    from networksolvers_b200 import _lib
    pairs = {"NsbTrunc": _lib.Trunc, "NsbExpand": _lib.Expand, "NsbKrylov": _lib.Krylov, "NsbExtractInfo": _lib.ExtractInfo,
             "NsbSolveInfo": _lib.SolveInfo, "NsbInsertInfo": _lib.InsertInfo}
    jl2c = {"Cdouble": "c_double", "Int64": "c_long", "Int32": "c_int"}
    for jname, cstruct in pairs.items():
        m = re.search(r"struct %s;\s*(.*?)\s*end" % jname, SRC)
        assert m, jname
        fields = [f.strip() for f in m.group(1).split(";") if f.strip()]
        names = [f.split("::")[0] for f in fields]
        types = [f.split("::")[1] for f in fields]
        assert names == [n for n, _ in cstruct._fields_], (jname, names)
        for t, (_, ct) in zip(types, cstruct._fields_):
            assert jl2c[t] == ct.__name__, (jname, t, ct)
