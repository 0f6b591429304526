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


def _strip_strings_and_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith('"""', i):
            i = src.index('"""', i + 3) + 3
            out.append(" ")
        elif src[i] == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            i = j + 1
            out.append('""')
        elif src[i] == "#":
            j = src.find("\n", i)
            i = j if j >= 0 else n
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def test_blocks_and_brackets_balance():
    """A coarse syntax lint (no Julia here): every bracket closes with its own kind, and outside brackets (where `end` is
    an index and `for` / `if` belong to comprehensions) every block opener has its `end`."""
    s = _strip_strings_and_comments(SRC)
    openers = {"function", "if", "for", "while", "struct", "module", "let", "do", "begin", "try", "quote", "macro"}
    stack, blocks = [], 0
    for m in re.finditer(r"[A-Za-z_][A-Za-z_0-9!]*|[\[\](){}]", s):
        t = m.group(0)
        line = s.count("\n", 0, m.start()) + 1
        if t in "[({":
            stack.append(t)
        elif t in "])}":
            assert stack and {"[": "]", "(": ")", "{": "}"}[stack.pop()] == t, f"bracket mismatch near stripped line {line}"
        elif not stack:
            if t in openers:
                blocks += 1
            elif t == "end":
                blocks -= 1
                assert blocks >= 0, f"`end` without an opener near stripped line {line}"
    assert not stack and blocks == 0, (stack, blocks)


# names of the reference's module the shim extends or calls (`ns.<name>`), each defined under /root/reference/src at the
# time of writing (file:line): the list is pinned here because the reference tree does not travel with the repository
REFERENCE_NAMES = {
    "applyexp": "src/applyexp.jl:62", "applyexp_sweep_printer": "src/applyexp.jl:50", "current_region": "src/iterators.jl:57",
    "current_time": "src/applyexp.jl:12", "default_expansion_factor": "src/subspace/subspace.jl", "default_max_expand": "src/subspace/subspace.jl",
    "eigenvalue": "src/eigsolve.jl", "eigsolve": "src/eigsolve.jl:45", "eigsolve_solver": "src/local_solvers/eigsolve.jl:3",
    "euler_sweep": "src/region_plans/euler_plans.jl:4", "exponentiate_solver": "src/local_solvers/exponentiate.jl",
    "extracter": "src/extracter.jl:3", "inserter": "src/inserter.jl:3", "next_region": "src/iterators.jl:62", "operator": "src/eigsolve.jl",
    "permute_indices": "src/permute_indices.jl", "problem": "src/iterators.jl:55", "process_real_times": "src/applyexp.jl:91",
    "region_plan": "src/iterators.jl:102", "runge_kutta_solver": "src/local_solvers/runge_kutta.jl:16", "state": "src/eigsolve.jl",
    "sweep_iterator": "src/iterators.jl:35", "sweep_solve": "src/sweep_solve.jl:10", "tdvp_regions": "src/region_plans/tdvp_region_plans.jl:35",
    "truncation_parameters": "src/truncation_parameters.jl", "updater": "src/eigsolve.jl:14",
}


def test_every_reference_name_the_shim_uses_exists_in_the_reference():
    used = set(re.findall(r"\bns\.([A-Za-z_][A-Za-z_0-9!]*)", _strip_strings_and_comments(SRC)))
    assert used <= set(REFERENCE_NAMES), used - set(REFERENCE_NAMES)
    ref = "/root/reference/src"
    if os.path.isdir(ref):        # in the build container the pinned list is re-checked against the tree itself
        text = ""
        for d, _, files in os.walk(ref):
            for f in files:
                if f.endswith(".jl"):
                    text += open(os.path.join(d, f)).read() + "\n"
        for nm in used:
            assert re.search(r"(?:^|\n)\s*(?:function\s+)?%s\s*\(" % re.escape(nm), text) or re.search(r"%s\(.*\)\s*=" % re.escape(nm), text), nm


def test_the_shim_does_not_overwrite_reference_methods():
    """Methods added to the reference's functions must dispatch on this module's own types: an untyped redefinition would
    replace the reference's method for every problem type (the sweep printers are therefore separate functions here)."""
    s = _strip_strings_and_comments(SRC)
    defs = list(re.finditer(r"(?:^|\n)(?:function\s+)?ns\.(\w+)\(([^;)]*)", s))      # definitions start in column 0 in this file
    assert len(defs) >= 11
    for m in defs:
        name, first_args = m.group(1), m.group(2)
        assert "::B200" in first_args or "::Union{B200" in first_args, (name, first_args)


def test_every_ccall_names_its_symbol_literally():
    """Julia resolves `ccall((:symbol, lib), ...)` at compile time: the symbol must be a literal, not a variable."""
    s = _strip_strings_and_comments(SRC)
    assert len(re.findall(r"ccall\(\(", s)) >= 40
    assert re.findall(r"ccall\(\((?!:nsb_\w+, lib\))", s) == []
