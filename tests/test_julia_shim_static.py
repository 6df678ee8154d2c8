"""Julia is not installed in this image, so `julia/DynamicOEDCuda.jl` (the `ccall` binding that sits
at the reference's drop-in seam, /root/reference/src/problem.jl:96-101) cannot run here.  It is
therefore checked MECHANICALLY against the ABI it binds: struct field order / widths, every `ccall`
tuple against the header prototype, the integer codes, and the JSON keys `julia/export_ir.jl` writes
against what `ir.py` reads.  One transposed field would segfault the first real user."""
import ctypes as C
import os
import re

from dynamicoed.jl_b200 import ir, models, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "dynoed.h")).read()
SHIM = open(os.path.join(ROOT, "julia", "DynamicOEDCuda.jl")).read()
EXPORT = open(os.path.join(ROOT, "julia", "export_ir.jl")).read()

# Julia type <-> C type <-> ctypes type
JL_FIELD = {"UInt32": ("uint32_t", C.c_uint32), "Int32": ("int32_t", C.c_int32), "Cstring": ("const char*", C.c_char_p),
            "Ptr{Float64}": ("const double*", C.POINTER(C.c_double)), "Ptr{Int32}": ("const int32_t*", C.POINTER(C.c_int32))}
# what a ccall argument type may be for each C parameter type
JL_ARG = {
    "dynoed_ctx*": {"Ptr{Cvoid}"}, "const dynoed_ctx*": {"Ptr{Cvoid}"}, "void*": {"Ptr{Cvoid}"},
    "dynoed_ctx**": {"Ref{Ptr{Cvoid}}"}, "void**": {"Ref{Ptr{Cvoid}}"},
    "const dynoed_problem_desc*": {"Ref{ProblemDesc}"},
    "const double*": {"Ptr{Float64}"}, "double*": {"Ptr{Float64}", "Ref{Float64}"},
    "int64_t": {"Int64"}, "int64_t*": {"Ref{Int64}", "Ptr{Int64}"}, "uint32_t": {"UInt32"},
    "double": {"Cdouble", "Float64"}, "int": {"Cint"}, "size_t": {"Csize_t"},
}
JL_RET = {"int": "Cint", "void": "Cvoid", "const char*": "Cstring"}


def _norm_c(t):
    t = re.sub(r"\s+", " ", t.strip())
    return t.replace(" *", "*").replace("* ", "*")


def c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(.*?)(\w+)$", decl)
        out.append((m.group(2), _norm_c(m.group(1))))
    return out


def c_prototypes():
    protos = {}
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    for m in re.finditer(r"^\s*([\w\s\*]+?)\s*\b(dynoed_\w+)\s*\(([^)]*)\)\s*;", text, re.M):
        ret, name, args = _norm_c(m.group(1)), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append(_norm_c(mm.group(1)))
        protos[name] = (ret, params)
    return protos


def _split_top(s):
    """split a Julia tuple body at top-level commas (braces nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\[\]\),\s*(\w+),\s*\(", SHIM):
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(SHIM[j], 0)
            j += 1
        calls.append((m.group(1), m.group(2), _split_top(SHIM[i:j - 1])))
    return calls


def test_problem_desc_layout_matches_header_and_ctypes():
    jl = re.search(r"struct ProblemDesc\n(.*?)\nend", SHIM, re.S).group(1)
    jl_fields = [tuple(x.strip() for x in line.split("::")) for line in jl.strip().splitlines()]
    c_fields = c_struct_fields("dynoed_problem_desc")
    py_fields = runtime.ProblemDesc._fields_
    assert len(jl_fields) == len(c_fields) == len(py_fields) == 20
    for (jn, jt), (cn, ct), (pn, pt) in zip(jl_fields, c_fields, py_fields):
        assert jn == cn == pn, (jn, cn, pn)
        want_c, want_py = JL_FIELD[jt]
        assert ct == want_c, (jn, jt, ct)
        assert pt is want_py or pt == want_py, (jn, jt, pt)
    # Julia lays an isbits struct out like C (natural alignment): recompute the offsets from the
    # Julia field types and compare them with what ctypes (= the C compiler's rule) uses
    off = 0
    for jn, jt in jl_fields:
        size = 4 if jt in ("UInt32", "Int32") else 8
        off = (off + size - 1) // size * size
        assert getattr(runtime.ProblemDesc, jn).offset == off, (jn, off)
        off += size
    assert C.sizeof(runtime.ProblemDesc) == (off + 7) // 8 * 8


def test_every_ccall_matches_its_prototype():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert {c[0] for c in calls} >= {"dynoed_create", "dynoed_destroy", "dynoed_eval_batch", "dynoed_hessian_batch",
                                     "dynoed_eval_batch_host_async", "dynoed_wait", "dynoed_argmin",
                                     "dynoed_last_error", "dynoed_host_alloc", "dynoed_host_free"}
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in include/dynoed.h"
        cret, cparams = protos[name]
        assert JL_RET[cret] == ret, (name, cret, ret)
        assert len(args) == len(cparams), (name, args, cparams)
        for a, cp in zip(args, cparams):
            assert a in JL_ARG[cp], (name, a, cp)


def test_every_header_symbol_is_exported_and_bound_by_the_python_host():
    protos = c_prototypes()
    assert set(protos) == set(runtime.EXPORTS), set(protos) ^ set(runtime.EXPORTS)


def test_integer_codes_match_the_defines():
    defs = {m.group(1): int(m.group(2).rstrip("u")) for m in re.finditer(r"#define (DYNOED_\w+) \(?(-?\d+u?)\)?", HEADER)}
    methods = dict(re.findall(r":(\w+) => Cint\((\d+)\)", re.search(r"const METHODS = Dict\((.*?)\)\n", SHIM, re.S).group(1)))
    assert {k: int(v) for k, v in methods.items()} == {
        "rk4": defs["DYNOED_RK4"], "tsit5": defs["DYNOED_TSIT5"], "ros2": defs["DYNOED_ROS2"],
        "ros3p": defs["DYNOED_ROS3P"], "rodas3": defs["DYNOED_RODAS3"]}
    crit = dict(re.findall(r"criterion_id\(::(\w+)Criterion\) = Cint\((\d+)\)", SHIM))
    assert {k: int(v) for k, v in crit.items()} == {
        "FisherA": defs["DYNOED_FISHER_A"], "FisherD": defs["DYNOED_FISHER_D"], "FisherE": defs["DYNOED_FISHER_E"],
        "A": defs["DYNOED_A"], "D": defs["DYNOED_D"], "E": defs["DYNOED_E"], "LogD": defs["DYNOED_LOG_D"]}
    assert int(re.search(r"const DYNOED_OK = Cint\((\d+)\)", SHIM).group(1)) == defs["DYNOED_OK"]
    # the Python host uses the same numbers
    assert (runtime.RK4, runtime.TSIT5, runtime.ROS2, runtime.ROS3P, runtime.RODAS3) == (0, 1, 2, 3, 4)
    assert runtime.GRAD_NO_CONTROLS == defs["DYNOED_GRAD_NO_CONTROLS"]


def test_shim_keeps_the_reference_signature():
    """/root/reference/src/problem.jl:96-101: (prob, AD, u0 = get_initial_variables(prob), p = NullParameters();
    integer_constraints = false, constraints = nothing, variable_type = Float64, kwargs...)."""
    sig = re.search(r"function Optimization\.OptimizationProblem\((.*?)\) where \{T\}", SHIM, re.S).group(1)
    sig = re.sub(r"\s+", " ", sig)
    for piece in ("prob::OEDProblem", "backend::CudaBackend", "u0::ComponentVector = DynamicOED.get_initial_variables(prob)",
                  "p = SciMLBase.NullParameters()", "integer_constraints::Bool = false", "constraints = nothing",
                  "variable_type::Type{T} = Float64", "kwargs..."):
        assert piece in sig, piece
    ref = open("/root/reference/src/problem.jl").read() if os.path.exists("/root/reference/src/problem.jl") else None
    if ref:   # only in the build container: the argument names of the method being shadowed
        rs = re.sub(r"\s+", " ", re.search(r"function Optimization\.OptimizationProblem\((.*?)\) where \{T\}", ref, re.S).group(1))
        for piece in ("prob::OEDProblem", "integer_constraints::Bool = false", "constraints = nothing",
                      "variable_type::Type{T} = Float64", "kwargs..."):
            assert piece in rs, piece


def D_simplified():
    import dynamicoed.jl_b200 as D
    return D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing()))


def test_export_ir_writes_the_keys_ir_py_reads():
    keys = lambda section: set(re.findall(r'\\"(\w+)\\": ', section))
    MAIN = EXPORT[EXPORT.index("function export_ir("):]          # the user-system part of the file
    top = keys(MAIN)
    want = ir.system_to_ir(models.lotka_volterra_fishing())
    assert set(want) <= top, set(want) - top
    st = MAIN[MAIN.index('\\"states\\": ['):MAIN.index('\\"parameters\\": [')]
    pa = MAIN[MAIN.index('\\"parameters\\": ['):MAIN.index('\\"equations\\": [')]
    eq = MAIN[MAIN.index('\\"equations\\": ['):MAIN.index('\\"observed\\": [')]
    ob = MAIN[MAIN.index('\\"observed\\": ['):MAIN.index('augmented &&')]
    # the "augmented" section (reference-built simplified system) and its canonical identifiers
    AUG = EXPORT[EXPORT.index("function augmented_alias("):EXPORT.index("function export_ir(")]
    sec = ir.augmented_to_ir(D_simplified())
    assert keys(AUG) == {"equations"} | set(sec["equations"][0])
    assert "augmented" in top
    for ident in ('"x$(k-1)"', '"G_$(i)_$(j)"', '"F_$(i)_$(j)"', '"p$(k-1)"', '"ic$(get_initial_condition_id(v))"', '"w$(i)"'):
        assert ident in AUG, ident
    assert {e["state"] for e in sec["equations"]} == {"x0", "x1", "G_1_1", "G_2_1", "G_1_2", "G_2_2", "F_1_1", "F_1_2", "F_2_2"}
    assert keys(st) - {"states"} == set(want["states"][0])
    assert keys(pa) - {"parameters"} == set(want["parameters"][0])
    assert keys(eq) - {"equations"} == set(want["equations"][0])
    assert keys(ob) - {"observed"} == set(want["observed"][0])
    assert f'"{ir.FORMAT}"'.replace('"', '\\"') in EXPORT
