"""Model interchange ("dynoed-ir/1"): the UN-augmented system exactly as the user declares it in
the reference — states, parameters with their metadata (``tunable``, ``input``, ``bounds``,
``measurement_rate``, integer-ness; /root/reference/src/augment.jl:11-96), equations
``D(x) ~ rhs`` / ``0 ~ g``, observed equations, ``tspan`` — as JSON with infix expression strings.

This is the hand-over point between a Julia host and the kernel emitter: `julia/export_ir.jl`
writes the file from a ModelingToolkit ``ODESystem`` (unexecuted here — no Julia in the image),
and

    python -m dynamicoed.jl_b200.ir build model.json

augments it (``OEDSystem`` -> ``structural_simplify``, same equations as
/root/reference/src/augment.jl:128-244), prints the right-hand side into the sm_100a kernel
template (`emit.py`) and compiles it ONCE into ``libdynoed_<hash>.so``; the Julia shim then only
``ccall``s that library.  `system_to_ir` / `system_from_ir` round-trip every built-in model
(tests/test_ir.py), which is what pins the format.
"""
from __future__ import annotations

import json
import sys

import sympy as sp

from .symbolic import (Differential, Eq, Num, ODESystem, getbounds, getdefault,
                       get_measurement_rate, independent_variable, is_measured, isinput, istunable)

FORMAT = "dynoed-ir/1"


def _expr_str(e, names):
    return sp.sstr(sp.sympify(e).xreplace(names))


def system_to_ir(sys_: ODESystem) -> dict:
    """JSON-able description of an un-augmented `ODESystem`."""
    t = sys_.iv
    plain = {}           # symbols with characters that are not identifier-safe get safe aliases

    def alias(v, prefix, k):
        plain[v] = sp.Symbol(f"{prefix}{k}")
        return plain[v].name
    states, params = [], []
    for k, v in enumerate(sys_.get_states()):
        lo, hi = getbounds(v)
        states.append(dict(id=alias(v, "x", k), name=v.name, default=getdefault(v),
                           tunable=bool(istunable(v, False)), bounds=[lo, hi]))
    for k, v in enumerate(sys_.parameters()):
        lo, hi = getbounds(v)
        rate = get_measurement_rate(v) if is_measured(v) else None
        params.append(dict(id=alias(v, "p", k), name=v.name, default=getdefault(v),
                           tunable=bool(istunable(v, False)), input=bool(isinput(v)),
                           bounds=[lo, hi], measurement_rate=rate,
                           rate_is_count=isinstance(rate, int) and not isinstance(rate, bool),
                           integer=bool(v.meta.get("integer", False))))
    plain[t] = sp.Symbol("t")
    eqs = []
    for e in sys_.eqs:
        lhs = sp.sympify(e.lhs)
        if isinstance(lhs, Num) and lhs.role == "derivative":
            eqs.append(dict(lhs=f"D({plain[lhs.of].name})", rhs=_expr_str(e.rhs, plain)))
        else:
            eqs.append(dict(lhs=_expr_str(lhs, plain), rhs=_expr_str(e.rhs, plain)))
    obs = []
    for e in sys_.observed:
        y = e.lhs
        rate = get_measurement_rate(y)
        obs.append(dict(name=y.name, rhs=_expr_str(e.rhs, plain), measurement_rate=rate,
                        rate_is_count=isinstance(rate, int) and not isinstance(rate, bool)))

    def clean(d):
        for k in ("bounds",):
            d[k] = [None if (b is None or b in (float("inf"), float("-inf"))) else float(b) for b in d[k]]
        return d
    return dict(format=FORMAT, name=sys_.name, iv="t", tspan=list(sys_.tspan) if sys_.tspan else None,
                states=[clean(s) for s in states], parameters=[clean(p) for p in params],
                equations=eqs, observed=obs)


def _rate(r, is_count):
    if r is None:
        return None
    return int(r) if is_count else float(r)


def system_from_ir(ir: dict) -> ODESystem:
    assert ir.get("format") == FORMAT, f"unknown IR format {ir.get('format')!r}"
    t = independent_variable(ir.get("iv", "t"))
    local = {"t": t}
    states, params = [], []
    for s in ir["states"]:
        meta = dict(default=s.get("default"), tunable=bool(s.get("tunable", False)))
        b = s.get("bounds") or [None, None]
        if b[0] is not None or b[1] is not None:
            meta["bounds"] = (-sp.oo if b[0] is None else b[0], sp.oo if b[1] is None else b[1])
            meta["bounds"] = tuple(float(v) for v in meta["bounds"])
        v = Num(s["name"], "state", **meta)
        states.append(v)
        local[s["id"]] = v
    for p in ir["parameters"]:
        meta = dict(default=p.get("default"), tunable=bool(p.get("tunable", False)))
        if p.get("input"):
            meta["input"] = True
        if p.get("integer"):
            meta["integer"] = True
        r = _rate(p.get("measurement_rate"), p.get("rate_is_count", False))
        if r is not None:
            meta["measurement_rate"] = r
        b = p.get("bounds") or [None, None]
        if b[0] is not None or b[1] is not None:
            meta["bounds"] = (float("-inf") if b[0] is None else float(b[0]),
                              float("inf") if b[1] is None else float(b[1]))
        v = Num(p["name"], "parameter", **meta)
        params.append(v)
        local[p["id"]] = v
    D = Differential(t)
    eqs = []
    for e in ir["equations"]:
        rhs = sp.sympify(e["rhs"], locals=local)
        lhs = e["lhs"].strip()
        if lhs.startswith("D(") and lhs.endswith(")"):
            eqs.append(Eq(D(local[lhs[2:-1].strip()]), rhs))
        else:
            eqs.append(Eq(sp.sympify(lhs, locals=local), rhs))
    observed = []
    for o in ir["observed"]:
        y = Num(o["name"], "observed", measurement_rate=_rate(o["measurement_rate"], o.get("rate_is_count", False)))
        observed.append(Eq(y, sp.sympify(o["rhs"], locals=local)))
    return ODESystem(eqs, t, states=states, ps=params, tspan=ir.get("tspan"), observed=observed,
                     name=ir.get("name", "sys"))


# ----------------------------------------------------------------------------------------------
# Optional section "augmented": the explicit right-hand sides of the REFERENCE-built system,
# `full_equations(structural_simplify(OEDSystem(sys)))`, one entry per differential state, written by
# julia/export_ir.jl with canonical identifiers that depend only on positions the reference fixes
# (/root/reference/src/augment.jl:233-243: states = [x; vec(G); F[triu]; vec(Q)], parameters =
# [p; unknown initial conditions; w]):
#     x<k>      k-th state of the user system (0-based, as in "states")
#     G_<i>_<j> sensitivity of state i w.r.t. tunable j (1-based, the reference's G[i,j])
#     F_<i>_<j> Fisher entry (i <= j, 1-based)
#     p<k>      k-th parameter of the user system (0-based, as in "parameters")
#     ic<i>     unknown initial condition of state i (1-based)          w<i>  measurement weight i (1-based)
# When the section is present, `build_from_ir` refuses to emit a kernel unless the host's own augmentation
# (augment.py) reproduces every one of those right-hand sides: the reference's simplified system is the
# authority for the equations that reach the kernel template, the Python derivation only adds the
# structure (Jacobian blocks, adjoint) the emitter needs.
# ----------------------------------------------------------------------------------------------
def _canonical_symbols(simp, user_ir=None):
    """{our symbol: canonical identifier} for a SimplifiedOEDSystem."""
    aug = simp.aug
    names = {}
    for k, v in enumerate(aug.x):
        names[v] = sp.Symbol(f"x{k}")
    nx, np_ = aug.G.shape
    for i in range(nx):
        for j in range(np_):
            names[aug.G[i, j]] = sp.Symbol(f"G_{i + 1}_{j + 1}")
    for i in range(np_):
        for j in range(i, np_):
            names[aug.F[i, j]] = sp.Symbol(f"F_{i + 1}_{j + 1}")
    for k, v in enumerate(aug.base.parameters()):
        names[v] = sp.Symbol(f"p{k}")
    for v, i in zip(aug.x_ic, aug.unknown_initial_conditions):
        names[v] = sp.Symbol(f"ic{i}")
    for i, v in enumerate(aug.w, start=1):
        names[v] = sp.Symbol(f"w{i}")
    names[aug.iv] = sp.Symbol("t")
    return names


def augmented_to_ir(simp) -> dict:
    """The "augmented" section as THIS host derives it (what export_ir.jl writes from the reference's
    own `structural_simplify(OEDSystem(sys))`)."""
    names = _canonical_symbols(simp)
    eqs = []
    for v, r in zip(simp.x, simp.f):
        eqs.append(dict(state=names[v].name, rhs=_expr_str(r, names)))
    for j in range(simp.np_):
        for i in range(simp.nx):
            eqs.append(dict(state=names[simp.G[i, j]].name, rhs=_expr_str(simp.dG[i, j], names)))
    for v, r in zip(simp.F_triu, simp.dF):
        eqs.append(dict(state=names[v].name, rhs=_expr_str(r, names)))
    return dict(equations=eqs)


def check_augmented(section: dict, simp, rtol: float = 1e-10, points: int = 4) -> int:
    """Compare the reference-exported right-hand sides with this host's, numerically at random points (the two
    sides differ in expression form, not in value).  Raises ValueError naming the first disagreement."""
    import random
    ours = {e["state"]: sp.sympify(e["rhs"]) for e in augmented_to_ir(simp)["equations"]}
    theirs = {}
    for e in section["equations"]:
        theirs[e["state"]] = sp.sympify(e["rhs"].replace("^", "**"))
    missing = sorted(set(ours) - set(theirs))
    extra = sorted(set(theirs) - set(ours))
    if missing or extra:
        raise ValueError(f"augmented systems differ in their differential states: the reference's system lacks "
                         f"{missing}, this host's lacks {extra}")
    rng = random.Random(20262)
    syms = sorted(set().union(*[e.free_symbols for e in ours.values()], *[e.free_symbols for e in theirs.values()]),
                  key=lambda v: v.name)
    for _ in range(points):
        at = {v: rng.uniform(0.5, 1.5) for v in syms}
        for st in ours:
            a, b = float(ours[st].evalf(30, subs=at)), float(theirs[st].evalf(30, subs=at))
            if abs(a - b) > rtol * max(1.0, abs(a), abs(b)):
                raise ValueError(f"augmented right-hand side of {st} differs from the reference's: "
                                 f"{ours[st]}  vs  {theirs[st]}")
    return len(ours)


def build_from_ir(path: str) -> dict:
    """IR file -> emitted model -> compiled library.  Returns {library, model, n_var, ...}."""
    from .augment import OEDSystem, structural_simplify
    from .problem import _resolve_model
    with open(path) as fh:
        ir = json.load(fh)
    simp = structural_simplify(OEDSystem(system_from_ir(ir)))
    checked = check_augmented(ir["augmented"], simp) if ir.get("augmented") else 0
    lib, name, info = _resolve_model(simp)
    from . import runtime
    return dict(library=lib or runtime.LIB_PATH, model=name, nx=info["nx"], np=info["np"],
                n_obs=info["n_obs"], n_ctrl=info["n_ctrl"], n_ic=info["n_ic"],
                theta_names=info["theta_names"], state_names=info["state_names"],
                augmented_equations_checked=checked)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "build":
        print(json.dumps(build_from_ir(sys.argv[2]), ensure_ascii=False))
    elif len(sys.argv) == 3 and sys.argv[1] == "export":
        from . import models
        from .augment import OEDSystem, structural_simplify
        sys0 = models.BUILTIN[sys.argv[2]]()
        out = system_to_ir(sys0)
        out["augmented"] = augmented_to_ir(structural_simplify(OEDSystem(sys0)))
        print(json.dumps(out, indent=1, ensure_ascii=False))
    else:
        print("usage: python -m dynamicoed.jl_b200.ir build <model.json> | export <builtin name>")
