"""Minimal symbolic front-end (sympy based) standing in for the ModelingToolkit surface the
reference's hot path consumes.

The reference declares models with ``@variables`` / ``@parameters`` carrying metadata
(``tunable``, ``input``, ``bounds``, ``measurement_rate`` ...; see
/root/reference/src/augment.jl:11-96 and e.g. /root/reference/test/references/LotkaVolterra.jl:10-36).
This module provides the same vocabulary on top of sympy so that the augmentation
(`augment.py`), the grid compiler (`discretize.py`) and the CUDA emitter (`emit.py`) can be
written against it.  It is host-side, one-off (cold path) code.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Iterable, Sequence

import sympy as sp

_counter = itertools.count()


class Num(sp.Symbol):
    """A sympy symbol with ModelingToolkit-like metadata.

    kind: "state" | "parameter" | "derivative" | "observed" | "iv"
    meta keys (all optional): default, tunable, input, bounds, measurement_rate, description,
    integer, measurement_function, fisher_state, unweighted_information_gain, variable_ic
    """

    def __new__(cls, name, kind="parameter", uid=None, **meta):
        obj = sp.Symbol.__xnew__(cls, name, real=True)
        obj.role = kind
        obj.meta = dict(meta)
        obj.order = next(_counter)
        # identity = (name, declaration): two models may both declare a "p[1]" with different
        # metadata; they must not compare equal (sympy caches expressions by equality).
        obj.uid = uid if uid is not None else obj.order
        return obj

    def _hashable_content(self):
        return super()._hashable_content() + (self.uid,)

    # sympy re-creates atoms through __getnewargs_ex__ on pickling only; identity is kept in
    # subs/xreplace, so metadata survives every manipulation used in this package.
    def with_meta(self, **meta):
        new = Num(self.name, self.role, **{**self.meta, **meta})
        return new


def _names(spec) -> list[str]:
    if isinstance(spec, str):
        return [s for s in spec.replace(",", " ").split() if s]
    return list(spec)


def variables(spec, defaults=None, **meta):
    """``@variables x(t)=0.5 y(t)=0.7 [meta...]`` -> state variables."""
    names = _names(spec)
    defaults = _broadcast(defaults, len(names))
    out = [Num(n, "state", **_with_default(meta, d)) for n, d in zip(names, defaults)]
    return out[0] if len(out) == 1 and isinstance(spec, str) and len(names) == 1 else out


def parameters(spec, defaults=None, **meta):
    """``@parameters p[1:4]=[...] [meta...]`` -> parameters (default: not tunable, not input)."""
    names = _names(spec)
    defaults = _broadcast(defaults, len(names))
    out = [Num(n, "parameter", **_with_default(meta, d)) for n, d in zip(names, defaults)]
    return out[0] if len(out) == 1 and isinstance(spec, str) and len(names) == 1 else out


def observed_variables(spec, **meta):
    """``@variables obs(t)[1:2] [measurement_rate=0.1]`` -> left-hand sides of observed eqs."""
    names = _names(spec)
    out = [Num(n, "observed", **meta) for n in names]
    return out[0] if len(out) == 1 and isinstance(spec, str) and len(names) == 1 else out


def _broadcast(d, n):
    if d is None:
        return [None] * n
    if isinstance(d, (int, float)):
        return [d] * n
    d = list(d)
    assert len(d) == n, "number of defaults does not match number of names"
    return d


def _with_default(meta, d):
    m = dict(meta)
    if d is not None:
        m["default"] = d
    return m


def independent_variable(name="t"):
    return Num(name, "iv")


class Differential:
    """``D = Differential(t); D(x)`` -> a marker symbol for dx/dt."""

    def __init__(self, iv):
        self.iv = iv

    def __call__(self, x):
        assert isinstance(x, Num), "Differential applies to a declared variable"
        return _derivative(x)


def _derivative(x: "Num") -> "Num":
    d = Num(f"D({x.name})", "derivative", uid=("D", x.uid))
    d.of = x
    return d


@dataclass(frozen=True)
class Equation:
    lhs: sp.Expr
    rhs: sp.Expr

    def residual(self):
        # the reference's convention, /root/reference/src/augment.jl:116: rhs - lhs
        return sp.sympify(self.rhs) - sp.sympify(self.lhs)


def Eq(lhs, rhs) -> Equation:
    return Equation(sp.sympify(lhs), sp.sympify(rhs))


# --- metadata accessors (names follow /root/reference/src/augment.jl:15-96 and MTK 8.72) -------

def getdefault(x, default=None):
    return x.meta.get("default", default)


def istunable(x, default=False):
    # ModelingToolkit 8.x: variables without `tunable` metadata are NOT tunable.
    return bool(x.meta.get("tunable", default))


def isinput(x):
    return bool(x.meta.get("input", False))


def getbounds(x):
    return tuple(x.meta.get("bounds", (-float("inf"), float("inf"))))


def is_measured(x):
    return "measurement_rate" in x.meta


def get_measurement_rate(x):
    return x.meta.get("measurement_rate", 0.0)


def is_measurement_function(x):
    return bool(getattr(x, "meta", {}).get("measurement_function", False))


def is_fisher_state(x):
    return bool(getattr(x, "meta", {}).get("fisher_state", False))


def is_information_gain(x):
    return bool(getattr(x, "meta", {}).get("unweighted_information_gain", False))


def is_initial_condition(x):
    return "variable_ic" in getattr(x, "meta", {})


def get_initial_condition_id(x):
    return x.meta.get("variable_ic", 0)


def _symbols_in(exprs: Iterable[sp.Expr]) -> list[Num]:
    seen = {}
    for e in exprs:
        for s in sp.sympify(e).free_symbols:
            if isinstance(s, Num):
                seen[s] = s
    return sorted(seen.values(), key=lambda v: v.order)


class ODESystem:
    """A first-order ODE / semi-explicit DAE system ``D(x_i) ~ f_i`` and ``0 ~ g_j`` with
    observed equations ``y_k ~ h_k(x)``.

    States and parameters are discovered from the equations and ordered by declaration order
    unless given explicitly (``states=``, ``ps=``).
    """

    def __init__(self, eqs: Sequence[Equation], t: Num | None = None, states=None, ps=None,
                 *, tspan=None, observed: Sequence[Equation] = (), name="sys", defaults=None):
        self.eqs = list(eqs)
        self.iv = t if t is not None else independent_variable("t")
        self.tspan = tuple(float(v) for v in tspan) if tspan is not None else None
        self.observed = list(observed)
        self.name = name
        syms = _symbols_in([e.lhs for e in self.eqs] + [e.rhs for e in self.eqs] +
                           [e.rhs for e in self.observed])
        derivs = [s for s in syms if s.role == "derivative"]
        found_states = {s: s for s in syms if s.role == "state"}
        for d in derivs:
            found_states.setdefault(d.of, d.of)
        self.states = list(states) if states is not None else sorted(
            found_states.values(), key=lambda v: v.order)
        self.ps = list(ps) if ps is not None else [s for s in syms if s.role == "parameter"]
        self.defaults = dict(defaults or {})
        for v in self.states + self.ps:
            if v not in self.defaults and getdefault(v) is not None:
                self.defaults[v] = getdefault(v)

    # ModelingToolkit-style accessors
    def equations(self):
        return self.eqs

    def parameters(self):
        return self.ps

    def get_states(self):
        return self.states

    def derivative_of(self, x: Num) -> Num:
        return _derivative(x)
