"""Symbolic sensitivity/Fisher augmentation of an ODE/DAE system (host side, cold path).

Mirrors the behaviour of /root/reference/src/augment.jl:
  * ``construct_jacobians``      (:113-126)  residual convention rhs - lhs, f_xdot, f_x, f_p, h_x
  * ``build_augmented_system``   (:128-244)  G0 incl. unknown-IC block (:182-188), F' (:223-226),
                                              sensitivity DAE 0 = f_xdot G' + f_x G + f_p (:231),
                                              state order [x; vec(G); F[triu]; vec(Q)] (:240),
                                              parameter order [p U p_tunable; w] (:241)
  * ``OEDSystem``                (:246-248)
and of the ModelingToolkit passes the reference applies afterwards (``structural_simplify``,
``dae_index_lowering``; call sites /root/reference/test/references/Rober.jl:23), restricted to
what the hot path needs: explicit tearing of algebraic states.

It is written from the equations, not translated: the result is an explicit, substitution-free
set of sympy expressions that `emit.py` prints into the CUDA kernel template.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import sympy as sp

from .symbolic import (_derivative, Equation, ODESystem, Num, get_measurement_rate, getbounds,
                       getdefault, is_measured, isinput, istunable)

_SUB = str.maketrans("0123456789", "₀₁₂₃₄₅₆₇₈₉")


def _sub(i: int) -> str:
    return str(i).translate(_SUB)


def triu_pairs(n: int):
    """(i, j) pairs, 0-based, of the upper triangle in column-major order — the packing
    ``x[j(j-1)/2 + i]`` of /root/reference/src/fisher.jl:3-5."""
    return [(i, j) for j in range(n) for i in range(j + 1)]


def construct_jacobians(sys: ODESystem, p):
    """f_xdot, f_x, f_p, h_x of the residuals ``rhs - lhs`` (augment.jl:113-126)."""
    res = sp.Matrix([e.residual() for e in sys.eqs])
    x = sys.states
    dx = [sys.derivative_of(xi) for xi in x]
    fx = res.jacobian(x)
    dfddx = res.jacobian(dx)
    fp = res.jacobian(list(p)) if len(p) else sp.zeros(len(sys.eqs), 0)
    obs = [e.rhs for e in sys.observed] if sys.observed else list(x)
    hx = sp.Matrix(obs).jacobian(x)
    return dfddx, fx, fp, hx


@dataclass
class AugmentedSystem:
    """Result of ``OEDSystem(sys)``: the (un-simplified) augmented DAE."""
    name: str
    base: ODESystem
    iv: Num
    x: list
    p_tunable: list
    controls: list
    x_ic: list                      # initial-condition parameters (appended to p_tunable)
    unknown_initial_conditions: list  # 1-based state indices, like the reference
    w: list
    G: sp.Matrix
    F: sp.Matrix
    Q: sp.Matrix
    G_init: sp.Matrix
    hx: sp.Matrix
    eqs: list
    states: list
    ps: list
    observed: list
    tspan: tuple | None
    defaults: dict = field(default_factory=dict)

    def equations(self):
        return self.eqs

    def parameters(self):
        return self.ps

    def get_states(self):
        return self.states

    @property
    def np_(self):
        return len(self.p_tunable)

    @property
    def nx(self):
        return len(self.x)

    @property
    def n_obs(self):
        return len(self.w)


def build_augmented_system(sys: ODESystem, name: str | None = None) -> AugmentedSystem:
    T = float
    p = sys.parameters()
    c = [q for q in p if isinput(q)]
    p_tunable = [q for q in p if istunable(q) and q not in c]
    x = sys.get_states()
    x_ic, unknown_ic = [], []
    for i, xi in enumerate(x, start=1):
        if istunable(xi):
            xi0 = Num(f"{xi.name}₀", "parameter", default=getdefault(xi), tunable=True,
                           bounds=getbounds(xi), variable_ic=i,
                           description=f"Initial condition of state {xi.name}")
            unknown_ic.append(i)
            x_ic.append(xi0)
            p_tunable.append(xi0)
    t = sys.iv
    obs = sys.observed
    assert len(obs) > 0, "Please defined `observed` equations to use optimal experimental design."
    assert all(is_measured(e.lhs) for e in obs), \
        "Not all observed equations have measurement rates associated to them!"
    assert all(is_measured(ci) for ci in c), \
        "Not all controls have rates associated to them!"
    np_, nx, n_obs = len(p_tunable), len(x), len(obs)
    dfx, fx, fp, hx = construct_jacobians(sys, p_tunable)
    assert fx.shape[0] == fp.shape[0] == nx, "The size of the state equations and the jacobian does not match"

    G_init = sp.zeros(nx, np_)
    if x_ic:
        # literal restatement of augment.jl:186 (quirk Q3 of SURVEY.md: column np-(i-1) for state i)
        rows = [i - 1 for i in unknown_ic]
        cols = [np_ - (i - 1) - 1 for i in unknown_ic]
        if any(cc < 0 for cc in cols) or rows != list(range(len(rows))):
            warnings.warn("unknown initial conditions are not the leading states: the reference's "
                          "G0 column rule (augment.jl:186) is only a permutation in that case")
        for a, r in enumerate(rows):
            for b, cc in enumerate(cols):
                G_init[r, cc] = 1 if a == b else 0

    F = sp.Matrix(np_, np_, lambda i, j: Num(f"F[{i+1},{j+1}]", "state", default=0.0,
                                                   fisher_state=True,
                                                   description="Fisher Information Matrix"))
    G = sp.Matrix(nx, np_, lambda i, j: Num(f"G[{i+1},{j+1}]", "state",
                                                  default=T(G_init[i, j]),
                                                  description="Sensitivity State"))
    Q = sp.Matrix(n_obs, np_, lambda i, j: Num(
        f"Q[{i+1},{j+1}]", "state", unweighted_information_gain=True,
        description="Unweighted Fisher Information Derivative"))
    w = []
    for i, e in enumerate(obs, start=1):
        w.append(Num(f"w{_sub(i)}", "parameter", default=0.0, tunable=True, integer=True,
                          measurement_function=True, bounds=(0, 1),
                          measurement_rate=get_measurement_rate(e.lhs),
                          description=f"Measurement function of {e.lhs.name}"))

    idx = triu_pairs(np_)
    D = sys.derivative_of
    dF = sp.zeros(np_, np_)
    for i, wi in enumerate(w):
        row = hx[i:i + 1, :] * G
        dF += wi * (row.T * row)
    new_obs = [Equation(_deriv(F[i, j]), dF[i, j]) for (i, j) in idx]

    dG = sp.Matrix(nx, np_, lambda i, j: _deriv(G[i, j]))
    sens = dfx * dG + fx * G + fp
    sens_eqs = [Equation(sp.Integer(0), sens[i, j]) for j in range(np_) for i in range(nx)]
    HG = hx * G
    q_eqs = [Equation(Q[i, j], HG[i, j]) for j in range(np_) for i in range(n_obs)]

    states = list(x) + [G[i, j] for j in range(np_) for i in range(nx)] + \
        [F[i, j] for (i, j) in idx] + [Q[i, j] for j in range(np_) for i in range(n_obs)]
    ps = list(p) + [q for q in p_tunable if q not in p] + w
    defaults = dict(sys.defaults)
    for v in states + ps:
        if getdefault(v) is not None:
            defaults.setdefault(v, getdefault(v))
    return AugmentedSystem(name=name or sys.name, base=sys, iv=t, x=list(x), p_tunable=p_tunable,
                           controls=c, x_ic=x_ic, unknown_initial_conditions=unknown_ic, w=w,
                           G=G, F=F, Q=Q, G_init=G_init, hx=hx,
                           eqs=list(sys.eqs) + sens_eqs + new_obs + q_eqs,
                           states=states, ps=ps, observed=list(sys.observed), tspan=sys.tspan,
                           defaults=defaults)


def _deriv(v: Num) -> Num:
    return _derivative(v)


def OEDSystem(sys: ODESystem, name: str | None = None) -> AugmentedSystem:
    """augment.jl:246-248."""
    return build_augmented_system(sys, name=name)


def dae_index_lowering(sys: AugmentedSystem) -> AugmentedSystem:
    """The hot path only supports semi-explicit index-1 systems whose algebraic equations can be
    torn explicitly.  For those, index lowering is the identity; the
    solvability check happens in `structural_simplify`."""
    return sys


@dataclass
class SimplifiedOEDSystem:
    """Explicit ODE form of the augmented system after tearing (what ``ODAEProblem`` integrates,
    /root/reference/src/problem.jl:82): states ``[x_d; vec(G_d); F_triu]``."""
    name: str
    aug: AugmentedSystem
    iv: Num
    x: list                 # differential states
    x_index: list           # their 0-based indices in the original state list
    G: sp.Matrix            # nx_d x np symbols
    F_triu: list            # nF symbols
    f: sp.Matrix            # nx_d rhs
    dG: sp.Matrix           # nx_d x np rhs
    dF: list                # nF rhs (weighted by w)
    Qexpr: sp.Matrix        # n_obs x np, in terms of x_d, G_d
    states: list
    observed: list          # torn variables + Q + original observed
    alg_solution: dict
    u0: list                # default initial values for `states`
    # passthroughs
    p_tunable: list = field(default_factory=list)
    controls: list = field(default_factory=list)
    x_ic: list = field(default_factory=list)
    w: list = field(default_factory=list)
    ps: list = field(default_factory=list)
    tspan: tuple | None = None
    defaults: dict = field(default_factory=dict)

    @property
    def np_(self):
        return len(self.p_tunable)

    @property
    def nx(self):
        return len(self.x)

    @property
    def n_obs(self):
        return len(self.w)

    def equations(self):
        eqs = [Equation(_deriv(v), r) for v, r in zip(self.x, self.f)]
        eqs += [Equation(_deriv(self.G[i, j]), self.dG[i, j])
                for j in range(self.np_) for i in range(self.nx)]
        eqs += [Equation(_deriv(v), r) for v, r in zip(self.F_triu, self.dF)]
        return eqs

    def parameters(self):
        return self.ps

    def get_states(self):
        return self.states


def structural_simplify(aug: AugmentedSystem) -> SimplifiedOEDSystem:
    if isinstance(aug, SimplifiedOEDSystem):
        return aug
    base = aug.base
    x_all = aug.x
    nx_all, np_ = len(x_all), aug.np_
    dsyms = [base.derivative_of(xi) for xi in x_all]
    res = [e.residual() for e in base.eqs]
    # split rows into differential / algebraic
    diff_rows = [r for r, e in enumerate(res) if any(e.has(d) for d in dsyms)]
    alg_rows = [r for r in range(len(res)) if r not in diff_rows]
    diff_state_idx = [i for i, d in enumerate(dsyms) if any(e.has(d) for e in res)]
    alg_state_idx = [i for i in range(nx_all) if i not in diff_state_idx]
    assert len(alg_rows) == len(alg_state_idx), \
        "number of algebraic equations and algebraic states differ (higher-index DAE unsupported)"
    assert all(i + 1 not in aug.unknown_initial_conditions for i in alg_state_idx), \
        "unknown initial conditions on algebraic states are not supported"

    # 1. tear algebraic states
    alg_solution = {}
    if alg_rows:
        alg_states = [x_all[i] for i in alg_state_idx]
        sol = sp.solve([res[r] for r in alg_rows], alg_states, dict=True)
        assert len(sol) == 1, "algebraic equations must have a unique explicit solution"
        alg_solution = {k: sp.simplify(v) for k, v in sol[0].items()}
    # 2. solve differential rows for the derivatives
    dvars = [dsyms[i] for i in diff_state_idx]
    dsol = sp.solve([res[r] for r in diff_rows], dvars, dict=True)
    assert len(dsol) == 1, "differential equations must be uniquely solvable for D(x)"
    f = sp.Matrix([dsol[0][d] for d in dvars]).subs(alg_solution)

    # 3. sensitivities: solve 0 = f_xdot*G' + f_x*G + f_p for G' (differential rows) and for the
    #    algebraic rows of G (augment.jl:231)
    G = aug.G
    dGsym = sp.Matrix(nx_all, np_, lambda i, j: _deriv(G[i, j]))
    sens = [e.rhs for e in aug.eqs[len(base.eqs):len(base.eqs) + nx_all * np_]]
    sens = sp.Matrix(nx_all, np_, lambda i, j: sens[j * nx_all + i])
    G_alg_sol = {}
    if alg_rows:
        unknowns = [G[i, j] for i in alg_state_idx for j in range(np_)]
        eqs = [sens[r, j] for r in alg_rows for j in range(np_)]
        gs = sp.solve(eqs, unknowns, dict=True)
        assert len(gs) == 1
        G_alg_sol = gs[0]
    unknowns = [dGsym[i, j] for i in diff_state_idx for j in range(np_)]
    eqs = [sens[r, j] for r in diff_rows for j in range(np_)]
    gsol = sp.solve(eqs, unknowns, dict=True) if unknowns else [{}]
    assert len(gsol) == 1
    nxd = len(diff_state_idx)
    Gd = sp.Matrix(nxd, np_, lambda a, j: G[diff_state_idx[a], j])

    def clean(e):
        e = sp.sympify(e).subs(G_alg_sol).subs(alg_solution)
        return sp.expand(e) if e.is_polynomial() else e

    dG = sp.Matrix(nxd, np_, lambda a, j: clean(gsol[0][dGsym[diff_state_idx[a], j]]))
    # 4. F' and Q with torn variables substituted
    nF = np_ * (np_ + 1) // 2
    off = len(base.eqs) + nx_all * np_
    dF = [clean(e.rhs) for e in aug.eqs[off:off + nF]]
    Qexpr = (aug.hx * G).applyfunc(clean)
    F_triu = [aug.F[i, j] for (i, j) in triu_pairs(np_)]
    xd = [x_all[i] for i in diff_state_idx]
    states = xd + [Gd[a, j] for j in range(np_) for a in range(nxd)] + F_triu
    observed = list(aug.observed)
    observed += [Equation(k, v) for k, v in alg_solution.items()]
    observed += [Equation(k, clean(v)) for k, v in G_alg_sol.items()]
    observed += [Equation(aug.Q[i, j], Qexpr[i, j]) for j in range(np_) for i in range(aug.n_obs)]
    u0 = [aug.defaults.get(v, getdefault(v, 0.0)) for v in xd]
    u0 += [float(aug.G_init[diff_state_idx[a], j]) for j in range(np_) for a in range(nxd)]
    u0 += [0.0] * nF
    return SimplifiedOEDSystem(name=aug.name, aug=aug, iv=aug.iv, x=xd, x_index=diff_state_idx,
                               G=Gd, F_triu=F_triu, f=f, dG=dG, dF=dF, Qexpr=Qexpr,
                               states=states, observed=observed, alg_solution=alg_solution,
                               u0=u0, p_tunable=aug.p_tunable, controls=aug.controls,
                               x_ic=aug.x_ic, w=aug.w, ps=aug.ps, tspan=aug.tspan,
                               defaults=aug.defaults)
