"""Problem assembly and the optimiser boundary — the drop-in seam of the hot path.

Mirrors /root/reference/src/problem.jl:
  * ``OEDProblem(sys, criterion; alg, tspan, diffeqoptions)``          (:10-30)
  * ``states(::OEDProblem)``  symbolic optimisation variables           (:32-66)
  * ``get_timegrids``, ``get_initial_variables``                        (:68-78, :92-94)
  * ``build_predictor``  p -> (X, t)                                    (:80-90)
  * ``OptimizationProblem(prob, AD, u0, p; integer_constraints, constraints)`` (:96-161):
    objective closure (:112-121), constraints (:123-128), bounds (:131-134), integrality (:137-146)

Where the reference hands Optimization.jl an objective that is differentiated by ForwardDiff,
this hands the optimiser ``f`` AND ``grad`` closures that call ``libdynoed_cuda.so``; nothing in
here integrates or differentiates on the CPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import sympy as sp

from . import runtime
from .augment import SimplifiedOEDSystem, structural_simplify
from .criteria import AbstractInformationCriterion
from .discretize import (ComponentVector, Timegrid, design_layout, generate_initial_variables,
                         generate_variable_bounds)
from .symbolic import getdefault

_GEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "generated")


# --- fixed-step integrators standing in for the OrdinaryDiffEq `alg` argument -------------------
@dataclass(frozen=True)
class RK4:
    """Classic Runge-Kutta 4 with ``substeps`` equal steps per merged-grid interval."""
    substeps: int = 4
    method: int = runtime.RK4


@dataclass(frozen=True)
class Tsit5:
    """The Tsit5 tableau (the reference's default ``alg``, problem.jl:25) with FIXED steps."""
    substeps: int = 4
    method: int = runtime.TSIT5


@dataclass(frozen=True)
class ROS2:
    """Fixed-grid Rosenbrock method ROS2 (2 stages, order 2, L-stable) for the stiff / DAE
    configuration — stands where the reference passes ``Rodas5()``
    (/root/reference/test/references/Rober.jl:29).  ``substeps`` equal steps per merged interval,
    or ``edges``: one array of relative substep edges in [0, 1] per merged interval (graded grid)."""
    substeps: int = 4
    edges: tuple | None = None
    method: int = runtime.ROS2


@dataclass(frozen=True)
class ROS3P:
    """Fixed-grid Rosenbrock method ROS3P (3 stages, order 3, A-stable); see `ROS2`."""
    substeps: int = 4
    edges: tuple | None = None
    method: int = runtime.ROS3P


@dataclass(frozen=True)
class RODAS3:
    """Fixed-grid Rosenbrock method RODAS3 (4 stages, order 3, stiffly accurate, L-stable); see `ROS2`."""
    substeps: int = 4
    edges: tuple | None = None
    method: int = runtime.RODAS3


def graded_edges(n_intervals: int, n_first: int, t_min_rel: float, substeps: int) -> tuple:
    """Substep edges for a stiff start-up transient: ``n_first`` geometrically spaced steps from
    ``t_min_rel`` (relative to the first interval) up to its end, ``substeps`` equal steps in every
    later interval (Robertson)."""
    first = np.concatenate([[0.0], np.geomspace(t_min_rel, 1.0, n_first)])
    rest = np.arange(substeps + 1, dtype=np.float64) / substeps
    return tuple([first] + [rest] * (n_intervals - 1))


def _resolve_model(sys: SimplifiedOEDSystem):
    """(library path | None, model name, emitter info).  A system whose emitted code equals a
    compiled-in model uses ``libdynoed_cuda.so``; anything else is emitted and compiled once."""
    cached = getattr(sys, "_dynoed_model", None)
    if cached is not None:
        return cached
    from .emit import emit_model
    src, info = emit_model("USERMODEL", sys)
    found = None
    for f in sorted(os.listdir(_GEN)):
        if f.startswith("model_") and f.endswith(".cuh"):
            name = f[len("model_"):-len(".cuh")]
            if open(os.path.join(_GEN, f)).read() == src.replace("USERMODEL", name):
                found = (None, name, dict(info, name=name))
                break
    if found is None:
        from .build import build_user_model
        import hashlib
        name = "u" + hashlib.sha256(src.encode()).hexdigest()[:12]
        src = src.replace("USERMODEL", name)
        info = dict(info, name=name)
        found = (build_user_model(src, info), name, info)
    sys._dynoed_model = found
    return found


class OEDProblem:
    """problem.jl:10-30."""

    def __init__(self, system, objective: AbstractInformationCriterion, *, alg=None, tspan=None,
                 diffeqoptions=None, device: int = 0, **kwargs):
        system = structural_simplify(system)
        self.system = system
        self.objective = objective
        self.timegrid = Timegrid.from_system(system, tspan if tspan is not None else system.tspan)
        self.alg = alg if alg is not None else Tsit5()
        self.diffeq_options = dict(diffeqoptions or {})
        self.device = device
        self._ctx = None

    @property
    def layout(self):
        return design_layout(self.system, self.timegrid)

    def context(self) -> runtime.Context:
        """The GPU evaluator for this problem (created on first use)."""
        if self._ctx is None:
            libpath, name, info = _resolve_model(self.system)
            lay = self.layout
            nc = len(self.system.controls)
            self._ctx = runtime.Context(
                name, criterion=self.objective.id, method=self.alg.method,
                tgrid=self.timegrid.tpoints, indicators=self.timegrid.indicators - 1,
                ctrl_offsets=lay.control_offsets, w_offsets=lay.w_offsets,
                ic_offset=lay.ic_offset, tau_offset=lay.tau_offset, n_var=lay.n_var,
                substeps=self.alg.substeps, step_edges=getattr(self.alg, "edges", None),
                theta=info["theta"], x0=info["x0"], G0=info["G0"],
                device=self.device, library_path=libpath)
        return self._ctx


def states(prob: OEDProblem) -> ComponentVector:
    """Symbolic optimisation variables for constraint building (problem.jl:32-66)."""
    lay = prob.layout
    syms = np.empty(lay.n_var, dtype=object)
    for n, o, l in zip(lay.control_names, lay.control_offsets, lay.control_lengths):
        for k in range(l):
            syms[o + k] = sp.Symbol(f"{n}_{k+1}", real=True)
    for i, n in enumerate(lay.ic_names):
        syms[lay.ic_offset + i] = sp.Symbol(n, real=True)
    for n, o, l in zip(lay.w_names, lay.w_offsets, lay.w_lengths):
        for k in range(l):
            syms[o + k] = sp.Symbol(f"{n}_{k+1}", real=True)
    syms[lay.tau_offset] = sp.Symbol("τ", real=True)
    cv = ComponentVector.__new__(ComponentVector)
    object.__setattr__(cv, "data", syms)
    object.__setattr__(cv, "axes", lay.axes())
    return cv


def get_timegrids(prob: OEDProblem) -> dict:
    """problem.jl:68-78: {variable name: list of (t0, t1)} sorted by name."""
    g = prob.timegrid
    return {k: list(v) for k, v in sorted(zip(g.variables, g.timegrids))}


def get_initial_variables(prob: OEDProblem) -> ComponentVector:
    return generate_initial_variables(prob.system, prob.timegrid)


def build_predictor(prob: OEDProblem):
    """problem.jl:80-90: returns p -> (X, t) with X [n_aug, N+1] at the merged-grid points."""
    ctx = prob.context()
    t = prob.timegrid.tpoints

    def predictor(p):
        X = ctx.trajectory(np.asarray(p, dtype=np.float64))[0]
        return X.T.copy(), t.copy()
    return predictor


@dataclass
class IntervalSolution:
    """One merged-grid interval of `continuous_predict`: the stand-in for the `ODESolution` the reference
    returns per interval (/root/reference/src/utilities.jl:11-15).  ``t`` [k+1] are the integrator's own
    step points of the interval (both ends included, like `sol.t`), ``u`` [n_aug, k+1] the augmented state
    ``[x; vec(G); F_triu]`` there (like `sol[:, :]`)."""
    t: np.ndarray
    u: np.ndarray

    def __getitem__(self, key):
        return self.u[key]

    def __len__(self):
        return len(self.t)


def _dense_context(prob: OEDProblem):
    """A second evaluator whose OUTPUT grid is the integrator's step grid: every step of the original
    problem becomes a merged interval of its own (one step each, same step sizes, each variable keeps the
    value of the interval the step came from), so `dynoed_trajectory` returns the state after every step.
    With equal substeps the step size is the same single constant as in the original problem: x and G are
    bit-identical at the grid points, F agrees to rounding (its quadrature is folded once per interval)."""
    if getattr(prob, "_dense", None) is None:
        tp = np.asarray(prob.timegrid.tpoints, dtype=np.float64)
        N = len(tp) - 1
        edges = getattr(prob.alg, "edges", None)
        rel = [np.asarray(e, dtype=np.float64) for e in edges] if edges is not None else \
            [np.arange(prob.alg.substeps + 1, dtype=np.float64) / prob.alg.substeps] * N
        t_dense, owner = [tp[0]], []
        for j in range(N):
            L = tp[j + 1] - tp[j]
            for k in range(1, len(rel[j])):
                t_dense.append(tp[j + 1] if k == len(rel[j]) - 1 else tp[j] + rel[j][k] * L)
                owner.append(j)
        ind = np.asarray(prob.timegrid.indicators - 1)[:, owner]
        libpath, name, info = _resolve_model(prob.system)
        lay = prob.layout
        ctx = runtime.Context(
            name, criterion=prob.objective.id, method=prob.alg.method, tgrid=np.asarray(t_dense),
            indicators=ind, ctrl_offsets=lay.control_offsets, w_offsets=lay.w_offsets,
            ic_offset=lay.ic_offset, tau_offset=lay.tau_offset, n_var=lay.n_var, substeps=1,
            theta=info["theta"], x0=info["x0"], G0=info["G0"], device=prob.device, library_path=libpath)
        counts = [len(r) - 1 for r in rel]
        prob._dense = (ctx, np.asarray(t_dense), np.concatenate([[0], np.cumsum(counts)]))
    return prob._dense


def continuous_predict(prob: OEDProblem, result) -> list:
    """utilities.jl:1-17: the solution on every merged interval, restarted from the end of the previous one
    — here: the state after every integrator step of the GPU evaluation, one `IntervalSolution` per merged
    interval (its first column is the previous interval's last)."""
    ctx, t_dense, first = _dense_context(prob)
    X = ctx.trajectory(np.asarray(getattr(result, "data", result), dtype=np.float64))[0].T   # [n_aug, n_steps+1]
    return [IntervalSolution(t_dense[first[j]:first[j + 1] + 1].copy(), X[:, first[j]:first[j + 1] + 1].copy())
            for j in range(len(first) - 1)]


def extract_fisher(prob: OEDProblem, result) -> np.ndarray:
    """utilities.jl:37-45: F(t_f) as a symmetric np x np matrix."""
    from .criteria import _symmetric_from_vector
    ctx = prob.context()
    X = ctx.trajectory(np.asarray(getattr(result, "data", result), dtype=np.float64))[0]
    return _symmetric_from_vector(X[-1, ctx.info.nz:])


def compute_local_information_gain(Qs):
    """utilities.jl:47-54: P_i(t_k) = Q_i(t_k)^T Q_i(t_k) for a list of n_obs x np matrices Q(t_k)."""
    n_h = np.asarray(Qs[0]).shape[0]
    return [[np.outer(np.asarray(Q)[i], np.asarray(Q)[i]) for Q in Qs] for i in range(n_h)]


def compute_information_gain(prob: OEDProblem, result, dense: bool = False):
    """utilities.jl:56-70: ``(F, Ps, Πs, np, sols)`` with ``Ps[i][k] = Q_i(t_k)^T Q_i(t_k)`` and
    ``Πs[i][k] = F^-1 Ps[i][k] F^-1`` (local / global information gain of observable i).  The time points are
    the merged-grid points, or with ``dense=True`` the integrator's step points (what the reference's saved
    solution points are for a fixed-step run); all of it computed by `dynoed_information_gain` on the GPU."""
    from .criteria import _symmetric_from_vector
    des = np.asarray(getattr(result, "data", result), dtype=np.float64)
    ctx = _dense_context(prob)[0] if dense else prob.context()
    F, P, Pi = ctx.information_gain(des)
    n_obs = ctx.info.n_obs
    Ps = [[P[0, k, i] for k in range(P.shape[1])] for i in range(n_obs)]
    Pis = [[Pi[0, k, i] for k in range(Pi.shape[1])] for i in range(n_obs)]
    sols = continuous_predict(prob, des) if dense else None
    return _symmetric_from_vector(F[0]), Ps, Pis, ctx.info.np, sols


# --- constraints (ModelingToolkit.ConstraintsSystem stand-in) -----------------------------------
@dataclass
class Constraint:
    """lhs - rhs  with bounds  lcons <= expr <= ucons."""
    expr: sp.Expr
    lb: float
    ub: float


def leq(a, b) -> Constraint:
    """a ≲ b  ->  a - b in (-inf, 0]"""
    return Constraint(sp.sympify(a) - sp.sympify(b), -np.inf, 0.0)


def geq(a, b) -> Constraint:
    """a ≳ b  ->  a - b in [0, inf)"""
    return Constraint(sp.sympify(a) - sp.sympify(b), 0.0, np.inf)


@dataclass
class ConstraintsSystem:
    constraints: list
    variables: object = None
    name: str = "constraint_set"


@dataclass
class OptimizationFunction:
    f: object
    grad: object
    cons: object = None
    cons_j: object = None
    syms: list = field(default_factory=list)
    hess: object = None


@dataclass
class OptimizationProblemT:
    f: OptimizationFunction
    u0: np.ndarray
    p: object
    lb: np.ndarray
    ub: np.ndarray
    int: np.ndarray
    lcons: object
    ucons: object
    oed: OEDProblem = None


def OptimizationProblem(prob: OEDProblem, AD=None, u0=None, p=None, *, integer_constraints=False,
                        constraints=None, variable_type=np.float64, **kwargs) -> OptimizationProblemT:
    """problem.jl:96-161.  ``AD`` is accepted for signature compatibility and ignored: the
    gradient comes from the GPU adjoint sweep, not from an AD backend."""
    u0v = get_initial_variables(prob) if u0 is None else u0
    u0a = np.asarray(u0v, dtype=variable_type).copy()
    ctx = prob.context()
    n_var = ctx.n_var

    # Ipopt asks for f(u) and grad f(u) at the same point in separate callbacks: evaluate both in
    # ONE launch (the gradient kernel also returns the objective) and serve the second from a
    # one-entry cache keyed by the design's bytes.
    cache = {"key": None, "c": None, "g": None}

    def _both(u):
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(1, n_var))
        key = u.tobytes()
        if cache["key"] != key:
            c, g, _ = ctx.evaluate(u, grad=True, fim=False)
            cache.update(key=key, c=float(c[0]), g=g[0].copy())
        return cache

    def objective(u, _p=None):
        return _both(u)["c"]

    def gradient(G, u, _p=None):
        g = _both(u)["g"]
        if G is not None:
            G[:] = g
            return G
        return g.copy()

    def hessian(H, u, _p=None):
        """Hessian of the objective: central differences of the exact gradient, 2 n_var + 1 designs
        in one launch (`dynoed_hessian_batch`); also refreshes the f / grad cache at u."""
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(1, n_var))
        c, g, Hm = ctx.hessian(u[0])
        cache.update(key=u.tobytes(), c=float(c), g=g.copy())
        if H is not None:
            H[:] = Hm
            return H
        return Hm

    cons = cons_j = lcons = ucons = None
    if isinstance(constraints, ConstraintsSystem):
        syms = list(states(prob).data)
        exprs = [c.expr for c in constraints.constraints]
        fn = sp.lambdify(syms, exprs, "numpy")
        jn = sp.lambdify(syms, sp.Matrix(exprs).jacobian(syms), "numpy")
        cons = lambda u, _p=None: np.asarray(fn(*u), dtype=np.float64)
        cons_j = lambda u, _p=None: np.asarray(jn(*u), dtype=np.float64)
        lcons = np.array([c.lb for c in constraints.constraints])
        ucons = np.array([c.ub for c in constraints.constraints])

    lb = np.asarray(generate_variable_bounds(prob.system, prob.timegrid, True), dtype=variable_type)
    ub = np.asarray(generate_variable_bounds(prob.system, prob.timegrid, False), dtype=variable_type)
    assert np.all(lb <= u0a) and np.all(u0a <= ub), \
        "The initial variables are not within the bounds. Please check the input!"
    lay = prob.layout
    integrality = np.zeros(n_var, dtype=bool)
    if integer_constraints:
        for o, l in zip(lay.w_offsets, lay.w_lengths):
            integrality[o:o + l] = True                       # measurements are always integer
        for c, o, l in zip(prob.system.controls, lay.control_offsets, lay.control_lengths):
            integrality[o:o + l] = bool(c.meta.get("integer", False))
    syms = [v.name for v in list(prob.system.x_ic) + list(prob.system.controls) + list(prob.system.w)]
    opt_f = OptimizationFunction(objective, gradient, cons, cons_j, syms, hessian)
    return OptimizationProblemT(opt_f, u0a, p, lb, ub, integrality, lcons, ucons, prob)


@dataclass
class OptimizationSolution:
    u: np.ndarray
    objective: float
    success: bool
    message: str
    iterations: int
    evaluations: int


def solve(opt_prob: OptimizationProblemT, method: str = "SLSQP", **options) -> OptimizationSolution:
    """Stand-in for ``solve(optimization_problem, Ipopt.Optimizer())`` (/root/reference/README.md:62):
    hands the GPU objective / gradient closures, bounds and constraints to a scipy NLP solver
    (Ipopt itself is not in this image).  Relaxed problems only; for ``integer_constraints`` use
    the batch evaluator (exhaustive / branch-and-bound over `Context.eval_fixed_controls`)."""
    from scipy.optimize import NonlinearConstraint, minimize
    f = opt_prob.f
    cons = []
    if f.cons is not None:
        cons = [NonlinearConstraint(f.cons, opt_prob.lcons, opt_prob.ucons, jac=f.cons_j)]
    calls = [0]

    def fun(u):
        calls[0] += 1
        return f.f(u)
    opts = dict(ftol=1e-12, maxiter=500) if method == "SLSQP" else {}
    opts.update(options)
    extra = {}
    if method == "trust-constr" and f.hess is not None:   # exact-Hessian mode (Ipopt's default)
        extra["hess"] = lambda u: f.hess(None, u)
    r = minimize(fun, opt_prob.u0, jac=lambda u: f.grad(None, u), method=method,
                 bounds=list(zip(opt_prob.lb, opt_prob.ub)), constraints=cons, options=opts, **extra)
    return OptimizationSolution(np.asarray(r.x), float(r.fun), bool(r.success), str(r.message),
                                int(getattr(r, "nit", 0)), calls[0])
