"""CPU ORACLE (test infrastructure, not product code) — numpy/sympy restatement of the
reference's hot path for ANY model declared with the symbolic front-end.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product path (``dynamicoed.jl_b200``) never does and has no CPU fallback.

What it restates, from the equations (not a translation of the Julia sources):
  * augmentation: residual convention rhs-lhs, Jacobians f_x, f_p, h_x
    (/root/reference/src/augment.jl:113-126), G0 with the unknown-IC block (:182-188),
    sensitivity DAE 0 = f_xdot G' + f_x G + f_p (:231), F' = sum_i w_i (h_x[i,:]G)^T(h_x[i,:]G)
    on the upper triangle in column-major order (:221-226), state order [x; vec(G); F_triu] (:240)
  * sequential integration with a restart at every merged-grid point, piecewise-constant
    controls/weights picked through the indicator matrix, unknown ICs applied at interval 1 only
    (/root/reference/src/discretize.jl:247-265, :305-344)
  * objective  c(Symmetric(F + tau I)) + tau  (/root/reference/src/problem.jl:115-120,
    /root/reference/src/fisher.jl:3-15) for the six criteria (fisher.jl:26-107)
  * derivatives the way the reference obtains them — forward-mode AD pushed through the
    integrator (AutoForwardDiff, /root/reference/src/problem.jl:154): `Jet` carries one tangent
    per design variable through every arithmetic operation of the fixed-step integrator.

PARITY PIN STATUS: the integrator here is a FIXED-step explicit RK (classic RK4 or the Tsit5
tableau) or a fixed-grid Rosenbrock (ROS2) method, because the 1e-10 parity target is only
meaningful on a fixed tableau and grid.  The criteria/packing/grid parts are
pinned by the reference's tight goldens (test/criteria.jl, test/discretize.jl,
test/mtk_extensions.jl); the integration results are pinned only by the reference's loose optima
(atol 1e-2) and by independent tight-tolerance scipy integration => "integration: parity
unpinned beyond 1e-2 against the reference itself" (Julia is not installed here).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import sympy as sp

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import dynamicoed.jl_b200.symbolic as S            # noqa: E402  (metadata vocabulary only)
from dynamicoed.jl_b200.discretize import Timegrid, design_layout  # noqa: E402  (golden-pinned)

FISHER_A, FISHER_D, FISHER_E, CRIT_A, CRIT_D, CRIT_E = range(6)
LOG_D = 6      # -log det(F + tau I): extension (BASELINE north_star), not one of the reference's six
CRITERIA = {"FisherACriterion": 0, "FisherDCriterion": 1, "FisherECriterion": 2,
            "ACriterion": 3, "DCriterion": 4, "ECriterion": 5}


# ----------------------------------------------------------------------------------------------
# forward-mode AD number (batched): value [B], tangents [B, nd]
# ----------------------------------------------------------------------------------------------
class Jet:
    __slots__ = ("v", "d")
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @staticmethod
    def lift(x, like):
        if isinstance(x, Jet):
            return x
        return Jet(np.broadcast_to(np.asarray(x, dtype=np.float64), like.v.shape).copy(),
                   np.zeros_like(like.d))

    def __add__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v + o.v, self.d + o.d)
        return Jet(self.v + o, self.d)
    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v - o.v, self.d - o.d)
        return Jet(self.v - o, self.d)

    def __rsub__(self, o):
        return Jet(o - self.v, -self.d)

    def __neg__(self):
        return Jet(-self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v * o.v, self.d * o.v[:, None] + o.d * self.v[:, None])
        return Jet(self.v * o, self.d * o)
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Jet):
            q = self.v / o.v
            return Jet(q, (self.d - o.d * q[:, None]) / o.v[:, None])
        return Jet(self.v / o, self.d / o)

    def __rtruediv__(self, o):
        q = o / self.v
        return Jet(q, -self.d * (q / self.v)[:, None])

    def __pow__(self, e):
        if isinstance(e, Jet):
            return jexp(e * jlog(self))
        if e == 2:
            return self * self
        if isinstance(e, (int, np.integer)) and e >= 0:
            out = 1.0
            for _ in range(int(e)):
                out = self * out
            return out
        val = self.v ** e
        return Jet(val, self.d * (e * self.v ** (e - 1))[:, None])


def jexp(x):
    if isinstance(x, Jet):
        e = np.exp(x.v)
        return Jet(e, x.d * e[:, None])
    return np.exp(x)


def jlog(x):
    if isinstance(x, Jet):
        return Jet(np.log(x.v), x.d / x.v[:, None])
    return np.log(x)


def jsqrt(x):
    if isinstance(x, Jet):
        s = np.sqrt(x.v)
        return Jet(s, x.d / (2.0 * s)[:, None])
    return np.sqrt(x)


def jsin(x):
    if isinstance(x, Jet):
        return Jet(np.sin(x.v), x.d * np.cos(x.v)[:, None])
    return np.sin(x)


def jcos(x):
    if isinstance(x, Jet):
        return Jet(np.cos(x.v), -x.d * np.sin(x.v)[:, None])
    return np.cos(x)


_JETMOD = {"exp": jexp, "log": jlog, "sqrt": jsqrt, "sin": jsin, "cos": jcos}


# ----------------------------------------------------------------------------------------------
# integrator tableaus (fixed step)
# ----------------------------------------------------------------------------------------------
def tableau(name: str):
    """Explicit RK tableaus as (A, b, c) with exact rational/decimal coefficients as doubles."""
    if name == "rk4":
        A = [[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]]
        b = [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0]
        c = [0.0, 0.5, 0.5, 1.0]
        return A, b, c
    if name == "tsit5":
        # Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the first column
        # simplifying assumption" — the method OrdinaryDiffEq.Tsit5 (the reference's default alg,
        # /root/reference/src/problem.jl:25) steps with; used here with FIXED steps (b7 = 0, so
        # the FSAL stage does not enter the solution).
        c = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0]
        A = [[],
             [0.161],
             [-0.008480655492356989, 0.335480655492357],
             [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
             [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
             [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
              -0.028269050394068383]]
        b = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
             -3.290069515436081, 2.324710524099774]
        return A, b, c
    raise ValueError(name)


# ----------------------------------------------------------------------------------------------
# criteria (fisher.jl) on plain matrices + their matrix gradients
# ----------------------------------------------------------------------------------------------
def symmetric_from_vector(x):
    """fisher.jl:3-10 — packed upper triangle (column-major, x[j(j-1)/2+i]) -> symmetric matrix."""
    x = np.asarray(x)
    n = int(math.sqrt(2 * x.shape[-1] + 0.25) - 0.5)
    M = np.zeros(x.shape[:-1] + (n, n), dtype=x.dtype)
    k = 0
    for j in range(n):
        for i in range(j + 1):
            M[..., i, j] = x[..., k]
            M[..., j, i] = x[..., k]
            k += 1
    return M


def vector_from_symmetric(M):
    n = M.shape[-1]
    return np.stack([M[..., i, j] for j in range(n) for i in range(j + 1)], axis=-1)


def criterion(crit: int, F, tau=0.0):
    """c(F, tau) = c(Symmetric(F + tau I)) (fisher.jl:12-15, :26-107). F: [..., n, n]."""
    F = np.asarray(F, dtype=np.float64)
    n = F.shape[-1]
    M = F + np.asarray(tau, dtype=np.float64)[..., None, None] * np.eye(n)
    M = np.triu(M) + np.swapaxes(np.triu(M, 1), -1, -2)      # Symmetric(): upper triangle rules
    if crit == FISHER_A:
        return -np.trace(M, axis1=-2, axis2=-1)
    if crit == FISHER_D:
        return -np.linalg.det(M)
    if crit == FISHER_E:
        return -np.min(np.linalg.eigvalsh(M), axis=-1)
    if crit == CRIT_A:
        return np.sum(1.0 / np.linalg.eigvalsh(M), axis=-1)
    if crit == CRIT_D:
        return 1.0 / np.linalg.det(M)
    if crit == CRIT_E:
        return np.max(1.0 / np.linalg.eigvalsh(M), axis=-1)
    if crit == LOG_D:
        return -np.linalg.slogdet(M)[1]
    raise ValueError(crit)


def criterion_matrix_gradient(crit: int, F, tau=0.0):
    """d c(M)/dM for symmetric M = F + tau I (checked against finite
    differences in tests/test_oracle.py)."""
    F = np.asarray(F, dtype=np.float64)
    n = F.shape[-1]
    M = F + np.asarray(tau, dtype=np.float64)[..., None, None] * np.eye(n)
    if crit == FISHER_A:
        return -np.broadcast_to(np.eye(n), M.shape).copy()
    if crit == FISHER_D:
        return -np.linalg.det(M)[..., None, None] * np.linalg.inv(M)
    if crit == CRIT_A:
        Mi = np.linalg.inv(M)
        return -Mi @ Mi
    if crit == CRIT_D:
        return -np.linalg.inv(M) / np.linalg.det(M)[..., None, None]
    if crit == LOG_D:
        return -np.linalg.inv(M)
    lam, V = np.linalg.eigh(M)
    v = V[..., :, 0]
    vv = v[..., :, None] * v[..., None, :]
    if crit == FISHER_E:
        return -vv
    if crit == CRIT_E:
        return -vv / (lam[..., 0] ** 2)[..., None, None]
    raise ValueError(crit)


# ----------------------------------------------------------------------------------------------
# the oracle
# ----------------------------------------------------------------------------------------------
class Oracle:
    """Batched CPU evaluation of objective, F and (forward-mode) gradient for one model.

    sys       : un-augmented ``ODESystem`` (the oracle augments it itself, numerically)
    criterion : 0..5 (FisherA, FisherD, FisherE, A, D, E)
    method    : "rk4" | "tsit5" (explicit, fixed step) | "ros2" | "ros3p" (fixed-grid Rosenbrock,
                values only in this oracle)
    substeps  : int (uniform per merged interval) or list of per-interval substep *edge arrays*
                (relative positions in [0,1], first 0 and last 1) for graded grids
    theta     : optional dict {parameter name: value} overriding defaults
    """

    def __init__(self, sys, criterion=FISHER_A, method="rk4", substeps=4, tspan=None, theta=None):
        self.sys = sys
        self.crit = int(criterion)
        self.method = method
        self._classify()
        self.tspan = tuple(tspan) if tspan is not None else sys.tspan
        self.grid = Timegrid.build(self.tspan, *self.controls, *self.w)
        self.layout = design_layout(self, self.grid)
        self.N = len(self.grid.timespans)
        self.n_var = self.layout.n_var
        if isinstance(substeps, (int, np.integer)):
            m = int(substeps)
            self.edges = [np.arange(m + 1, dtype=np.float64) / m for _ in range(self.N)]
            self.uniform = m
        else:
            self.edges = [np.asarray(e, dtype=np.float64) for e in substeps]
            self.uniform = None
        self.theta = {p.name: float(S.getdefault(p)) for p in self.fixed_and_tunable}
        if theta:
            self.theta.update(theta)
        self._build_numeric()

    # -- augment.jl:134-188 -------------------------------------------------------------------
    def _classify(self):
        sys = self.sys
        p = sys.parameters()
        self.controls = [q for q in p if S.isinput(q)]
        tun = [q for q in p if S.istunable(q) and q not in self.controls]
        self.x_all = sys.get_states()
        self.x_ic, self.ic_state = [], []
        for i, xi in enumerate(self.x_all):
            if S.istunable(xi):
                self.ic_state.append(i)
                ic = S.Num(f"{xi.name}₀", "parameter", default=S.getdefault(xi),
                                bounds=S.getbounds(xi), tunable=True, variable_ic=i + 1)
                self.x_ic.append(ic)
        self.p_tunable = tun + self.x_ic
        self.np_ = len(self.p_tunable)
        self.n_ord = len(tun)
        self.fixed_and_tunable = [q for q in p if q not in self.controls]
        self.w = []
        for i, e in enumerate(sys.observed, start=1):
            assert S.is_measured(e.lhs)
            name = "w" + str(i).translate(str.maketrans("0123456789", "₀₁₂₃₄₅₆₇₈₉"))
            self.w.append(S.Num(name, "parameter", default=0.0, bounds=(0, 1),
                                     measurement_rate=S.get_measurement_rate(e.lhs)))
        self.n_obs = len(self.w)
        nx = len(self.x_all)
        G0 = np.zeros((nx, self.np_))
        for a, r in enumerate(self.ic_state):          # augment.jl:186, literal (quirk Q3)
            for b, r2 in enumerate(self.ic_state):
                G0[r, self.np_ - r2 - 1] = 1.0 if a == b else 0.0
        self.G0_all = G0

    def _build_numeric(self):
        sys = self.sys
        x = self.x_all
        dsyms = [sys.derivative_of(xi) for xi in x]
        res = [e.residual() for e in sys.eqs]
        diff_rows = [r for r, e in enumerate(res) if any(e.has(d) for d in dsyms)]
        alg_rows = [r for r in range(len(res)) if r not in diff_rows]
        d_idx = [i for i, d in enumerate(dsyms) if any(e.has(d) for e in res)]
        a_idx = [i for i in range(len(x)) if i not in d_idx]
        assert len(alg_rows) == len(a_idx)
        self.d_idx, self.a_idx = d_idx, a_idx
        xd = [x[i] for i in d_idx]
        xa = [x[i] for i in a_idx]
        # explicit rows must read D(x_i) ~ rhs_i : f_xdot = -I on them
        fdot = sp.Matrix([res[r] for r in diff_rows]).jacobian([dsyms[i] for i in d_idx])
        assert fdot == -sp.eye(len(d_idx)), "oracle supports rows of the form D(x) ~ rhs only"
        f = sp.Matrix([res[r] + dsyms[d_idx[k]] for k, r in enumerate(diff_rows)])
        phi = {}
        if a_idx:
            sol = sp.solve([res[r] for r in alg_rows], xa, dict=True)
            assert len(sol) == 1
            phi = sol[0]
        tun_ord = self.p_tunable[:self.n_ord]
        # partial Jacobians of the *un-torn* residuals, evaluated with x_a = phi(x_d, p)
        fx_d = f.jacobian(xd)
        fx_a = f.jacobian(xa) if xa else sp.zeros(len(xd), 0)
        fp = f.jacobian(tun_ord) if tun_ord else sp.zeros(len(xd), 0)
        phi_v = sp.Matrix([phi[v] for v in xa]) if xa else sp.zeros(0, 1)
        phi_x = phi_v.jacobian(xd) if xa else sp.zeros(0, len(xd))
        phi_p = phi_v.jacobian(tun_ord) if (xa and tun_ord) else sp.zeros(len(xa), len(tun_ord))
        h = sp.Matrix([e.rhs for e in sys.observed])
        hx_d = h.jacobian(xd)
        hx_a = h.jacobian(xa) if xa else sp.zeros(len(h), 0)
        args = [sys.iv] + xd + xa + list(self.fixed_and_tunable) + list(self.controls)

        def lam(M):
            flat = [sp.sympify(e) for e in M]
            fn = sp.lambdify(args, flat, modules=[_JETMOD, "numpy"], cse=False)
            shape = M.shape
            return lambda *a: _reshape(fn(*a), shape)
        self._f, self._fxd, self._fxa, self._fp = lam(f), lam(fx_d), lam(fx_a), lam(fp)
        self._phi, self._phix, self._phip = lam(phi_v), lam(phi_x), lam(phi_p)
        self._hxd, self._hxa = lam(hx_d), lam(hx_a)
        self.nxd, self.nxa = len(xd), len(xa)
        self.nF = self.np_ * (self.np_ + 1) // 2
        self.n_aug = self.nxd + self.nxd * self.np_ + self.nF
        self.x0 = np.array([float(S.getdefault(v)) for v in xd])
        self.G0 = self.G0_all[d_idx, :]

    # -- the augmented right-hand side on [x_d; vec(G_d); F_triu] ------------------------------
    def rhs(self, t, z, u, w):
        nxd, np_ = self.nxd, self.np_
        xd = z[:nxd]
        G = [[z[nxd + j * nxd + i] for j in range(np_)] for i in range(nxd)]
        th = [self.theta[p.name] for p in self.fixed_and_tunable]
        pre = [t] + list(xd)
        xa = [r[0] for r in self._phi(*pre, *([0.0] * self.nxa), *th, *u)] if self.nxa else []
        a = pre + list(xa) + th + list(u)
        f, fxd, fxa, fp = self._f(*a), self._fxd(*a), self._fxa(*a), self._fp(*a)
        hxd, hxa = self._hxd(*a), self._hxa(*a)
        n_ord = self.n_ord
        # sensitivities of torn states: G_a = phi_x G_d + phi_p   (0 = f_x G + f_p on alg. rows)
        Ga = []
        if self.nxa:
            phix, phip = self._phix(*a), self._phip(*a)
            for r in range(self.nxa):
                row = []
                for j in range(np_):
                    s = 0.0
                    for k in range(nxd):
                        s = s + phix[r][k] * G[k][j]
                    if j < n_ord:
                        s = s + phip[r][j]
                    row.append(s)
                Ga.append(row)
        dz = [None] * len(z)
        for i in range(nxd):
            dz[i] = f[i][0]
        for j in range(np_):
            for i in range(nxd):
                s = 0.0
                for k in range(nxd):
                    s = s + fxd[i][k] * G[k][j]
                for k in range(self.nxa):
                    s = s + fxa[i][k] * Ga[k][j]
                if j < n_ord:
                    s = s + fp[i][j]
                dz[nxd + j * nxd + i] = s
        # Q = h_x G (all states incl. torn ones), F' = sum_i w_i Q_i^T Q_i on triu, column-major
        Q = []
        for o in range(self.n_obs):
            row = []
            for j in range(np_):
                s = 0.0
                for k in range(nxd):
                    s = s + hxd[o][k] * G[k][j]
                for k in range(self.nxa):
                    s = s + hxa[o][k] * Ga[k][j]
                row.append(s)
            Q.append(row)
        k = nxd + nxd * np_
        for j in range(np_):
            for i in range(j + 1):
                s = 0.0
                for o in range(self.n_obs):
                    s = s + w[o] * (Q[o][i] * Q[o][j])
                dz[k] = s
                k += 1
        return dz, Q

    # -- one fixed step --------------------------------------------------------------------------
    def _step_rk(self, t, z, h, u, w, tab):
        A, b, c = tab
        ks = []
        for s in range(len(b)):
            zs = z
            if s > 0:
                zs = list(z)
                for j, a in enumerate(A[s]):
                    if a != 0.0:
                        zs = [zi + (h * a) * kj for zi, kj in zip(zs, ks[j])]
            k, _ = self.rhs(t + c[s] * h, zs, u, w)
            ks.append(k)
        out = list(z)
        for s, bs in enumerate(b):
            out = [zi + (h * bs) * ki for zi, ki in zip(out, ks[s])]
        return out

    def _step_ros2(self, t, z, h, u, w):
        from . import rosenbrock_oracle
        if any(isinstance(v, Jet) for v in z):
            raise NotImplementedError("the numpy oracle's Rosenbrock step carries values only; "
                                      "gradients of the stiff configuration come from the C++ oracle")
        return rosenbrock_oracle.step(self, t, z, h, u, w)

    # -- sequential solve (discretize.jl:305-344) ------------------------------------------------
    def integrate(self, designs, with_tangents=False, save=False):
        designs = np.atleast_2d(np.asarray(designs, dtype=np.float64))
        B, nv = designs.shape
        assert nv == self.n_var
        lay = self.layout
        if with_tangents:
            eye = np.eye(nv)
            col = [Jet(designs[:, k].copy(), np.broadcast_to(eye[k], (B, nv)).copy())
                   for k in range(nv)]
            zero = Jet(np.zeros(B), np.zeros((B, nv)))
            const = lambda v: Jet(np.full(B, float(v)), np.zeros((B, nv)))
        else:
            col = [designs[:, k].copy() for k in range(nv)]
            const = lambda v: np.full(B, float(v))
        z = [const(v) for v in self.x0] + \
            [const(self.G0[i, j]) for j in range(self.np_) for i in range(self.nxd)] + \
            [const(0.0) for _ in range(self.nF)]
        # unknown ICs overwrite u0 at interval 1 (discretize.jl:260-263)
        for a, st in enumerate(self.ic_state):
            z[self.d_idx.index(st)] = col[lay.ic_offset + a]
        tab = tableau(self.method) if self.method in ("rk4", "tsit5") else None
        nc = len(self.controls)
        traj = [list(z)] if save else None
        hbar = None
        if self.uniform is not None:        # single dt on a grid that is uniform up to rounding
            from .cpp_oracle import uniform_step
            hbar = uniform_step(self.grid.tpoints, self.uniform)
        for j, (t0, t1) in enumerate(self.grid.timespans):
            u = [col[lay.control_offsets[c] + int(self.grid.indicators[c, j]) - 1]
                 for c in range(nc)]
            w = [col[lay.w_offsets[o] + int(self.grid.indicators[nc + o, j]) - 1]
                 for o in range(self.n_obs)]
            e = self.edges[j]
            L = t1 - t0
            for s in range(len(e) - 1):
                if self.uniform is not None:
                    h = hbar if hbar is not None else L / self.uniform
                    ts = t0 + s * h
                else:
                    ts = t0 + e[s] * L
                    h = (t0 + e[s + 1] * L) - ts
                if tab is not None:
                    z = self._step_rk(ts, z, h, u, w, tab)
                else:
                    z = self._step_ros2(ts, z, h, u, w)
            if save:
                traj.append(list(z))
        return (z, traj, col) if save else (z, None, col)

    # -- objective (problem.jl:115-120) and its forward-mode gradient ---------------------------
    def evaluate(self, designs, grad=False):
        designs = np.atleast_2d(np.asarray(designs, dtype=np.float64))
        z, _, col = self.integrate(designs, with_tangents=grad)
        k0 = self.nxd + self.nxd * self.np_
        Fz = z[k0:k0 + self.nF]
        tau = designs[:, self.layout.tau_offset]
        if not grad:
            Fv = np.stack(Fz, axis=-1)
            obj = criterion(self.crit, symmetric_from_vector(Fv), tau) + tau
            return obj, None, Fv
        Fv = np.stack([f.v for f in Fz], axis=-1)                   # [B, nF]
        dF = np.stack([f.d for f in Fz], axis=1)                    # [B, nF, nv]
        Fm = symmetric_from_vector(Fv)
        obj = criterion(self.crit, Fm, tau) + tau
        Lam = criterion_matrix_gradient(self.crit, Fm, tau)         # [B, n, n]
        dM = symmetric_from_vector(np.swapaxes(dF, 1, 2))           # [B, nv, n, n]
        dM[:, self.layout.tau_offset] += np.eye(self.np_)
        g = np.einsum("bij,bvij->bv", Lam, dM)
        g[:, self.layout.tau_offset] += 1.0
        return obj, g, Fv

    def trajectory(self, design):
        """X [n_aug, N+1] at the merged-grid points (what the predictor of problem.jl:80-90
        returns with saveat = interval ends)."""
        z, traj, _ = self.integrate(design, save=True)
        return np.array([[float(np.asarray(v).reshape(-1)[0]) for v in zz] for zz in traj]).T


def _reshape(flat, shape):
    r, c = shape
    return [[flat[i * c + j] for j in range(c)] for i in range(r)]
