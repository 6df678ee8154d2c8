"""CPU ORACLE (test infrastructure, not product code) — the reference's ACTUAL integration:
ADAPTIVE Tsit5 as OrdinaryDiffEq.jl (compat "6", /root/reference/Project.toml:18-29) runs it when
DynamicOED calls ``solve(prob_i, Tsit5(); saveat = (t_i, t_{i+1}))`` once per merged interval
(/root/reference/src/discretize.jl:305-344; default alg /root/reference/src/problem.jl:25; the
reference drops ``diffeq_options``, so every solve runs with the package defaults).

OrdinaryDiffEq is an un-vendored dependency (no Manifest.toml is committed); this file restates its
published algorithm from the package's documented defaults:
  * tolerances abstol = 1e-6, reltol = 1e-3; error norm = RMS of  err_i / (abstol + max(|u_prev,i|, |u_i|) reltol)
    over ALL states of the simplified augmented system [x; vec(G); F_triu] (F is a state there)
  * Tsit5 tableau with the embedded error weights btilde (7 stages, the 7th is f(u_new))
  * PI step-size controller: beta1 = 7/(10 k), beta2 = 2/(5 k) with k = 5 (7/50, 2/25), gamma = 9/10,
    qmin = 1/5, qmax = 10, qsteady = [1, 1], qoldinit = 1e-4; accept iff EEst <= 1;
    rejected: dt /= min(1/qmin, EEst^beta1 / gamma)
  * initial step by Hairer-Wanner's algorithm as OrdinaryDiffEq implements it (order divisor 5,
    smalldt 1e-6, capped by 100 dt0 and dtmax = interval length)
  * a fresh solve (fresh controller state, fresh initial step) for every merged interval, last
    step clipped to the interval end.

What it pins (tests/test_oracle.py): at the reference optimum w* = [0,0,0,1,1,1,1,0,0,0] of
/root/reference/test/references/1D.jl:57-64 the reference prints the objective
-1.2823515846720533e-2; the exact integral is -1.28137e-2, i.e. the printed digits carry the
adaptive integrator's own error (7.6e-4 relative) and are reproduced only by the same step
sequence.  Only tests import this module.
"""
from __future__ import annotations

import math

import numpy as np

# Tsitouras 5(4): a, b as in oed_oracle.tableau("tsit5"); c7 = 1, a7j = b_j (FSAL stage)
_C = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0]
_A = [[],
      [0.161],
      [-0.008480655492356989, 0.335480655492357],
      [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
      [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
      [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
       -0.028269050394068383]]
_B = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
      2.324710524099774]
# error weights  utilde = dt * sum_i btilde_i k_i  (b - bhat of the embedded 4th-order solution)
_BT = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
       0.5823571654525552, -0.45808210592918697, 0.015151515151515152]

ABSTOL, RELTOL = 1e-6, 1e-3
ORDER = 5
BETA2 = 2.0 / (5 * ORDER)
BETA1 = 7.0 / (10 * ORDER)
GAMMA, QMIN, QMAX, QOLDINIT = 0.9, 0.2, 10.0, 1e-4


def _rms(v):
    v = np.asarray(v, dtype=np.float64)
    return math.sqrt(float(np.sum(v * v)) / v.size)


def initial_dt(f, u0, t, dtmax):
    """OrdinaryDiffEq's ode_determine_initdt (forward time)."""
    smalldt = 1e-6
    sk = ABSTOL + np.abs(u0) * RELTOL
    f0 = f(t, u0)
    d0 = _rms(u0 / sk)
    d1 = _rms(f0 / sk)
    dt0 = smalldt if (d0 < 1e-5 or d1 < 1e-5) else (d0 / d1) / 100.0
    dt0 = min(dt0, dtmax)
    u1 = u0 + dt0 * f0
    f1 = f(t + dt0, u1)
    d2 = _rms((f1 - f0) / sk) / dt0
    m = max(d1, d2)
    dt1 = max(smalldt, dt0 * 1e-3) if m <= 1e-15 else 10.0 ** (-(2.0 + math.log10(m)) / ORDER)
    return min(100.0 * dt0, dt1, dtmax)


def solve_interval(f, u0, t0, t1, stats=None):
    """One `solve(prob, Tsit5(); saveat = (t0, t1))`: returns u(t1)."""
    u = np.array(u0, dtype=np.float64)
    t = t0
    dtmax = t1 - t0
    dt = initial_dt(f, u, t, dtmax)
    qold = QOLDINIT
    while t < t1:
        dt = min(dt, t1 - t)                       # modify_dt_for_tstops!
        k = [f(t, u)]
        for s in range(1, 6):
            us = u.copy()
            for j, a in enumerate(_A[s]):
                us = us + (dt * a) * k[j]
            k.append(f(t + _C[s] * dt, us))
        unew = u.copy()
        for s in range(6):
            unew = unew + (dt * _B[s]) * k[s]
        k.append(f(t + dt, unew))
        utilde = np.zeros_like(u)
        for s in range(7):
            utilde = utilde + (dt * _BT[s]) * k[s]
        EEst = _rms(utilde / (ABSTOL + np.maximum(np.abs(u), np.abs(unew)) * RELTOL))
        # stepsize_controller!(::PIController)
        if EEst == 0.0:
            q = 1.0 / QMAX
            q11 = 0.0
        else:
            q11 = EEst ** BETA1
            q = q11 / (qold ** BETA2)
            q = max(1.0 / QMAX, min(1.0 / QMIN, q / GAMMA))
        if stats is not None:
            stats.append((t, dt, EEst, EEst <= 1.0))
        if EEst <= 1.0:                            # accept
            qold = max(EEst, QOLDINIT)
            tn = t + dt
            t = t1 if abs(tn - t1) < 100.0 * np.finfo(float).eps * max(abs(tn), abs(t1), 1.0) else tn
            u = unew
            dt = min(dt / q, dtmax)
        else:                                      # step_reject_controller!
            dt = dt / min(1.0 / QMIN, q11 / GAMMA)
    return u


def integrate(oracle, design, stats=None):
    """The reference's sequential_solve for ONE design with the adaptive integrator: final
    [x; vec(G); F_triu] of the oracle's model (`oracle` is an oed_oracle.Oracle; its rhs is used as is)."""
    design = np.asarray(design, dtype=np.float64).reshape(-1)
    lay, grid = oracle.layout, oracle.grid
    nc = len(oracle.controls)
    z = np.concatenate([oracle.x0, oracle.G0.T.reshape(-1), np.zeros(oracle.nF)])
    for a, st in enumerate(oracle.ic_state):
        z[oracle.d_idx.index(st)] = design[lay.ic_offset + a]
    for j, (t0, t1) in enumerate(grid.timespans):
        u = [np.array([design[lay.control_offsets[c] + int(grid.indicators[c, j]) - 1]]) for c in range(nc)]
        w = [np.array([design[lay.w_offsets[o] + int(grid.indicators[nc + o, j]) - 1]]) for o in range(oracle.n_obs)]

        def f(t, y):
            dz, _ = oracle.rhs(t, [np.array([v]) for v in y], u, w)
            return np.array([float(np.asarray(v).reshape(-1)[0]) for v in dz])
        z = solve_interval(f, z, float(t0), float(t1), stats)
    return z


def objective(oracle, design, stats=None):
    """c(F(t_f) + tau I) + tau with the adaptive integration (problem.jl:115-120)."""
    from .oed_oracle import criterion, symmetric_from_vector
    z = integrate(oracle, design, stats)
    k0 = oracle.nxd + oracle.nxd * oracle.np_
    F = z[k0:k0 + oracle.nF]
    tau = float(np.asarray(design).reshape(-1)[oracle.layout.tau_offset])
    return float(criterion(oracle.crit, symmetric_from_vector(F[None, :]), np.array([tau]))[0] + tau), F
