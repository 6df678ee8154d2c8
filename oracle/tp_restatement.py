"""CPU RESTATEMENT (test infrastructure, not product code) of the ALGORITHM of the time-parallel
small-batch kernel, `tp_eval_kernel` (dynamicoed.jl_b200/csrc/dynoed_rosenbrock.cuh), on
top of the numpy oracle's own step function — so that the no-GPU test suite can check the three claims the
kernel rests on against the reference-style gradient (forward-mode duals through the whole integration,
/root/reference/src/problem.jl:154, restated in oracle/oed_oracle.py):

  1. given the states x_k, one integrator step maps the sensitivities G affinely, G_{k+1} = psi_k + Phi_k G_k
     (G' = f_x G + f_p, /root/reference/src/augment.jl:231), and Phi_k, psi_k come out of ONE dual-number
     step taken from (x_k, G = 0);
  2. the chain of these affine maps may be evaluated in chunks (compose per chunk, chain the chunk maps,
     replay each chunk from its boundary value);
  3. the gradient with respect to control values and unknown initial conditions is the discrete adjoint
     lambda_k = s_k + J_k^T lambda_{k+1} over per-step Jacobians J_k = d z_{k+1} / d z_k taken step by step
     (each step differentiated on its own, seeded with the identity), chunked the same way; d/dw follows from
     the interval quadratures, d/dtau = tr(Lambda) + 1.

Only tests/ may import this module.  Explicit tableaus only (the numpy oracle's Rosenbrock step carries
values only); one design at a time; steps are batched along the oracle's batch axis."""
from __future__ import annotations

import numpy as np

from .oed_oracle import (Jet, Oracle, criterion, criterion_matrix_gradient, symmetric_from_vector, tableau)


def _step_table(orc: Oracle):
    """(t, h, interval) of every integrator step, exactly as Oracle.integrate walks them."""
    hbar = None
    if orc.uniform is not None:
        from .cpp_oracle import uniform_step
        hbar = uniform_step(orc.grid.tpoints, orc.uniform)
    ts, hs, iv = [], [], []
    for j, (t0, t1) in enumerate(orc.grid.timespans):
        e, L = orc.edges[j], t1 - t0
        for s in range(len(e) - 1):
            if orc.uniform is not None:
                h = hbar if hbar is not None else L / orc.uniform
                t = t0 + s * h
            else:
                t = t0 + e[s] * L
                h = (t0 + e[s + 1] * L) - t
            ts.append(t); hs.append(h); iv.append(j)
    return np.array(ts), np.array(hs), np.array(iv)


def _chunks(n_steps: int, n_chunks: int):
    n_chunks = min(n_chunks, n_steps)
    return [(c * n_steps // n_chunks, (c + 1) * n_steps // n_chunks) for c in range(n_chunks)]


def evaluate(orc: Oracle, design, n_chunks: int = 16):
    """(objective, gradient [n_var], F [nF], diagnostics) of ONE design by the kernel's algorithm."""
    assert orc.method in ("rk4", "tsit5")
    design = np.asarray(design, dtype=np.float64).reshape(-1)
    lay, nx, npar, nF = orc.layout, orc.nxd, orc.np_, orc.nF
    NG, NZ = nx * npar, nx + nx * npar
    nc, nobs = len(orc.controls), orc.n_obs
    ND = NZ + nc
    tab = tableau(orc.method)
    t, h, iv = _step_table(orc)
    K = len(t)
    ctrl_idx = [[int(orc.grid.indicators[c, j]) - 1 for j in iv] for c in range(nc)]
    w_idx = [[int(orc.grid.indicators[nc + o, j]) - 1 for j in iv] for o in range(nobs)]
    u_k = [design[[lay.control_offsets[c] + i for i in ctrl_idx[c]]] for c in range(nc)]       # [K] each
    w_k = [design[[lay.w_offsets[o] + i for i in w_idx[o]]] for o in range(nobs)]

    # ---- phase 1: sequential sweep; only the STATES of it are used below ----
    z = [np.array([float(v)]) for v in orc.x0] + \
        [np.array([float(orc.G0[i, j])]) for j in range(npar) for i in range(nx)] + [np.zeros(1) for _ in range(nF)]
    for a, st in enumerate(orc.ic_state):
        z[orc.d_idx.index(st)] = np.array([design[lay.ic_offset + a]])
    xs, Gs_seq = np.empty((K, nx)), np.empty((K, NG))
    for k in range(K):
        xs[k] = [float(z[i][0]) for i in range(nx)]
        Gs_seq[k] = [float(z[nx + i][0]) for i in range(NG)]
        z = orc._step_rk(t[k], z, h[k], [np.array([u[k]]) for u in u_k], [np.array([w[k]]) for w in w_k], tab)

    # ---- phase 1b: G by the chain of affine maps; one dual-number step per step from (x_k, G = 0) ----
    nd = NG
    zero_d = np.zeros((K, nd))
    zj = [Jet(xs[:, i].copy(), zero_d.copy()) for i in range(nx)]
    for q in range(NG):
        d = zero_d.copy(); d[:, q] = 1.0
        zj.append(Jet(np.zeros(K), d))
    zj += [Jet(np.zeros(K), zero_d.copy()) for _ in range(nF)]
    # per-step (t, h) along the batch axis: a plain scalar when the grid has one step size, else dual numbers
    # with zero tangents (the oracle's Jet multiplies by Python scalars or by other Jets)
    h_b = (lambda like: float(h[0])) if np.all(h == h[0]) else (lambda like: Jet(h.copy(), np.zeros_like(like)))
    t_b = lambda like: Jet(t.copy(), np.zeros_like(like))
    out = orc._step_rk(t_b(zero_d), zj, h_b(zero_d), [Jet(u.copy(), zero_d.copy()) for u in u_k],
                       [Jet(w.copy(), zero_d.copy()) for w in w_k], tab)
    psi = np.stack([out[nx + a].v for a in range(NG)], axis=1)              # [K, NG]
    Phi = np.stack([out[nx + a].d for a in range(NG)], axis=1)              # [K, NG(out), NG(in)]
    G0 = np.array([float(orc.G0[i, j]) for j in range(npar) for i in range(nx)])
    chunks = _chunks(K, n_chunks)
    comp = []
    for lo, hi in chunks:                                                   # compose: G_hi = b + A G_lo
        A, b = np.eye(NG), np.zeros(NG)
        for k in range(lo, hi):
            A, b = Phi[k] @ A, psi[k] + Phi[k] @ b
        comp.append((A, b))
    bound = [G0]
    for A, b in comp:                                                       # chain the chunk maps
        bound.append(b + A @ bound[-1])
    Gs = np.empty((K, NG))
    for (lo, hi), g0 in zip(chunks, bound):                                 # replay
        g = g0
        for k in range(lo, hi):
            Gs[k] = g
            g = psi[k] + Phi[k] @ g
    diag = {"G_chain_vs_sweep": float(np.max(np.abs(Gs - Gs_seq)) / max(np.max(np.abs(Gs_seq)), 1e-300))}

    # ---- phase 2: every step on its own, seeded with the identity in (z_k, u) ----
    zero_d = np.zeros((K, ND))
    zj = []
    for a in range(NZ):
        d = zero_d.copy(); d[:, a] = 1.0
        zj.append(Jet((xs[:, a] if a < nx else Gs[:, a - nx]).copy(), d))
    zj += [Jet(np.zeros(K), zero_d.copy()) for _ in range(nF)]
    uj = []
    for c in range(nc):
        d = zero_d.copy(); d[:, NZ + c] = 1.0
        uj.append(Jet(u_k[c].copy(), d))
    # unit weights, one observable at a time: the UNWEIGHTED quadrature increments of the step (what the
    # kernel integrates; F is linear in w)
    dP_v, dP_d = np.empty((nobs, K, nF)), np.empty((nobs, K, nF, ND))
    J = None
    for o in range(nobs):
        wj = [Jet(np.full(K, 1.0 if oo == o else 0.0), zero_d.copy()) for oo in range(nobs)]
        out = orc._step_rk(t_b(zero_d), zj, h_b(zero_d), uj, wj, tab)
        if J is None:
            J = np.stack([out[a].d for a in range(NZ)], axis=1)             # [K, NZ(out), ND(in)]
        dP_v[o] = np.stack([out[NZ + f].v for f in range(nF)], axis=1)
        dP_d[o] = np.stack([out[NZ + f].d for f in range(nF)], axis=1)

    # ---- phase 2b: interval quadratures, F, criterion, Lambda ----
    N = orc.N
    pq = np.zeros((nobs, N, nF))
    for k in range(K):
        pq[:, iv[k]] += dP_v[:, k]
    w_iv = [[design[lay.w_offsets[o] + int(orc.grid.indicators[nc + o, j]) - 1] for j in range(N)] for o in range(nobs)]
    Fv = np.zeros(nF)
    for j in range(N):
        for o in range(nobs):
            Fv += w_iv[o][j] * pq[o, j]
    tau = design[lay.tau_offset]
    Fm = symmetric_from_vector(Fv[None])
    obj = float((criterion(orc.crit, Fm, np.array([tau])) + tau)[0])
    Lam = criterion_matrix_gradient(orc.crit, Fm, np.array([tau]))[0]      # [np, np]
    # d objective = sum_f Ls[f] dF[f] with the off-diagonal entries of the symmetric matrix counted twice
    eyeF = symmetric_from_vector(np.eye(nF))                                # [nF, np, np]: d M / d F_f
    Ls = np.einsum("ij,fij->f", Lam, eyeF)
    s = np.zeros((K, ND))
    for o in range(nobs):
        s += w_k[o][:, None] * np.einsum("f,kfd->kd", Ls, dP_d[o])

    grad = np.zeros(orc.n_var)
    for o in range(nobs):                                                   # d/dw from the interval quadratures
        for j in range(N):
            grad[lay.w_offsets[o] + int(orc.grid.indicators[nc + o, j]) - 1] += Ls @ pq[o, j]
    grad[lay.tau_offset] = np.trace(Lam) + 1.0

    # ---- phase 3: discrete adjoint lambda_k = s_k[:NZ] + J_k[:, :NZ]^T lambda_{k+1}, chunked ----
    Jz, Ju, sz, su = J[:, :, :NZ], J[:, :, NZ:], s[:, :NZ], s[:, NZ:]
    comp = []
    for lo, hi in chunks:                                                   # lambda_lo = b + A lambda_hi
        # composed column by column, as the kernel does: push e_q (and 0 with the sources) down the chunk,
        # later steps first
        cols = []
        for q in range(NZ + 1):
            v = np.zeros(NZ)
            if q < NZ:
                v[q] = 1.0
            for k in range(hi - 1, lo - 1, -1):
                v = (sz[k] if q == NZ else 0.0) + Jz[k].T @ v
            cols.append(v)
        comp.append((np.stack(cols[:NZ], axis=1), cols[NZ]))
    lam_hi = [None] * len(chunks)
    lam = np.zeros(NZ)
    for c in range(len(chunks) - 1, -1, -1):
        lam_hi[c] = lam
        A, b = comp[c]
        lam = b + A @ lam
    lam0 = lam
    part = np.zeros((len(chunks), nc, max([1] + [max(ci) + 1 for ci in ctrl_idx])))
    for c, (lo, hi) in enumerate(chunks):                                   # replay: control columns
        lam = lam_hi[c]
        for k in range(hi - 1, lo - 1, -1):
            for cc in range(nc):
                part[c, cc, ctrl_idx[cc][k]] += su[k, cc] + Ju[k][:, cc] @ lam
            lam = sz[k] + Jz[k].T @ lam
    for cc in range(nc):
        for idx in set(ctrl_idx[cc]):
            grad[lay.control_offsets[cc] + idx] = sum(part[c, cc, idx] for c in range(len(chunks) - 1, -1, -1))
    for a, st in enumerate(orc.ic_state):
        grad[lay.ic_offset + a] = lam0[orc.d_idx.index(st)]
    return obj, grad, Fv, diag
