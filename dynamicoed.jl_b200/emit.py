"""CUDA emitter: prints the simplified augmented system into device functions for the
hand-written sm_100a kernel template (`csrc/dynoed_kernels.cuh`).

Role in the north-star design: "host code builds the simplified RHS and emits its body into a
hand-written kernel template, compiled once".  The reference gets the same functions from
ModelingToolkit's code generation (``ODAEProblem``, /root/reference/src/problem.jl:82) and its
derivatives from ForwardDiff duals (:154); here the host emits

  consts  : sub-expressions that depend only on (controls, parameters) — hoisted out of the
            integrator's stage loop, recomputed once per merged-grid interval
  primal  : x' = f,  G' = J G + f_p  (J = df/dx; /root/reference/src/augment.jl:231),
            F'_triu = sum_i w_i (Q_i^T Q_i)_triu with Q = h_x G (:223-226, :238)
  adjoint : vector-Jacobian product of the (x, G) right-hand side plus the gradient of the
            Lambda-weighted Fisher integrand — the stage kernel of the discrete RK adjoint that
            yields d objective / d(controls, initial conditions, weights) in one backward sweep
  obs     : Q = h_x G (for the information-gain outputs, /root/reference/src/utilities.jl:19-54)

The matrix structure (J, f_p, h_x as named temporaries; products kept factored) is preserved so
that nvcc contracts almost everything into DFMA.  Flop counts of the emitted straight-line code
(FMA = 2, mul/add = 1) are frozen next to the code for the roofline accounting.
"""
from __future__ import annotations

import hashlib
import json
from collections import OrderedDict

import sympy as sp
from sympy.printing.c import C99CodePrinter

from .augment import SimplifiedOEDSystem, triu_pairs
from .symbolic import getdefault, is_initial_condition


class _Printer(C99CodePrinter):
    def __init__(self, names):
        super().__init__({"contract": False})
        self._names = names

    def _print_Symbol(self, expr):
        return self._names.get(expr, expr.name)

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                s = f"({s})"
            return "*".join([s] * int(e))
        if e == -1:
            return f"1.0/({self._print(b)})"
        if e.is_Integer and -4 <= int(e) <= -2:
            s = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                s = f"({s})"
            return "1.0/(" + "*".join([s] * int(-e)) + ")"
        return super()._print_Pow(expr)

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"


class _FmaPrinter(_Printer):
    """Prints every sum of products as ONE explicit fma chain seeded with its plain addends
    (``fma(z3, J01, fma(z2, J00, fp00))`` instead of ``z2*J00 + z3*J01 + fp00``): a C compiler may
    contract a*b+c but must not reassociate, so the left-to-right sum costs a DMUL, a DFMA and a
    DADD where the chain costs two DFMAs.  double blocks only (the dual-number templates keep
    operator syntax)."""

    def _factor(self, f):
        s = self._print(f)
        return f"({s})" if (f.is_Add or (f.is_Mul and len(f.args) > 1)) else s

    def _print_Add(self, expr):
        plain, prods = [], []
        for t in expr.as_ordered_terms():
            c, rest = t.as_coeff_Mul()
            fac = [] if rest == 1 else list(sp.Mul.make_args(rest))
            expanded = []
            for f in fac:          # x**2 is two factors
                if f.is_Pow and f.exp.is_Integer and 2 <= int(f.exp) <= 4:
                    expanded += [f.base] * int(f.exp)
                else:
                    expanded.append(f)
            nfac = len(expanded) + (0 if c in (1, -1) else 1)
            (prods if nfac >= 2 else plain).append((c, expanded, t))
        if plain:
            acc = super()._print_Add(sp.Add(*[t for _, _, t in plain])) if len(plain) > 1 \
                else self._print(plain[0][2])
        else:
            acc = self._print(prods.pop(0)[2])
        for c, fac, _ in prods:
            if c in (1, -1):
                a = ("-" if c == -1 else "") + self._factor(fac[0])
                rest = fac[1:]
            else:
                a, rest = self._print(c), fac
            b = "*".join(self._factor(f) for f in rest)
            acc = f"fma({a}, {b}, {acc})"
        return acc


def count_flops(expr) -> int:
    """mul/add = 1 (an FMA therefore counts 2), division / elementary function = 1."""
    expr = sp.sympify(expr)
    n = 0
    for node in sp.preorder_traversal(expr):
        if node.is_Add:
            n += len(node.args) - 1
        elif node.is_Mul:
            args = [a for a in node.args if a != -1 and a != 1]
            n += max(len(args) - 1, 0)
        elif node.is_Pow:
            e = node.exp
            if e.is_Integer:
                n += abs(int(e)) - 1 + (1 if int(e) < 0 else 0)
            else:
                n += 1
        elif node.is_Function:
            n += 1
    return n


class _Hoister:
    """Replaces maximal sub-expressions that depend only on `static` symbols by c[k] symbols."""

    def __init__(self, static):
        self.static = set(static)
        self.table = OrderedDict()

    def _is_static(self, e):
        return len(e.free_symbols) > 0 and e.free_symbols <= self.static

    def _const(self, e):
        if e.is_Symbol:
            return e
        if e.could_extract_minus_sign() and (-e).is_Symbol:
            return e
        if e not in self.table:
            self.table[e] = sp.Symbol(f"hc{len(self.table)}", real=True)
        return self.table[e]

    def __call__(self, e):
        e = sp.sympify(e)
        if e.is_Atom:
            return e
        if self._is_static(e):
            return self._const(e)
        args = [self(a) for a in e.args]
        if e.is_Add or e.is_Mul:
            st = [a for a in args if (a.free_symbols and a.free_symbols <= self.static | set(self.table.values()))
                  or (a.is_Number and e.is_Add)]
            dyn = [a for a in args if a not in st]
            nontrivial = [a for a in st if not a.is_Number]
            if len(nontrivial) >= 2 or (len(nontrivial) == 1 and len(st) >= 2 and e.is_Add):
                folded = e.func(*st)
                folded = folded.xreplace({v: k for k, v in self.table.items()})
                return e.func(self._const(folded), *dyn)
        return e.func(*args)


def _toposort(defs, inputs):
    """defs: list of (symbol, expr).  Stable topological order w.r.t. symbol dependencies."""
    defined = set(inputs)
    names = {s for s, _ in defs}
    out, pending = [], list(defs)
    while pending:
        rest, progressed = [], False
        for s, e in pending:
            deps = {d for d in e.free_symbols if d in names}
            if deps <= defined:
                out.append((s, e))
                defined.add(s)
                progressed = True
            else:
                rest.append((s, e))
        assert progressed, "cyclic definitions in emitted block"
        pending = rest
    return out


class _Block:
    def __init__(self):
        self.defs = []      # temporaries (symbol, expr)
        self.outs = []      # (c lvalue string, expr, accumulate?)

    def tmp(self, name, expr):
        expr = sp.sympify(expr)
        if expr.is_Number or expr.is_Symbol:
            return expr
        if expr.could_extract_minus_sign() and (-expr).is_Symbol:
            return expr
        s = sp.Symbol(name, real=True)
        self.defs.append((s, expr))
        return s

    def out(self, lvalue, expr, accumulate=False):
        self.outs.append((lvalue, sp.sympify(expr), accumulate))

    def render(self, printer, inputs, indent="    ", scalar="double"):
        exprs = [e for _, e in self.defs] + [e for _, e, _ in self.outs]
        repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("v", real=True), order="none")
        nd = len(self.defs)
        defs = list(repl) + [(s, red[i]) for i, (s, _) in enumerate(self.defs)]
        tnames = {s for s, _ in defs}
        defs = _toposort(defs, inputs)
        lines, flops = [], 0
        for s, e in defs:
            lines.append(f"{indent}const {scalar} {s.name} = {printer.doprint(e)};")
            flops += count_flops(e)
        for (lv, _, acc), e in zip(self.outs, red[nd:]):
            lines.append(f"{indent}{lv} {'+=' if acc else '='} {printer.doprint(e)};")
            flops += count_flops(e) + (1 if acc else 0)
        return "\n".join(lines), flops


def _horner(e, wrt):
    e = sp.sympify(e)
    try:
        if e.is_polynomial(*wrt) and e.free_symbols & set(wrt):
            return sp.horner(sp.expand(e), wrt=wrt)
    except Exception:
        pass
    return e


def emit_model(name: str, s: SimplifiedOEDSystem) -> tuple[str, dict]:
    nx, np_, n_obs = s.nx, s.np_, s.n_obs
    nc = len(s.controls)
    n_ic = len(s.x_ic)
    n_ord = np_ - n_ic
    nF = np_ * (np_ + 1) // 2
    nz = nx + nx * np_
    t = s.iv
    theta = [p for p in s.ps if p not in s.controls and p not in s.w and not is_initial_condition(p)]
    x = list(s.x)
    G = s.G
    p_ord = s.p_tunable[:n_ord]

    # --- structured pieces ---------------------------------------------------------------------
    f = sp.Matrix(s.f)
    J = f.jacobian(x)
    fp = f.jacobian(p_ord) if n_ord else sp.zeros(nx, 0)
    # consistency with the reference-faithful derivation (0 = f_xdot G' + f_x G + f_p, torn)
    fp_full = sp.Matrix(nx, np_, lambda i, j: fp[i, j] if j < n_ord else 0)
    chk = (J * G + fp_full - s.dG).applyfunc(sp.expand)
    assert all(sp.simplify(e) == 0 for e in chk), "structured G' differs from the augmented system"
    # observation sensitivities Q = Hc G + H0
    Hc = sp.Matrix(n_obs, nx, lambda i, c: sp.diff(s.Qexpr[i, 0], G[c, 0]))
    H0 = sp.Matrix(n_obs, np_, lambda i, j: sp.simplify(
        s.Qexpr[i, j] - sum(Hc[i, c] * G[c, j] for c in range(nx))))
    assert not any(e.has(*list(G)) for e in list(Hc) + list(H0)), "Q must be affine in G"
    # check F' against the augmented system
    wsym = list(s.w)
    for k, (i, j) in enumerate(triu_pairs(np_)):
        e = sum(wsym[o] * s.Qexpr[o, i] * s.Qexpr[o, j] for o in range(n_obs)) - s.dF[k]
        assert sp.simplify(sp.expand(e)) == 0, "structured F' differs from the augmented system"

    # --- C names ---------------------------------------------------------------------------------
    names = {t: "t"}
    for i, v in enumerate(x):
        names[v] = f"z[{i}]"
    for j in range(np_):
        for i in range(nx):
            names[G[i, j]] = f"z[{nx + j * nx + i}]"
    for k, v in enumerate(theta):
        names[v] = f"th[{k}]"
    for k, v in enumerate(s.controls):
        names[v] = f"u[{k}]"
    for k, v in enumerate(wsym):
        names[v] = f"w[{k}]"
    kb = [sp.Symbol(f"kb{a}", real=True) for a in range(nz)]
    # cwL[i*nF + k] = 2 w_i Lambda_k (packed symmetric; the caller forms it once per merged interval)
    cwLs = [[sp.Symbol(f"cwL{i}_{k}", real=True) for k in range(nF)] for i in range(n_obs)]
    for a, v in enumerate(kb):
        names[v] = f"kb[{a}]"
    for i in range(n_obs):
        for k, v in enumerate(cwLs[i]):
            names[v] = f"cwL[{i * nF + k}]"
    cwLfull = []
    for i in range(n_obs):
        m = sp.zeros(np_, np_)
        for k, (a, b_) in enumerate(triu_pairs(np_)):
            m[a, b_] = cwLs[i][k]
            m[b_, a] = cwLs[i][k]
        cwLfull.append(m)

    hoist = _Hoister(set(theta) | set(s.controls))
    wrt = x
    H = lambda e: hoist(_horner(e, wrt))

    def H_best(e):
        """Horner form or factored form, whichever is cheaper after hoisting (x' = x (a - c y) and
        J_00 = a - c y then share one sub-expression)."""
        import copy
        cands = [_horner(e, wrt)]
        try:
            cands.append(sp.factor(e))
        except Exception:
            pass
        best, cost = None, None
        for cnd in cands:
            trial = copy.deepcopy(hoist)
            k = count_flops(trial(cnd))
            if cost is None or k < cost:
                best, cost = cnd, k
        return hoist(best)

    # ---------------- primal ---------------------------------------------------------------------
    def build_primal():
        b = _Block()
        Jt = sp.Matrix(nx, nx, lambda i, c: b.tmp(f"J{i}{c}", H(J[i, c])))
        fpt = sp.Matrix(nx, max(n_ord, 0), lambda i, j: b.tmp(f"fp{i}{j}", H(fp[i, j]))) \
            if n_ord else sp.zeros(nx, 0)
        for i in range(nx):
            b.out(f"dz[{i}]", H_best(f[i]))
        for j in range(np_):
            for i in range(nx):
                e = sum(Jt[i, c] * G[c, j] for c in range(nx))
                if j < n_ord:
                    e = e + fpt[i, j]
                b.out(f"dz[{nx + j * nx + i}]", e)
        return b

    # ---------------- adjoint --------------------------------------------------------------------
    def build_adjoint():
        b = _Block()
        kx = kb[:nx]
        kG = sp.Matrix(nx, np_, lambda i, j: kb[nx + j * nx + i])
        Jt = sp.Matrix(nx, nx, lambda i, c: b.tmp(f"J{i}{c}", H(J[i, c])))
        Hct = sp.Matrix(n_obs, nx, lambda i, c: H(Hc[i, c]))
        Qe = sp.Matrix(n_obs, np_, lambda i, j: sum(Hct[i, c] * G[c, j] for c in range(nx)) + H(H0[i, j]))
        # rw_ij = (2 w_i Lambda Q_i^T)_j  — plain expressions: cse names the ones used more than once
        rw = sp.Matrix(n_obs, np_, lambda i, j: sum(cwLfull[i][j, k] * Qe[i, k] for k in range(np_)))
        # G-bar = J^T kG + sum_i Hc[i,:]^T rw_i
        for j in range(np_):
            for c in range(nx):
                e = sum(Jt[i, c] * kG[i, j] for i in range(nx))
                for i in range(n_obs):
                    if Hc[i, c] != 0:
                        e = e + Hct[i, c] * rw[i, j]
                b.out(f"zb[{nx + j * nx + c}]", e)
        # scalar S = kx.f + <W, J> + <kG_ord, fp> + sum_ic m_ic Hc_ic + sum_ij n_ij H0_ij;
        # x-bar = dS/dx, u-bar = dS/du (W, m, n treated as constants)
        Wsym = sp.Matrix(nx, nx, lambda i, c: sp.Symbol(f"W{i}{c}", real=True))
        msym = sp.Matrix(n_obs, nx, lambda i, c: sp.Symbol(f"m{i}{c}", real=True))
        nsym = sp.Matrix(n_obs, np_, lambda i, j: sp.Symbol(f"n{i}{j}", real=True))
        S = sum(kx[i] * f[i] for i in range(nx))
        S += sum(Wsym[i, c] * J[i, c] for i in range(nx) for c in range(nx))
        S += sum(kG[i, j] * fp[i, j] for i in range(nx) for j in range(n_ord))
        S += sum(msym[i, c] * Hc[i, c] for i in range(n_obs) for c in range(nx))
        S += sum(nsym[i, j] * H0[i, j] for i in range(n_obs) for j in range(np_))
        lin = list(kx) + list(Wsym) + list(kG) + list(msym) + list(nsym)
        grads = [sp.diff(S, v) for v in x] + [sp.diff(S, v) for v in s.controls]
        used = set().union(*[g.free_symbols for g in grads]) if grads else set()
        sub = {}
        for i in range(nx):
            for c in range(nx):
                if Wsym[i, c] in used:
                    sub[Wsym[i, c]] = b.tmp(f"W{i}{c}", sum(kG[i, j] * G[c, j] for j in range(np_)))
        for i in range(n_obs):
            for c in range(nx):
                if msym[i, c] in used:
                    sub[msym[i, c]] = b.tmp(f"m{i}{c}", sum(G[c, j] * rw[i, j] for j in range(np_)))
            for j in range(np_):
                if nsym[i, j] in used:
                    sub[nsym[i, j]] = b.tmp(f"n{i}{j}", rw[i, j])
        static = [v for v in theta] + list(hoist.table.values())

        def collect_lin(g):
            g = sp.expand(g)
            out = 0
            rest = g
            for v in lin:
                co = g.coeff(v)
                if co != 0:
                    out = out + v * H(co)
                    rest = sp.expand(rest - v * co)
            assert rest == 0, "adjoint scalar must be linear in the adjoint symbols"
            out = out.xreplace(sub)
            # -p3 x kx0 - p3 W00 - ...  ->  -p3 (x kx0 + W00) - ... when that is cheaper
            cands = []
            try:
                cands.append(sp.collect(sp.expand(out), [v for v in static if out.has(v)]))
            except Exception:
                pass
            try:
                # (kb5 - kb2) z1 instead of kb5 z1 - kb2 z1: common STATE factors of the adjoint symbols
                cands.append(sp.collect(sp.expand(out), [v for v in x if out.has(v)]))
                cands.append(sp.collect(sp.collect(sp.expand(out), [v for v in x if out.has(v)]),
                                        [v for v in static if out.has(v)]))
            except Exception:
                pass
            for col in cands:
                if count_flops(col) < count_flops(out):
                    out = col
            return out
        for a in range(nx):
            b.out(f"zb[{a}]", collect_lin(grads[a]))
        for k in range(nc):
            b.out(f"ub[{k}]", collect_lin(grads[nx + k]))
        return b

    def build_obs():
        b = _Block()
        for i in range(n_obs):
            for j in range(np_):
                b.out(f"Q[{i * np_ + j}]",
                      sum(H(Hc[i, c]) * G[c, j] for c in range(nx)) + H(H0[i, j]))
        return b

    # ---------------- Rosenbrock pieces (stiff / DAE configuration) ---------------------------
    # The Rosenbrock step works on y = [x; vec(G); P] with P the UNWEIGHTED per-observation
    # quadratures (F = sum_i w_i P_i is linear in w because nothing depends on F).  The Jacobian of
    # the augmented system is block lower triangular with the SAME diagonal block f_x for x and
    # for every column of G, so one nx x nx LU per step serves all
    # 1 + np right-hand sides; what remains are the off-diagonal products emitted here:
    #   ros_f    : g(y)  = [f; J G + f_p] and the quadrature integrands q_i = (Q_i^T Q_i)_triu
    #   ros_jac  : A     = f_x(x)
    #   ros_ling : B_j U_x = d(J G_j + f_p,j)/dx . U_x          (needs second derivatives of f)
    #   ros_linp : dq_i/dy . U                                   (x- and G-directions)
    Usym = [sp.Symbol(f"U{a}", real=True) for a in range(nz)]
    for a, v in enumerate(Usym):
        names[v] = f"U[{a}]"

    def build_ros_f():
        b = _Block()
        Jt = sp.Matrix(nx, nx, lambda i, c: b.tmp(f"J{i}{c}", H(J[i, c])))
        fpt = sp.Matrix(nx, max(n_ord, 0), lambda i, j: b.tmp(f"fp{i}{j}", H(fp[i, j]))) \
            if n_ord else sp.zeros(nx, 0)
        for i in range(nx):
            b.out(f"dz[{i}]", H(f[i]))
        for j in range(np_):
            for i in range(nx):
                e = sum(Jt[i, c] * G[c, j] for c in range(nx))
                if j < n_ord:
                    e = e + fpt[i, j]
                b.out(f"dz[{nx + j * nx + i}]", e)
        Hct = sp.Matrix(n_obs, nx, lambda i, c: b.tmp(f"h{i}{c}", H(Hc[i, c])))
        Qt = sp.Matrix(n_obs, np_, lambda i, j: b.tmp(
            f"Q{i}{j}", sum(Hct[i, c] * G[c, j] for c in range(nx)) + H(H0[i, j])))
        for o in range(n_obs):
            for k, (i, j) in enumerate(triu_pairs(np_)):
                b.out(f"dP[{o * nF + k}]", Qt[o, i] * Qt[o, j])
        return b

    def build_ros_jac():
        b = _Block()
        for i in range(nx):
            for c in range(nx):
                b.out(f"A[{i * nx + c}]", H(J[i, c]))
        return b

    def build_ros_ling():
        b = _Block()
        for j in range(np_):
            for i in range(nx):
                gij = sum(J[i, c] * G[c, j] for c in range(nx)) + (fp[i, j] if j < n_ord else 0)
                e = sum(H(sp.diff(gij, x[c])) * Usym[c] for c in range(nx))
                b.out(f"BU[{j * nx + i}]", e)
        return b

    # ros_hvp: (d f_x / d eps) U for the direction (xd, ud) of (x, controls) — what a forward-mode
    # direction needs to reuse the VALUE factorisation of  W = 1/(h gamma) I - f_x :
    #   W0 U1 = r1 - W1 U0 = r1 + (d f_x . (xd, ud)) U0
    xds = [sp.Symbol(f"xd{k}", real=True) for k in range(nx)]
    uds = [sp.Symbol(f"ud{k}", real=True) for k in range(nc)]
    for k, v in enumerate(xds):
        names[v] = f"xd[{k}]"
    for k, v in enumerate(uds):
        names[v] = f"ud[{k}]"

    def build_ros_hvp():
        b = _Block()
        for i in range(nx):
            e = 0
            for j in range(nx):
                dA = sum(sp.diff(J[i, j], x[k]) * xds[k] for k in range(nx)) + \
                    sum(sp.diff(J[i, j], s.controls[m]) * uds[m] for m in range(nc))
                if dA != 0:
                    e = e + H(sp.expand(dA)) * Usym[j]
            b.out(f"out[{i}]", e)
        return b

    def build_ros_linp():
        b = _Block()
        zs = x + [G[i, j] for j in range(np_) for i in range(nx)]
        for o in range(n_obs):
            for k, (i, j) in enumerate(triu_pairs(np_)):
                q = s.Qexpr[o, i] * s.Qexpr[o, j]
                e = sum(H(sp.diff(q, zs[a])) * Usym[a] for a in range(nz))
                b.out(f"dPU[{o * nF + k}]", e)
        return b

    # ---------------- second-order pieces: gradient with respect to unknown initial conditions ----
    # d/d(ic_a) of the Rosenbrock scheme on (x, G, P) IS the same scheme on the system extended by
    #   H_(j,a) = d G_j / d ic_a   (d x / d ic_a is the G column of that initial condition itself)
    #   dP_a    = d P   / d ic_a
    # because the scheme uses the exact Jacobian.  The extended Jacobian stays block lower triangular
    # with diagonal block f_x, so every H column is one more right-hand side for the step's LU.
    # Emitted per unique column q = (j, a) (H_(j,a) = H_(a',j') when both are initial conditions):
    #   ic2_rhs<q>  : g = J(Z) ZH + d(J G_j + f_p,j)/dx . G_a        at the stage state
    #   ic2_lin<q>  : b = dg/d(x, G) . U                              at y_n (off-diagonal blocks)
    #   ic2_qbase   : rows of dP that do not involve H:  rhs(Z) + lin(y_n; U)
    #   ic2_qcol<q> : the part of the dP rows that is linear in (H_q, U_Hq)
    # G column that IS d x / d ic_a: the one whose initial value is the unit vector of that state (the
    # reference assigns them in its own order, quirk Q3 of /root/reference/src/augment.jl:186); -1 if the
    # default G0 has no such column (then the second-order kernel is not offered)
    ic_state_ = [s.x.index(s.aug.x[i - 1]) for i in s.aug.unknown_initial_conditions]
    G0m = [[float(s.u0[nx + j * nx + i]) for j in range(np_)] for i in range(nx)]
    ga = []
    for a in range(n_ic):
        cols = [j for j in range(n_ord, np_)
                if all(G0m[i][j] == (1.0 if i == ic_state_[a] else 0.0) for i in range(nx))]
        ga.append(cols[0] if cols else -1)
    ic2_possible = n_ic > 0 and all(c >= 0 for c in ga) and len(set(ga)) == n_ic
    if not ic2_possible:
        ga = [n_ord + a for a in range(n_ic)]
    col_ic = {c: a for a, c in enumerate(ga)}
    hcols, hidx = [], {}
    for a in range(n_ic):
        for j in range(np_):
            key = (j, a)
            if j in col_ic and col_ic[j] < a:        # H_(ga[a'], a) = d2 x / d ic_a' d ic_a = H_(ga[a], a')
                key = (ga[a], col_ic[j])
            if key not in hidx:
                hidx[key] = len(hcols)
                hcols.append(key)
            hidx[(j, a)] = hidx[key]
    nh = len(hcols)
    nQ = n_obs * nF
    ic2_code, ic2_flops = {}, {}
    if n_ic:
        names2 = dict(names)
        for i, v in enumerate(x):
            names2[v] = f"Z[{i}]"
        for j in range(np_):
            for i in range(nx):
                names2[G[i, j]] = f"Z[{nx + j * nx + i}]"
        xn = [sp.Symbol(f"xn{i}", real=True) for i in range(nx)]
        Gn = sp.Matrix(nx, np_, lambda i, j: sp.Symbol(f"Gn{i}_{j}", real=True))
        for i, v in enumerate(xn):
            names2[v] = f"zn[{i}]"
        for j in range(np_):
            for i in range(nx):
                names2[Gn[i, j]] = f"zn[{nx + j * nx + i}]"
        ZH = [[sp.Symbol(f"ZH{q}_{i}", real=True) for i in range(nx)] for q in range(nh)]
        HN = [[sp.Symbol(f"HN{q}_{i}", real=True) for i in range(nx)] for q in range(nh)]
        UH = [[sp.Symbol(f"UH{q}_{i}", real=True) for i in range(nx)] for q in range(nh)]
        for q in range(nh):
            for i in range(nx):
                names2[ZH[q][i]] = f"ZH[{i}]"
                names2[HN[q][i]] = f"Hn[{i}]"
                names2[UH[q][i]] = f"UH[{i}]"
        to_n = {x[i]: xn[i] for i in range(nx)}
        to_n.update({G[i, j]: Gn[i, j] for i in range(nx) for j in range(np_)})
        to_n.update({ZH[q][i]: HN[q][i] for q in range(nh) for i in range(nx)})
        zsyms = x + [G[i, j] for j in range(np_) for i in range(nx)]
        g1 = [[sum(J[i, c] * G[c, j] for c in range(nx)) + (fp[i, j] if j < n_ord else 0)
               for i in range(nx)] for j in range(np_)]
        g2 = [[sum(J[i, c] * ZH[q][c] for c in range(nx)) +
               sum(sp.diff(g1[j][i], x[c]) * G[c, ga[a]] for c in range(nx))
               for i in range(nx)] for q, (j, a) in enumerate(hcols)]
        q2 = []          # rows a*nQ + o*nF + k
        for a in range(n_ic):
            for o in range(n_obs):
                for k, (i, j) in enumerate(triu_pairs(np_)):
                    q1 = s.Qexpr[o, i] * s.Qexpr[o, j]
                    e = sum(sp.diff(q1, x[c]) * G[c, ga[a]] for c in range(nx))
                    e += sum(sp.diff(q1, G[c, jj]) * ZH[hidx[(jj, a)]][c]
                             for c in range(nx) for jj in range(np_))
                    q2.append(sp.expand(e))
        allZH = [v for col in ZH for v in col]
        allHN = [v for col in HN for v in col]
        allUH = [v for col in UH for v in col]
        hz = lambda e: _horner(sp.expand(e), x + xn)
        pr2 = _FmaPrinter(names2)
        inputs2 = set(names2.keys())

        def chain(blocks_q, signature):
            body, fl = [], []
            for q, b in enumerate(blocks_q):
                code_q, f_q = b.render(pr2, inputs2, indent="            ")
                kw = "if" if q == 0 else "else if"
                body.append(f"        {kw} constexpr (Q == {q}) {{\n{code_q}\n        }}")
                fl.append(f_q)
            return "\n".join(body), fl

        rhs_b, lin_b, qcol_b = [], [], []
        for q in range(nh):
            b = _Block()
            for i in range(nx):
                b.out(f"g[{i}]", hz(g2[q][i]))
            rhs_b.append(b)
            b = _Block()
            for i in range(nx):
                e = sum(sp.diff(g2[q][i], v) * Usym[k] for k, v in enumerate(zsyms))
                b.out(f"b[{i}]", hz(sp.expand(e).xreplace(to_n)))
            lin_b.append(b)
        # quadrature-derivative rows: stage-state part and the coupling at y_n
        qlin = []
        for e in q2:
            l = sum(sp.diff(e, v) * Usym[k] for k, v in enumerate(zsyms))
            l += sum(sp.diff(e, ZH[q][i]) * UH[q][i] for q in range(nh) for i in range(nx))
            qlin.append(sp.expand(sp.expand(l).xreplace(to_n)))
        zeroH = {v: 0 for v in allZH + allHN + allUH}
        b = _Block()
        for r, e in enumerate(q2):
            b.out(f"rb[{r}]", hz(e.xreplace(zeroH)) + hz(qlin[r].xreplace(zeroH)))
        ic2_code["qbase"], ic2_flops["qbase"] = b.render(pr2, inputs2, indent="        ")
        for q in range(nh):
            b = _Block()
            for r, e in enumerate(q2):
                part = sum(sp.diff(e, v) * v for v in ZH[q])
                part += sum(sp.diff(qlin[r], v) * v for v in HN[q] + UH[q])
                part = sp.expand(part)
                if part != 0:
                    b.out(f"acc[{r}]", hz(part), accumulate=True)
            qcol_b.append(b)
        ic2_rows = [sum(1 << int(lv[4:-1]) for lv, _, _ in b.outs) for b in qcol_b]
        ic2_base_rows = sum(1 << r for r, e in enumerate(q2)
                            if (e.xreplace(zeroH) != 0 or qlin[r].xreplace(zeroH) != 0))
        ic2_code["rhs"], ic2_flops["rhs"] = chain(rhs_b, None)
        ic2_code["lin"], ic2_flops["lin"] = chain(lin_b, None)
        ic2_code["qcol"], ic2_flops["qcol"] = chain(qcol_b, None)

    blocks = OrderedDict(primal_z=build_primal(),
                         adjoint=build_adjoint(), obs=build_obs(), ros_f=build_ros_f(),
                         ros_jac=build_ros_jac(), ros_ling=build_ros_ling(),
                         ros_linp=build_ros_linp(), ros_hvp=build_ros_hvp())
    for k, v in enumerate(hoist.table.values()):
        names[v] = f"c[{k}]"
    pr = _Printer(names)
    prf = _FmaPrinter(names)       # double blocks: explicit fma chains
    inputs = set(names.keys())
    code, flops = {}, {}
    for key, blk in blocks.items():
        tmpl = key.startswith("ros_") and key != "ros_hvp"
        code[key], flops[key] = blk.render(pr if tmpl else prf, inputs, scalar="T" if tmpl else "double")
    # what one forward stage of the explicit integrators evaluates: (x, G) right-hand side + Q
    flops["primal"] = flops["primal_z"] + flops["obs"]
    if n_ic:
        flops["ic2_col"] = sum(ic2_flops["rhs"]) + sum(ic2_flops["lin"])
        flops["ic2_q"] = ic2_flops["qbase"] + sum(ic2_flops["qcol"])
    else:
        flops["ic2_col"] = flops["ic2_q"] = 0
    cb = _Block()
    for k, e in enumerate(hoist.table.keys()):
        cb.out(f"c[{k}]", e)
    code["consts"], flops["consts"] = cb.render(prf, inputs)
    code["consts_t"], _ = cb.render(pr, inputs, scalar="T")
    n_const = max(len(hoist.table), 1)
    autonomous = not any(sp.sympify(e).has(t) for e in list(f) + list(s.Qexpr))

    ic_state = [s.x.index(s.aug.x[i - 1]) for i in s.aug.unknown_initial_conditions]
    x0 = [float(v) for v in s.u0[:nx]]
    G0 = [float(v) for v in s.u0[nx:nz]]
    th0 = [float(s.defaults.get(p, getdefault(p, 0.0))) for p in theta]
    info = dict(name=name, nx=nx, np=np_, n_obs=n_obs, n_ctrl=nc, n_ic=n_ic, n_theta=len(theta),
                n_const=len(hoist.table), nz=nz, nF=nF, nh=nh, ic_state=ic_state, x0=x0, G0=G0,
                theta=th0, theta_names=[p.name for p in theta],
                control_names=[c.name for c in s.controls], w_names=[w.name for w in s.w],
                ic_names=[p.name for p in s.x_ic], tunable_names=[p.name for p in s.p_tunable],
                state_names=[v.name for v in s.states], flops=flops)

    def arr(vals, fmt="{!r}"):
        return ", ".join(fmt.format(v) for v in vals) if vals else "0"

    cname = "Model_" + name
    ic2_src = ""
    if n_ic:
        ic2_src = f"""    // ---- second-order sensitivities for the unknown-initial-condition gradient (see emit.py) ----
    // unique H columns q -> (parameter column j, initial condition a): {', '.join(f'{q}:({j},{a})' for q, (j, a) in enumerate(hcols))}
    // Z: stage state (x, G), zn: (x, G) at the step start, U: this stage's increments of (x, G) — arrays or
    // accessors with operator[]; ZH / Hn / UH: the column's own 6-vectors; dP rows: a*NOBS*NF + o*NF + k
    static constexpr unsigned long long IC2_QBASE_ROWS = {ic2_base_rows}ull;   // dP rows with an H-free part
    __host__ __device__ static constexpr unsigned long long ic2_qrows(int q) {{   // dP rows column q feeds
        constexpr unsigned long long v[{max(nh, 1)}] = {{{', '.join(f'{r}ull' for r in ic2_rows)}}};
        return v[q];
    }}
    template <int Q, class ZA>
    __device__ __forceinline__ static void ic2_rhs(double t, const ZA& Z, const double* __restrict__ ZH,
            const double* __restrict__ u, const double* __restrict__ th, double* __restrict__ g) {{
{ic2_code['rhs']}
    }}
    template <int Q, class NA, class UA>
    __device__ __forceinline__ static void ic2_lin(double t, const NA& zn, const double* __restrict__ Hn,
            const UA& U, const double* __restrict__ u, const double* __restrict__ th,
            double* __restrict__ b) {{
{ic2_code['lin']}
    }}
    template <class ZA, class NA, class UA>
    __device__ __forceinline__ static void ic2_qbase(double t, const ZA& Z, const NA& zn, const UA& U,
            const double* __restrict__ u, const double* __restrict__ th, double* __restrict__ rb) {{
{ic2_code['qbase']}
    }}
    template <int Q, class ZA, class NA, class UA>
    __device__ __forceinline__ static void ic2_qcol(double t, const ZA& Z, const double* __restrict__ ZH,
            const NA& zn, const double* __restrict__ Hn, const UA& U, const double* __restrict__ UH,
            const double* __restrict__ u, const double* __restrict__ th, double* __restrict__ acc) {{
{ic2_code['qcol']}
    }}
"""
    src = f"""// GENERATED by dynamicoed.jl_b200/emit.py — do not edit.  Model "{name}".
// states  : {', '.join(info['state_names'])}
// tunables: {', '.join(info['tunable_names'])}   controls: {', '.join(info['control_names']) or '-'}
// theta   : {', '.join(info['theta_names'])}
// flops (FMA=2): {json.dumps(flops)}
struct {cname} {{
    static constexpr int NX = {nx}, NP = {np_}, NOBS = {n_obs}, NCTRL = {nc}, NIC = {n_ic};
    static constexpr int NTH = {len(theta)}, NCONST = {n_const}, NZ = {nz}, NF = {nF};
    static constexpr int FLOPS_PRIMAL = {flops['primal']}, FLOPS_PRIMAL_Z = {flops['primal_z']};
    static constexpr int FLOPS_ADJOINT = {flops['adjoint']}, FLOPS_CONSTS = {flops['consts']};
    static constexpr int FLOPS_OBS = {flops['obs']};
    static constexpr int FLOPS_ROS_F = {flops['ros_f']};
    static constexpr int FLOPS_ROS_JAC = {flops['ros_jac']}, FLOPS_ROS_LING = {flops['ros_ling']};
    static constexpr int FLOPS_ROS_LINP = {flops['ros_linp']}, FLOPS_ROS_HVP = {flops['ros_hvp']};
    static constexpr int NH = {nh}, FLOPS_IC2_COL = {flops['ic2_col']}, FLOPS_IC2_Q = {flops['ic2_q']};
    static constexpr bool IC2_OK = {'true' if (ic2_possible and n_obs * nF * n_ic <= 64) else 'false'};
    // G column that is d x / d ic_a (its initial value must be the unit vector of state ic_state(a))
    __host__ __device__ static constexpr int ic2_gcol(int a) {{
        constexpr int v[{max(n_ic, 1)}] = {{{arr(ga, '{}')}}};
        return v[a];
    }}
    static constexpr const char* NAME = "{name}";
    __host__ __device__ static constexpr int ic_state(int i) {{
        constexpr int v[{max(n_ic, 1)}] = {{{arr(ic_state, '{}')}}};
        return v[i];
    }}
    static inline const double* default_x0() {{ static const double v[] = {{{arr(x0)}}}; return v; }}
    static inline const double* default_G0() {{ static const double v[] = {{{arr(G0)}}}; return v; }}
    static inline const double* default_theta() {{ static const double v[] = {{{arr(th0)}}}; return v; }}

    __device__ __forceinline__ static void consts(const double* __restrict__ u,
                                                  const double* __restrict__ th,
                                                  double* __restrict__ c) {{
{code['consts']}
    }}
    // x' = f, G' = J G + f_p  (the Fisher quadratures are formed from Q = obs() by the kernel template)
    __device__ __forceinline__ static void primal_z(double t, const double* __restrict__ z,
            const double* __restrict__ u, const double* __restrict__ th,
            const double* __restrict__ c, double* __restrict__ dz) {{
{code['primal_z']}
    }}
    // zb = (d g/d z)^T kb + sum_i w_i d(Q_i L Q_i^T)/dz ;  ub = (d g/d u)^T kb
    // cwL[i*NF + k] = 2 w_i L_k  (L = packed symmetric d criterion / d F; formed once per interval)
    __device__ __forceinline__ static void adjoint(double t, const double* __restrict__ z,
            const double* __restrict__ u, const double* __restrict__ cwL,
            const double* __restrict__ th, const double* __restrict__ c,
            const double* __restrict__ kb,
            double* __restrict__ zb, double* __restrict__ ub) {{
{code['adjoint']}
    }}
    // Q = h_x G, row-major [NOBS][NP]
    __device__ __forceinline__ static void obs(double t, const double* __restrict__ z,
            const double* __restrict__ u, const double* __restrict__ th,
            const double* __restrict__ c, double* __restrict__ Q) {{
{code['obs']}
    }}
    // ---- Rosenbrock pieces; T = double or a forward-mode dual number ------------------------
    static constexpr bool AUTONOMOUS = {'true' if autonomous else 'false'};
    // the hoisted (control, parameter) sub-expressions for a dual-number control
    template <class T>
    __device__ __forceinline__ static void consts_t(const T* __restrict__ u,
                                                    const double* __restrict__ th,
                                                    T* __restrict__ c) {{
{code['consts_t']}
    }}
    // g(y) = [f; J G + f_p] and the unweighted quadrature integrands q_i = (Q_i^T Q_i)_triu
    template <class T>
    __device__ __forceinline__ static void ros_f(double t, const T* __restrict__ z,
            const T* __restrict__ u, const double* __restrict__ th,
            const T* __restrict__ c, T* __restrict__ dz, T* __restrict__ dP) {{
{code['ros_f']}
    }}
    // A = f_x(x), row-major [NX][NX]
    template <class T>
    __device__ __forceinline__ static void ros_jac(double t, const T* __restrict__ z,
            const T* __restrict__ u, const double* __restrict__ th,
            const T* __restrict__ c, T* __restrict__ A) {{
{code['ros_jac']}
    }}
    // BU[j*NX+i] = d(J G_j + f_p,j)_i/dx . U_x   (off-diagonal block of the augmented Jacobian)
    template <class T>
    __device__ __forceinline__ static void ros_ling(double t, const T* __restrict__ z,
            const T* __restrict__ u, const double* __restrict__ th,
            const T* __restrict__ c, const T* __restrict__ U, T* __restrict__ BU) {{
{code['ros_ling']}
    }}
    // dPU = dq/dy . U  (rows of the quadratures)
    template <class T>
    __device__ __forceinline__ static void ros_linp(double t, const T* __restrict__ z,
            const T* __restrict__ u, const double* __restrict__ th,
            const T* __restrict__ c, const T* __restrict__ U, T* __restrict__ dPU) {{
{code['ros_linp']}
    }}
    // out = (d f_x / d eps) U  along the direction (xd, ud) of (states, controls): lets a forward-mode
    // direction be solved with the VALUE factorisation of 1/(h gamma) I - f_x
    __device__ __forceinline__ static void ros_hvp(double t, const double* __restrict__ z,
            const double* __restrict__ xd, const double* __restrict__ u, const double* __restrict__ ud,
            const double* __restrict__ th, const double* __restrict__ c,
            const double* __restrict__ U, double* __restrict__ out) {{
{code['ros_hvp']}
    }}
{ic2_src}}};
"""
    info["sha"] = hashlib.sha256(src.encode()).hexdigest()[:16]
    return src, info


def emit_builtin(out_dir: str):
    """Regenerate ``csrc/generated/*`` for the models compiled into libdynoed_cuda.so."""
    import os
    from . import models
    from .augment import OEDSystem, structural_simplify
    os.makedirs(out_dir, exist_ok=True)
    infos, includes = [], []
    for name, ctor in models.BUILTIN.items():
        sysm = structural_simplify(OEDSystem(ctor()))
        src, info = emit_model(name, sysm)
        with open(os.path.join(out_dir, f"model_{name}.cuh"), "w") as fh:
            fh.write(src)
        infos.append(info)
        includes.append(f'#include "model_{name}.cuh"')
    with open(os.path.join(out_dir, "models.cuh"), "w") as fh:
        fh.write("// GENERATED by dynamicoed.jl_b200/emit.py — do not edit.\n#pragma once\n")
        fh.write("\n".join(includes) + "\n")
        fh.write("#define DYNOED_FOR_EACH_MODEL(X) " +
                 " ".join(f"X(Model_{i['name']})" for i in infos) + "\n")
    with open(os.path.join(out_dir, "models.json"), "w") as fh:
        json.dump(infos, fh, indent=1, ensure_ascii=False)
    return infos


if __name__ == "__main__":
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    infos = emit_builtin(os.path.join(here, "csrc", "generated"))
    for i in infos:
        print(i["name"], i["flops"], file=sys.stderr)
