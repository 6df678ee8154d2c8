"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracles, the committed
golden fixtures and the reference's own goldens.  Tolerance: relative 1e-10 on F entries,
criterion values and gradients (BASELINE.json north_star); gradients are compared relative to
the largest gradient entry of the design (FisherE/D conditioning note)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
RTOL = 1e-10
CRITS = [D.FisherACriterion, D.FisherDCriterion, D.FisherECriterion, D.ACriterion, D.DCriterion,
         D.ECriterion]
CASES = {"one_d": models.one_d, "readme": models.readme, "lv_fishing": models.lotka_volterra_fishing,
         "lv_ic": models.lotka_volterra_ic}
ALGS = {"rk4": D.RK4, "tsit5": D.Tsit5}


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def rel_rows(a, b):
    den = np.maximum(np.max(np.abs(b), axis=1), 1e-300)
    return float(np.max(np.max(np.abs(a - b), axis=1) / den))


def random_designs(name, n_var, n, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    d = rng.random((n, n_var))
    d[:, -1] = 1e-3 + (1 - 1e-3) * d[:, -1]
    if name == "lv_ic":
        d[:, 0] = 0.1 + 0.9 * d[:, 0]
    return d


def make_problem(name, crit_cls, method="rk4", m=4, **kw):
    return D.OEDProblem(D.OEDSystem(CASES[name]()), crit_cls(), alg=ALGS[method](m), **kw)


@pytest.fixture(scope="module")
def oracles():
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    cache = {}

    def get(name, crit, method, m):
        key = (name, method, m)
        if key not in cache:
            o = Oracle(CASES[name](), 0, method, m)
            cache[key] = CppOracle(name, o)
        cache[key].set_criterion(crit)
        return cache[key]
    return get


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("method", ["rk4", "tsit5"])
def test_golden_fixtures(name, method, built_library):
    z = np.load(os.path.join(GOLD, f"{name}_{method}_m4.npz"))
    for crit, cls in enumerate(CRITS):
        ctx = make_problem(name, cls, method).context()
        c, g, F = ctx.evaluate(z["designs"])
        assert rel(c, z[f"crit{crit}"]) < RTOL
        assert rel(F, z["F"]) < RTOL
        assert rel_rows(g, z[f"grad{crit}"]) < RTOL, (name, method, crit)
    X = ctx.trajectory(z["designs"][2])[0].T
    assert np.allclose(X, z["X0"], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("method,m", [("rk4", 4), ("rk4", 1), ("tsit5", 2)])
def test_random_designs_vs_cpp_oracle(name, method, m, oracles, built_library):
    n = 2048 if name == "lv_fishing" else 512
    for crit, cls in enumerate(CRITS):
        ctx = make_problem(name, cls, method, m).context()
        des = random_designs(name, ctx.n_var, n, 20260 + crit)
        c, g, F = ctx.evaluate(des)
        oc, og, oF = oracles(name, crit, method, m).evaluate(des, grad=True)
        assert rel(c, oc) < RTOL and rel(F, oF) < RTOL
        assert rel_rows(g, og) < RTOL, (name, method, m, crit, rel_rows(g, og))
        i, v = ctx.argmin()
        assert i == int(np.argmin(c)) and v == c[i]


@pytest.mark.parametrize("name", ["readme", "lv_fishing", "lv_ic"])
def test_log_det_criterion_extension(name, oracles, built_library):
    """-log det(F + tau I) (criterion id 6; BASELINE north_star's "FisherD logdet F" — not one of the
    reference's six): objective, F and the full gradient against the C++ oracle (np = 1 and np = 2 closed
    forms), the host-side criterion call against numpy, and consistency with the reference's D criterion
    on the same designs: LogD + tau-part == log(DCriterion) where both are defined."""
    ctx = make_problem(name, D.LogDCriterion, "rk4", 4).context()
    des = random_designs(name, ctx.n_var, 600, 77)
    c, g, F = ctx.evaluate(des)
    oc, og, oF = oracles(name, 6, "rk4", 4).evaluate(des, grad=True)
    assert rel(c, oc) < RTOL and rel(F, oF) < RTOL and rel_rows(g, og) < RTOL
    cd = make_problem(name, D.DCriterion, "rk4", 4).context().evaluate(des, grad=False)[0]
    tau = des[:, -1]
    ok = cd - tau > 0
    assert ok.sum() > 500 and np.allclose((c - tau)[ok], np.log((cd - tau)[ok]), rtol=1e-9, atol=1e-12)
    A = np.random.default_rng(5).standard_normal((4, 7))
    M = A @ A.T
    assert np.isclose(D.LogDCriterion()(M, 0.5), -np.linalg.slogdet(M + 0.5 * np.eye(4))[1], rtol=1e-12)
    assert np.allclose(D.LogDCriterion().gradient(M, 0.5), -np.linalg.inv(M + 0.5 * np.eye(4)), rtol=1e-10)


def test_continuous_predict_returns_the_state_after_every_step(built_library):
    """utilities.jl:1-17 (`continuous_predict`): one solution per merged interval.  Here: the state after
    every integrator step.  At the grid points x and G must be BIT-identical to `build_predictor` (same single
    step size; F agrees to rounding), between them it is the same scheme's intermediate state: checked against the closed form of
    the 1-D system (x = exp(p t), G = t exp(p t), test/references/1D.jl:14-24) and, for a graded Rosenbrock
    grid, against the grid-point trajectory to rounding."""
    prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
    des = random_designs("lv_fishing", 71, 1, 3)[0]
    sols = D.continuous_predict(prob, des)
    X, t = D.build_predictor(prob)(des)
    assert len(sols) == 30 and all(len(s_) == 5 and s_.u.shape == (9, 5) for s_ in sols)
    for j, s_ in enumerate(sols):
        assert s_.t[0] == t[j] and s_.t[-1] == t[j + 1] and np.all(np.diff(s_.t) > 0)
        # x and G: the same steps with the same single step size -> bit-identical; F: the quadrature is folded
        # once per (now shorter) interval -> equal to rounding
        assert np.array_equal(s_.u[:6, 0], X[:6, j]) and np.array_equal(s_.u[:6, -1], X[:6, j + 1])
        assert np.allclose(s_.u[6:, -1], X[6:, j + 1], rtol=1e-13, atol=1e-300)
        if j:
            assert np.array_equal(s_.u[:, 0], sols[j - 1].u[:, -1])
    p1 = D.OEDProblem(D.OEDSystem(models.one_d()), D.FisherACriterion(), alg=D.Tsit5(8))
    w = np.array([0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1.0])
    th = p1.context().info
    for s_ in D.continuous_predict(p1, w):
        lam = np.log(s_.u[0, -1] / s_.u[0, 0]) / (s_.t[-1] - s_.t[0])          # x' = p x: recover p from the data
        assert np.allclose(s_.u[0], s_.u[0, 0] * np.exp(lam * (s_.t - s_.t[0])), rtol=1e-9)
        assert np.allclose(s_.u[1], s_.t * s_.u[0], rtol=1e-8, atol=1e-12)      # G = t x
        assert np.all(np.diff(s_.u[2]) >= 0)                                    # F' = w G^2 >= 0
    edges = D.graded_edges(10, 12, 1e-6, 3)
    pr = D.OEDProblem(D.OEDSystem(models.robertson()), D.DCriterion(), alg=D.ROS3P(edges=edges))
    dr = stiff_designs("robertson", pr.context().n_var, 1, 5)[0]
    sr = D.continuous_predict(pr, dr)
    Xr, tr_ = D.build_predictor(pr)(dr)
    assert [len(s_) - 1 for s_ in sr] == [len(e) - 1 for e in edges]
    for j, s_ in enumerate(sr):
        assert rel(s_.u[:, -1], Xr[:, j + 1]) < 1e-11


def test_ragged_and_empty_batches(oracles, built_library):
    ctx = make_problem("lv_fishing", D.DCriterion).context()
    orc = oracles("lv_fishing", 4, "rk4", 4)
    des = random_designs("lv_fishing", 71, 1000, 5)
    full = ctx.evaluate(des)
    # batches up to time_parallel_max run on the time-parallel kernel, larger ones on the direction-parallel
    # kernel: bitwise equal within a kernel whatever the batch size, equal to rounding between the two
    tpmax = ctx.kernel_stats(True)["time_parallel_max"]
    assert 128 <= tpmax < 257
    tp_full = ctx.evaluate(des[:tpmax])
    for a, b in zip(tp_full, full):
        assert rel_rows(a.reshape(tpmax, -1), b[:tpmax].reshape(tpmax, -1)) < 1e-12
    for n in (1, 2, 31, 127, 128, 129, 257, 1000):
        c, g, F = ctx.evaluate(des[:n])
        ref = tp_full if n <= tpmax else full
        assert np.array_equal(c, ref[0][:n]) and np.array_equal(g, ref[1][:n])
        assert np.array_equal(F, ref[2][:n])
    oc, og, oF = orc.evaluate(des[:129], grad=True)
    assert rel(full[0][:129], oc) < RTOL and rel_rows(full[1][:129], og) < RTOL
    # empty batch: OK, nothing written
    ctx.eval_batch(np.zeros((0, 71)), np.zeros(0), np.zeros((0, 71)), None, n=0)
    with pytest.raises(runtime.DynoedError):
        ctx.argmin()
    # objective only / no fim
    c2, g2, f2 = ctx.evaluate(des[:300], grad=False, fim=False)
    assert g2 is None and f2 is None and rel(c2, full[0][:300]) < 1e-13   # another kernel: not bitwise


def test_host_and_device_paths_agree_bitwise(built_library):
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    des = random_designs("lv_fishing", 71, 20000, 6)
    c, g, F = ctx.evaluate(des)                                   # host path (chunked pipeline)
    d = torch.from_numpy(des).cuda()
    dc = torch.empty(len(des), dtype=torch.float64, device="cuda")
    dg = torch.empty_like(d)
    dF = torch.empty((len(des), 3), dtype=torch.float64, device="cuda")
    ctx.eval_batch(d, dc, dg, dF)                                 # device path (one launch)
    assert np.array_equal(dc.cpu().numpy(), c) and np.array_equal(dg.cpu().numpy(), g)
    assert np.array_equal(dF.cpu().numpy(), F)
    assert ctx.argmin()[0] == int(np.argmin(c))
    # unaligned device pointer -> non-TMA staging path, same numbers
    buf = torch.empty(len(des) * 71 + 1, dtype=torch.float64, device="cuda")
    d2 = buf[1:].view(len(des), 71)
    d2.copy_(d)
    ctx.eval_batch(d2, dc, dg, dF)
    assert np.array_equal(dc.cpu().numpy(), c) and np.array_equal(dg.cpu().numpy(), g)
    # mixing host and device pointers is an error, not a crash
    with pytest.raises(runtime.DynoedError):
        ctx.eval_batch(d, c, None, None)


def test_criteria_kat_on_gpu(built_library):
    """/root/reference/test/criteria.jl:14-30 through dynoed_criterion_batch."""
    F = np.array([[5.68145, -0.148312, 2.00382, 5.1841, 3.57814],
                  [-0.148312, 2.17276, -1.18576, 1.57651, -3.95544],
                  [2.00382, -1.18576, 3.8703, -0.14832, 2.69545],
                  [5.1841, 1.57651, -0.14832, 10.9617, 0.0774284],
                  [3.57814, -3.95544, 2.69545, 0.0774284, 24.8495]])
    res = [-47.53571, -2180.4107236837117, -1.0326129713287733, 2.103928493261123,
           0.0004586291881332073, 0.9684170427504832]
    for cls, r in zip(CRITS, res):
        assert np.isclose(cls()(F, 0.0), r, rtol=1.5e-8, atol=0)
        assert np.isclose(cls()(F), r, rtol=1.5e-8, atol=0)
    from oracle.oed_oracle import criterion, criterion_matrix_gradient
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 6, 8, 12):
        A = rng.standard_normal((n, n + 3))
        M = A @ A.T
        for crit, cls in enumerate(CRITS):
            assert np.isclose(cls()(M, 0.25), criterion(crit, M, 0.25), rtol=1e-11)
            L = cls().gradient(M, 0.25)
            Lo = criterion_matrix_gradient(crit, M, 0.25)
            assert np.max(np.abs(L - Lo)) <= 1e-9 * np.max(np.abs(Lo))


def test_full_size_properties(oracles, built_library):
    """BASELINE.json config 3 size (65,536 designs): linearity of F in w, determinism, arg-min,
    and spot parity on a 4,096-design subset."""
    import torch
    n = 65536
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    des = random_designs("lv_fishing", 71, n, 20262)
    d = torch.from_numpy(des).cuda()
    outs = []
    for rep in range(2):
        c = torch.empty(n, dtype=torch.float64, device="cuda")
        g = torch.empty_like(d)
        F = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        ctx.eval_batch(d, c, g, F)
        outs.append((c.cpu().numpy(), g.cpu().numpy(), F.cpu().numpy()))
    assert all(np.array_equal(a, b) for a, b in zip(*outs))           # deterministic
    c, g, F = outs[0]
    assert ctx.argmin() == (int(np.argmin(c)), float(c.min()))
    # linearity: same controls, w3 = 0.25 w1 + 0.5 w2  =>  F3 = 0.25 F1 + 0.5 F2
    half = n // 2
    d1, d2 = des[:half].copy(), des[:half].copy()
    d2[:, 10:70] = des[half:, 10:70]
    d3 = d1.copy()
    d3[:, 10:70] = 0.25 * d1[:, 10:70] + 0.5 * d2[:, 10:70]
    Fs = [ctx.evaluate(x, grad=False)[2] for x in (d1, d2, d3)]
    assert rel(Fs[2], 0.25 * Fs[0] + 0.5 * Fs[1]) < 1e-13
    # d obj / d tau = tr(Lambda) + 1 and FisherD: Lambda = -adj(M) => -(F11 + F22 + 2 tau) + 1
    assert rel(g[:, 70], -(F[:, 0] + F[:, 2] + 2 * des[:, 70]) + 1) < 1e-13
    idx = np.random.default_rng(0).choice(n, 4096, replace=False)
    oc, og, oF = oracles("lv_fishing", 1, "rk4", 4).evaluate(des[idx], grad=True)
    assert rel(c[idx], oc) < RTOL and rel(F[idx], oF) < RTOL and rel_rows(g[idx], og) < RTOL


def test_one_call_at_multistart_size(built_library):
    """BASELINE.json config 5 size in ONE device-path call (1,048,576 + 77 designs, a ragged last tile):
    every row is the value the same design gets in a small batch (bit for bit: one design per lane, nothing
    depends on its position in the batch), the arg-min is the global one (index beyond 2^20), NaN rows lose."""
    import torch
    n = (1 << 20) + 77
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    d = torch.rand((n, 71), dtype=torch.float64, device="cuda", generator=gen)
    d[:, -1] = 1e-3 + 0.99 * d[:, -1]
    d[12345, 3] = float("nan")
    c = torch.empty(n, dtype=torch.float64, device="cuda")
    g = torch.empty_like(d)
    F = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    ctx.eval_batch(d, c, g, F)
    cc = torch.nan_to_num(c, nan=float("inf"))
    i, v = ctx.argmin()
    assert i == int(torch.argmin(cc)) and v == float(cc.min()) and bool(torch.isnan(c[12345]))
    for lo, hi in ((0, 4096), (524288 - 2048, 524288 + 2048), (n - 4096 - 77, n)):     # large enough for the same kernel
        cs, gs, Fs = ctx.evaluate(d[lo:hi].cpu().numpy())
        assert np.array_equal(cs, c[lo:hi].cpu().numpy()) and np.array_equal(gs, g[lo:hi].cpu().numpy())
        assert np.array_equal(Fs, F[lo:hi].cpu().numpy())


def test_set_criterion_and_predictor(oracles, built_library):
    prob = make_problem("lv_fishing", D.FisherACriterion)
    ctx = prob.context()
    des = random_designs("lv_fishing", 71, 64, 8)
    for crit in range(6):
        ctx.set_criterion(crit)
        c, g, _ = ctx.evaluate(des)
        oc, og, _ = oracles("lv_fishing", crit, "rk4", 4).evaluate(des, grad=True)
        assert rel(c, oc) < RTOL and rel_rows(g, og) < RTOL
    X, t = D.build_predictor(prob)(des[0])
    assert X.shape == (9, 31) and np.array_equal(t, prob.timegrid.tpoints)
    Xo = oracles("lv_fishing", 0, "rk4", 4).trajectory(des[0])
    assert np.allclose(X, Xo, rtol=1e-11, atol=1e-13)


def test_user_model_is_emitted_and_compiled_once(built_library):
    """A system that is not compiled in: emitted, built with nvcc, loaded, parity-checked."""
    from dynamicoed.jl_b200.symbolic import (Differential, Eq, ODESystem, independent_variable,
                                             observed_variables, parameters, variables)
    from oracle.oed_oracle import Oracle
    t = independent_variable("t")
    a, b = variables("a b", [1.0, 0.5])
    k = parameters("k[1] k[2]", [0.7, 1.3], tunable=True)
    v = parameters("v", 0.0, input=True, measurement_rate=0.5, bounds=(0.0, 1.0))
    obs = observed_variables("o[1] o[2]", measurement_rate=0.25)
    Dt = Differential(t)
    import sympy as sp
    sysm = ODESystem([Eq(Dt(a), -k[0] * a * b + v), Eq(Dt(b), k[0] * a * b - k[1] * b * sp.exp(-a))],
                     t, tspan=(0.0, 2.0), observed=[Eq(obs[0], a * b), Eq(obs[1], b)], name="user")
    prob = D.OEDProblem(D.OEDSystem(sysm), D.ACriterion(), alg=D.RK4(3))
    ctx = prob.context()
    des = random_designs("user", ctx.n_var, 200, 4)
    c, g, F = ctx.evaluate(des)
    oc, og, oF = Oracle(sysm, 3, "rk4", 3).evaluate(des, grad=True)
    assert rel(c, oc) < RTOL and rel(F, oF) < RTOL and rel_rows(g, og) < RTOL


def _solve(opt_prob, u0=None):
    from scipy.optimize import NonlinearConstraint, minimize
    f = opt_prob.f
    cons = []
    if f.cons is not None:
        cons = [NonlinearConstraint(f.cons, opt_prob.lcons, opt_prob.ucons, jac=f.cons_j)]
    u0 = opt_prob.u0 if u0 is None else u0
    return minimize(f.f, u0, jac=lambda u: f.grad(None, u), bounds=list(zip(opt_prob.lb, opt_prob.ub)),
                    constraints=cons, method="SLSQP", options=dict(ftol=1e-13, maxiter=400))


def test_reference_1d_optimum_through_the_boundary(built_library):
    """/root/reference/test/references/1D.jl:29-66 with SLSQP standing in for Ipopt: objective and
    gradient come from the GPU through OptimizationProblem's f / grad hooks."""
    for i, cls in enumerate(CRITS):
        prob = D.OEDProblem(D.structural_simplify(D.OEDSystem(models.one_d())), cls())
        v = D.states(prob)
        dts = {k: np.array([b - a for a, b in g]) for k, g in D.get_timegrids(prob).items()}
        cons = D.ConstraintsSystem([D.geq(0.2 - 0.5 * sum(dts["w₁"] * v.measurements["w₁"]), 0)])
        op = D.OptimizationProblem(prob, None, constraints=cons, integer_constraints=False)
        u0 = op.u0.copy()
        u0[:10] = 0.3
        r = _solve(op, u0)
        assert np.allclose(r.x[:10], [0, 0, 0, 1, 1, 1, 1, 0, 0, 0], atol=2e-3), (cls, r.x)
        want = -1.2823515846720533e-02 if i < 3 else 2.0
        assert abs(r.fun - want) <= 1e-2 + 1e-1 * abs(want)


def test_reference_lv_optimum_through_the_boundary(built_library):
    """/root/reference/test/references/LotkaVolterra.jl:8-155: u* = [0.62, 0, ...], w1*, w2* = last
    4 of 30, objective -4.85 (atol 1e-2)."""
    prob = D.OEDProblem(D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing())),
                        D.FisherACriterion())
    v = D.states(prob)
    dts = {k: np.array([b - a for a, b in g]) for k, g in D.get_timegrids(prob).items()}
    cons = D.ConstraintsSystem([
        D.geq(0.2 - 0.5 * sum(dts["w₁"] * v.measurements["w₁"]), 0),
        D.geq(0.2 - 0.5 * sum(dts["w₁"] * v.measurements["w₂"]), 0),
        D.geq(1.0, sum(dts["u"] * v.controls["u"]))])
    op = D.OptimizationProblem(prob, None, constraints=cons, integer_constraints=False)
    assert len(op.u0) == 71 and not op.int.any()
    r = _solve(op)
    assert abs(r.fun - (-4.85)) < 1e-2
    assert np.allclose(r.x[:10], [0.62] + [0] * 9, atol=1e-2)
    w_ref = np.array([0.0] * 26 + [1.0] * 4)
    assert np.allclose(r.x[10:40], w_ref, atol=1e-2) and np.allclose(r.x[40:70], w_ref, atol=1e-2)
    opi = D.OptimizationProblem(D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing(True)),
                                             D.DCriterion()), None, integer_constraints=True)
    assert opi.int[:10].all() and opi.int[10:70].all() and not opi.int[70]


def test_overlapped_back_to_back_launches_match_isolated_ones(built_library):
    """Consecutive launches on disjoint buffers that carry DYNOED_OVERLAP_OK may overlap their tails
    (programmatic dependent launch, two scratch sets behind the scratch-set guard).  Every batch must
    still equal its isolated evaluation bit for bit, the arg-min must be the last batch's; launches
    WITHOUT the flag, and launches that reuse a buffer of their predecessor, must not be overlapped."""
    import torch
    from dynamicoed.jl_b200.runtime import OVERLAP_OK
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    ks = ctx.kernel_stats(True)
    n = ks["grid"] * ks["block"] + 5000                  # 1.09 waves: a real partial last wave
    K = 5
    ins = [torch.from_numpy(random_designs("lv_fishing", 71, n, 100 + k)).cuda() for k in range(K)]
    ref = []
    for k in range(K):                                   # isolated: a sync between launches
        c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(ins[k])
        f = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        ctx.eval_batch(ins[k], c, g, f)
        ref.append((c.clone(), g.clone(), f.clone(), ctx.argmin()))
    over0 = ctx.kernel_stats(True)["overlapped_launches"]
    outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty_like(ins[k]),
             torch.empty((n, 3), dtype=torch.float64, device="cuda")) for k in range(K)]
    for k in range(K):                                   # the default is full stream order
        ctx.eval_batch(ins[k], *outs[k], async_=True)
    ctx.sync()
    assert ctx.kernel_stats(True)["overlapped_launches"] == over0
    for rep in range(3):
        for k in range(K):
            ctx.eval_batch(ins[k], *outs[k], async_=True, flags=OVERLAP_OK)
    ctx.sync()
    assert ctx.kernel_stats(True)["overlapped_launches"] - over0 == 3 * K
    for k in range(K):
        for a, b in zip(outs[k], ref[k][:3]):
            assert torch.equal(a, b)
    assert ctx.argmin() == ref[K - 1][3]
    # same output buffers twice in a row: stream order must be kept, so no overlap
    over1 = ctx.kernel_stats(True)["overlapped_launches"]
    ctx.eval_batch(ins[0], *outs[0], async_=True, flags=OVERLAP_OK)
    ctx.eval_batch(ins[1], *outs[0], async_=True, flags=OVERLAP_OK)
    ctx.sync()
    assert ctx.kernel_stats(True)["overlapped_launches"] == over1 + 1   # only the first of the two
    for a, b in zip(outs[0], ref[1][:3]):
        assert torch.equal(a, b)
    # inputs rewritten by a foreign kernel between two launches: without the flag the second launch sees them
    mod = ins[2].clone()
    ctx.eval_batch(ins[0], *outs[0], async_=True, flags=OVERLAP_OK)
    mod.copy_(ins[3])                                    # torch kernel on the same stream
    ctx.eval_batch(mod, *outs[1], async_=True)           # no flag: ordered after the copy
    ctx.sync()
    for a, b in zip(outs[1], ref[3][:3]):
        assert torch.equal(a, b)
    assert ctx.guard_stats()[1] == 0


def test_subwave_launches_overlap_behind_the_scratch_guard(built_library):
    """Batches smaller than one wave (what a branch-and-bound driver sends: 16-32 k designs) also run
    back to back under programmatic dependent launch.  Launch k+2 shares its scratch set with launch k
    and can become resident while k still runs; the scratch-set guard makes its CTAs wait.  Results
    must be bit-identical to isolated launches, no guard wait may time out."""
    import torch
    from dynamicoed.jl_b200.runtime import OVERLAP_OK
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    ks = ctx.kernel_stats(True)
    K = 6
    for n in (ks["grid"] * ks["block"] // 2 + 77, 9000):
        ins = [torch.from_numpy(random_designs("lv_fishing", 71, n, 300 + k)).cuda() for k in range(K)]
        ref = []
        for k in range(K):
            c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(ins[k])
            f = torch.empty((n, 3), dtype=torch.float64, device="cuda")
            ctx.eval_batch(ins[k], c, g, f)
            ref.append((c.clone(), g.clone(), f.clone(), ctx.argmin()))
        outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty_like(ins[k]),
                 torch.empty((n, 3), dtype=torch.float64, device="cuda")) for k in range(K)]
        over0 = ctx.kernel_stats(True)["overlapped_launches"]
        REPS = 16
        for rep in range(REPS):
            for k in range(K):
                ctx.eval_batch(ins[k], *outs[k], async_=True, flags=OVERLAP_OK)
        ctx.sync()
        assert ctx.kernel_stats(True)["overlapped_launches"] - over0 >= REPS * K - 1
        for k in range(K):
            for a, b in zip(outs[k], ref[k][:3]):
                assert torch.equal(a, b)
        # the scratch set's result slot must hold the LAST launch's batch: waiters on a set are served in
        # launch order (a plain lock let a later launch overtake an earlier one here, about one run in four)
        assert ctx.argmin() == ref[K - 1][3]
        for last in (2, 3):                       # ... also when the last launch uses the other scratch set
            for k in range(last + 1):
                ctx.eval_batch(ins[k], *outs[k], async_=True, flags=OVERLAP_OK)
            ctx.sync()
            assert ctx.argmin() == ref[last][3]
    waits, timeouts = ctx.guard_stats()
    assert timeouts == 0
    # objective-only launches take the same path
    n = 20000
    ins = [torch.from_numpy(random_designs("lv_fishing", 71, n, 400 + k)).cuda() for k in range(4)]
    cs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
    iso = []
    for k in range(4):
        ctx.eval_batch(ins[k], cs[k], None, None)
        iso.append(cs[k].clone())
    for rep in range(3):
        for k in range(4):
            ctx.eval_batch(ins[k], cs[k], None, None, async_=True, flags=OVERLAP_OK)
    ctx.sync()
    for k in range(4):
        assert torch.equal(cs[k], iso[k])
    assert ctx.guard_stats()[1] == 0


# ---------------------------------------------------------------------------------------------
# stiff configuration (BASELINE.json config 4): fixed-grid Rosenbrock kernels
# ---------------------------------------------------------------------------------------------
ROS = {"ros2": D.ROS2, "ros3p": D.ROS3P, "rodas3": D.RODAS3}
STIFF = {"robertson": (models.robertson, D.DCriterion, 4), "chem6": (models.chem6, D.FisherECriterion, 2)}


def stiff_designs(name, n_var, n, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    d = rng.random((n, n_var))
    d[:, -1] = 1e-3 + 0.9 * d[:, -1]
    if name == "chem6":                      # unknown initial conditions within their bounds
        d[:, 0] = 0.5 + d[:, 0]
        d[:, 1] = 0.1 + 0.9 * d[:, 1]
    return d


@pytest.mark.parametrize("name", list(STIFF))
@pytest.mark.parametrize("method", ["ros2", "ros3p", "rodas3"])
def test_rosenbrock_kernels_vs_cpp_oracle(name, method, built_library):
    """GPU (block-triangular solves with one nx x nx LU per step, symbolic second derivatives,
    dual parts for the unknown initial conditions) vs the C++ oracle (flat dense Jacobian by nested
    duals): F, objective and the full gradient to relative 1e-10, on a graded grid."""
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    ctor, crit_cls, crit_id = STIFF[name]
    edges = D.graded_edges(10, 30, 1e-6, 6)
    prob = D.OEDProblem(D.OEDSystem(ctor()), crit_cls(), alg=ROS[method](edges=edges))
    ctx = prob.context()
    orc = CppOracle(name, Oracle(ctor(), crit_id, method, list(edges)))
    des = stiff_designs(name, ctx.n_var, 300, 11)
    c, g, F = ctx.evaluate(des)
    oc, og, oF = orc.evaluate(des, grad=True)
    assert rel(F, oF) < RTOL
    # eigenvalue criteria of a small-lambda_min F inherit the ABSOLUTE error of F (|d lambda| <=
    # ||dF||): the objective is compared relative to max(|objective|, ||F||_max) per design
    den = np.maximum(np.abs(oc), np.max(np.abs(oF), axis=1))
    assert float(np.max(np.abs(c - oc) / den)) < RTOL
    assert rel_rows(g, og) < RTOL
    assert ctx.argmin()[0] == int(np.argmin(oc))
    # criterion only, device buffers, predictor at the grid points
    c2, _, _ = ctx.evaluate(des[:77], grad=False, fim=False)     # plain-double kernel (no dual parts)
    assert float(np.max(np.abs(c2 - c[:77]) / den[:77])) < RTOL
    X = ctx.trajectory(des[5])[0]
    assert rel(X.T, orc.trajectory(des[5])) < RTOL


def _pivoting_system():
    """A linear 3-state system dominated by rotation: W = 1/(h gamma) I - f_x has off-diagonal entries
    (40, 25) far above its diagonal (~5) at the step sizes used below, so partial pivoting MUST exchange
    rows at elimination steps 0 and 1."""
    from dynamicoed.jl_b200.symbolic import (Differential, Eq, ODESystem, independent_variable,
                                             observed_variables, parameters, variables)
    t = independent_variable("t")
    x1, x2, x3 = variables("x1 x2 x3", [1.0, 0.5, -0.5])
    k = parameters("k[1] k[2]", [0.7, 1.3], tunable=True)
    obs = observed_variables("o[1] o[2]", measurement_rate=0.25)
    Dt = Differential(t)
    return ODESystem([Eq(Dt(x1), -k[0] * x1 + 40.0 * x2), Eq(Dt(x2), -40.0 * x1 - k[1] * x2 + 25.0 * x3),
                      Eq(Dt(x3), -25.0 * x2 - 0.5 * x3)],
                     t, tspan=(0.0, 1.0), observed=[Eq(obs[0], x1), Eq(obs[1], x3)], name="piv3")


@pytest.mark.parametrize("method", ["ros2", "ros3p"])
def test_rosenbrock_lu_takes_the_row_exchange_path_when_it_must(method, built_library):
    """The Rosenbrock kernels factorise W speculatively without row exchanges and redo the factorisation
    with partial pivoting when some lane of the warp needs an exchange (SmallLU, factor_shifted).  This
    system needs exchanges at every step; F and the objective against the numpy oracle (LAPACK solves of
    the flat dense system), the weight gradient against central differences of the GPU objective."""
    from oracle.oed_oracle import Oracle
    sysm = _pivoting_system()
    gamma = {"ros2": 1.0 + 1.0 / np.sqrt(2.0), "ros3p": 0.7886751345948129}[method]
    W = np.eye(3) / (0.125 * gamma) - np.array([[-0.7, 40.0, 0.0], [-40.0, -1.3, 25.0], [0.0, -25.0, -0.5]])
    assert abs(W[1, 0]) > abs(W[0, 0])                   # partial pivoting exchanges rows 0 and 1
    prob = D.OEDProblem(D.OEDSystem(sysm), D.ACriterion(), alg=ROS[method](2))
    ctx = prob.context()
    rng = np.random.Generator(np.random.Philox(31))
    des = 0.1 + 0.9 * rng.random((70, ctx.n_var))
    c, g, F = ctx.evaluate(des)
    oc, _, oF = Oracle(sysm, D.ACriterion().id, method, 2).evaluate(des, grad=False)
    assert rel(F, oF) < RTOL and rel(c, oc) < RTOL
    big = np.tile(des, (30, 1))                          # 2,100 designs: the large-batch kernel
    cb, gb, Fb = ctx.evaluate(big)
    assert rel(Fb[:70], oF) < RTOL and rel(cb[:70], oc) < RTOL and rel_rows(gb[:70], g) < 1e-11
    eps = 1e-6
    for kk in (0, 3, ctx.n_var - 2, ctx.n_var - 1):
        dp, dm = des[:8].copy(), des[:8].copy()
        dp[:, kk] += eps
        dm[:, kk] -= eps
        fd = (ctx.evaluate(dp, grad=False, fim=False)[0] - ctx.evaluate(dm, grad=False, fim=False)[0]) / (2 * eps)
        assert np.max(np.abs(fd - g[:8, kk]) / np.maximum(np.abs(g[:8, kk]), 1e-3)) < 1e-5, kk


@pytest.mark.parametrize("method", ["ros2", "ros3p", "rodas3"])
def test_second_order_sensitivity_kernel_and_dual_number_kernels_agree(method, monkeypatch, built_library):
    """Unknown initial conditions under a Rosenbrock scheme: the large-batch gradient kernel (ros_ic_kernel:
    emitted second-order sensitivities, columns staged in shared memory), the forward-mode dual-number
    kernel it replaced (DYNOED_NO_IC2=1) and the direction-parallel small-batch kernel are three
    implementations of the same discrete derivative; each is checked against the C++ oracle, for the stiff
    6-state model and for Lotka-Volterra with an unknown initial condition AND an ordinary parameter
    (mixed second-order columns)."""
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    for name, ctor, crit_cls, crit_id, sub in (("chem6", models.chem6, D.FisherECriterion, 2, D.graded_edges(10, 24, 1e-6, 5)),
                                               ("lv_ic", models.lotka_volterra_ic, D.ACriterion, 3, 8)):
        alg = ROS[method](edges=sub) if not isinstance(sub, int) else ROS[method](sub)
        out = {}
        for variant, env in (("ic2", {"DYNOED_NO_DIRPAR": "1"}), ("dual", {"DYNOED_NO_DIRPAR": "1", "DYNOED_NO_IC2": "1"}),
                             ("dirpar", {})):
            for k in ("DYNOED_NO_DIRPAR", "DYNOED_NO_IC2"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            ctx = D.OEDProblem(D.OEDSystem(ctor()), crit_cls(), alg=alg).context()
            des = (stiff_designs if name == "chem6" else random_designs)(name, ctx.n_var, 333, 5)
            st0 = ctx.kernel_stats(True)
            out[variant] = ctx.evaluate(des) + (st0["dynamic_smem"], ctx.kernel_stats(True)["small_batch_launches"])
            ctx.close()
        for k in ("DYNOED_NO_DIRPAR", "DYNOED_NO_IC2"):
            monkeypatch.delenv(k, raising=False)
        assert out["ic2"][3] > out["dual"][3]                # the second-order kernel stages columns in shared memory
        assert out["ic2"][4] == 0 and out["dual"][4] == 0 and out["dirpar"][4] > 0
        orc = CppOracle(name, Oracle(ctor(), crit_id, method, list(sub) if not isinstance(sub, int) else sub))
        oc, og, oF = orc.evaluate(des, grad=True)
        den = np.maximum(np.abs(oc), np.max(np.abs(oF), axis=1))
        for variant, (c, g, F, _, _) in out.items():
            assert rel(F, oF) < RTOL and float(np.max(np.abs(c - oc) / den)) < RTOL, (name, method, variant)
            assert rel_rows(g, og) < RTOL, (name, method, variant, rel_rows(g, og))


def test_robertson_reference_golden_through_the_boundary(built_library):
    """test/references/Rober.jl:50-66 — DCriterion, measurements on the last 3 of 10 intervals:
    objective 0 +- 1e-2 (here 1/(F + tau) + tau with F = 740.6)."""
    edges = D.graded_edges(10, 160, 1e-6, 40)
    prob = D.OEDProblem(D.structural_simplify(D.dae_index_lowering(D.OEDSystem(models.robertson()))),
                        D.DCriterion(), alg=D.ROS3P(edges=edges))
    u = np.asarray(D.get_initial_variables(prob), dtype=np.float64)
    u[7:10] = 1.0
    u[-1] = 1e-3
    c, g, F = prob.context().evaluate(u[None])
    assert abs(F[0, 0] - 740.59) / 740.59 < 1e-4
    assert abs(c[0]) < 1e-2


@pytest.mark.parametrize("method", ["ros2", "rodas3"])
def test_rosenbrock_with_controls_full_gradient(method, built_library):
    """LV-fishing under a Rosenbrock integrator: the control gradient comes from one dual pass per
    control value (no adjoint sweep exists for the Rosenbrock schemes); with DYNOED_GRAD_NO_CONTROLS
    the control entries are NaN and the rest is unchanged."""
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=ROS[method](4))
    ctx = prob.context()
    des = random_designs("lv_fishing", 71, 200, 3)
    c, g, F = ctx.evaluate(des)
    oc, og, oF = CppOracle("lv_fishing", Oracle(models.lotka_volterra_fishing(), 1, method, 4)).evaluate(des, grad=True)
    assert rel(F, oF) < RTOL and rel(c, oc) < RTOL and rel_rows(g, og) < RTOL
    c2, g2, _ = ctx.evaluate(des, flags=runtime.GRAD_NO_CONTROLS)
    assert np.all(np.isnan(g2[:, :10])) and rel(c2, oc) < RTOL and rel_rows(g2[:, 10:], og[:, 10:]) < RTOL
    c3, _, F3 = ctx.evaluate(des, grad=False)
    assert rel(F3, oF) < RTOL and rel(c3, oc) < RTOL


def test_argmin_record_and_gathered_reduction(built_library):
    """The multi-GPU sweep gathers one 16-byte {value, local index} record per rank (NCCL) and
    reduces them with one small kernel; here the ranks' records are fabricated on one GPU."""
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    des = random_designs("lv_fishing", 71, 4096, 21)
    c, _, _ = ctx.evaluate(des, grad=False, fim=False)
    rec = torch.empty(2, dtype=torch.int64, device="cuda")
    ctx.argmin_record_async(rec)
    ctx.sync()
    assert int(rec[1]) == int(np.argmin(c)) and float(rec[:1].view(torch.float64)[0]) == float(np.min(c))
    world, shard = 5, 4096
    vals = np.array([0.5, -2.0, np.nan, -2.0, 1.0])
    idxs = np.array([7, 100, 3, 5, -1])
    gat = torch.from_numpy(np.stack([vals.view(np.int64), idxs], axis=1).copy()).cuda()
    out = torch.empty(3, dtype=torch.int64, device="cuda")
    ctx.argmin_gathered(gat, world, shard, out)
    ctx.sync()
    o = out.cpu().numpy()
    assert o[:1].view(np.float64)[0] == -2.0 and o[1] == 1 * shard + 100 and o[2] == 1   # tie -> smaller global index
    # history ring: the records of the last few batches, oldest first, also across the ring's wrap
    mins = []
    for k in range(70):
        sub = des[37 * k % 3000: 37 * k % 3000 + 512]
        ck, _, _ = ctx.evaluate(sub, grad=False, fim=False)
        mins.append((float(np.min(ck)), int(np.argmin(ck))))
    hist = torch.empty((6, 2), dtype=torch.int64, device="cuda")
    ctx.argmin_history_async(6, hist)
    ctx.sync()
    h = hist.cpu().numpy()
    for row, (v, i) in zip(h, mins[-6:]):
        assert row[:1].view(np.float64)[0] == v and row[1] == i
    # two "ranks" x 6 gathered batches in one reduction
    g2 = torch.stack([hist, hist]).contiguous()
    out2 = torch.empty((6, 3), dtype=torch.int64, device="cuda")
    ctx.argmin_gathered(g2, 2, 512, out2, groups=6)
    ctx.sync()
    o2 = out2.cpu().numpy()
    for row, (v, i) in zip(o2, mins[-6:]):
        assert row[:1].view(np.float64)[0] == v and row[1] == i and row[2] == 0


@pytest.mark.parametrize("name", ["lv_fishing", "readme", "lv_ic"])
def test_information_gain_matches_oracle(name, oracles, built_library):
    """compute_information_gain (/root/reference/src/utilities.jl:47-70) at the merged-grid points:
    P_i(t_j) = Q_i^T Q_i with Q = h_x G from the oracle's trajectory, Pi = F^-1 P F^-1."""
    from oracle.oed_oracle import symmetric_from_vector
    ctx = make_problem(name, D.FisherACriterion).context()
    orc = oracles(name, 0, "rk4", 4)
    o = orc.o                                            # the numpy oracle (for Q = h_x G)
    des = random_designs(name, ctx.n_var, 5, 31)
    F, P, Pi = ctx.information_gain(des)
    nx, np_, nobs, N = o.nxd, o.np_, o.n_obs, o.N
    nc = len(o.controls)
    for d in range(len(des)):
        X = orc.trajectory(des[d])                       # [n_aug, N+1]
        Ff = symmetric_from_vector(X[nx + nx * np_:, -1])
        assert rel(F[d], X[nx + nx * np_:, -1]) < RTOL
        Finv = np.linalg.inv(Ff)
        for j in range(N + 1):
            jj = min(j, N - 1)
            u = [np.array([des[d, o.layout.control_offsets[c] + int(o.grid.indicators[c, jj]) - 1]])
                 for c in range(nc)]
            z = [np.array([v]) for v in X[:, j]]
            _, Q = o.rhs(o.grid.tpoints[jj], z, u, [np.array([0.0])] * nobs)
            Q = np.array([[float(np.asarray(q).reshape(-1)[0]) for q in row] for row in Q])   # [nobs, np]
            for i in range(nobs):
                Pref = np.outer(Q[i], Q[i])
                assert np.max(np.abs(P[d, j, i] - Pref)) <= RTOL * max(np.max(np.abs(Pref)), 1e-300)
                Piref = Finv @ Pref @ Finv
                assert np.max(np.abs(Pi[d, j, i] - Piref)) <= 1e-8 * max(np.max(np.abs(Piref)), 1e-300)


def test_information_gain_host_mirrors(built_library):
    """The reference's names for the analysis helpers (/root/reference/src/utilities.jl:37-70) on top of the GPU
    calls: `extract_fisher`, `compute_local_information_gain`, `compute_information_gain` at the grid points and
    at every integrator step (`dense=True`); the dense result contains the grid-point result."""
    prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherACriterion(), alg=D.RK4(4))
    des = random_designs("lv_fishing", 71, 1, 41)[0]
    F, Ps, Pis, np_, sols = D.compute_information_gain(prob, des)
    assert np_ == 2 and sols is None and len(Ps) == 2 and len(Ps[0]) == 31 and Ps[0][0].shape == (2, 2)
    assert np.array_equal(F, D.extract_fisher(prob, des)) and np.allclose(F, F.T)
    Fi = np.linalg.inv(F)
    for i in range(2):
        for k in (0, 7, 30):
            assert np.allclose(Pis[i][k], Fi @ Ps[i][k] @ Fi, rtol=1e-9, atol=1e-300)
            assert np.linalg.matrix_rank(Ps[i][k], tol=1e-12 * max(np.abs(Ps[i][k]).max(), 1e-300)) <= 1   # Q_i^T Q_i
    X, _ = D.build_predictor(prob)(des)
    Q = [np.array([[X[2, k], X[4, k]], [X[3, k], X[5, k]]]) for k in range(31)]      # h = (x, y): Q = G
    loc = D.compute_local_information_gain(Q)
    assert all(np.allclose(loc[i][k], Ps[i][k], rtol=1e-12, atol=1e-300) for i in range(2) for k in range(31))
    Fd, Pd, Pid, _, sd = D.compute_information_gain(prob, des, dense=True)
    assert len(Pd[0]) == 121 and len(sd) == 30 and np.allclose(Fd, F, rtol=1e-13)
    assert all(np.allclose(Pd[i][4 * k], Ps[i][k], rtol=1e-12, atol=1e-300) for i in range(2) for k in range(31))


def test_fixed_controls_shortcut_matches_full_evaluation(oracles, built_library):
    """Integer / branch-and-bound batch mode (problem.jl:137-146): with the controls frozen, F is
    linear in the weights, so F = designs x basis.  The shortcut must reproduce the full
    integration (objective, F, d/dw, d/dtau) for {0,1} and relaxed measurement designs."""
    import torch
    for crit_cls in (D.FisherDCriterion, D.ACriterion, D.FisherECriterion):
        ctx = make_problem("lv_fishing", crit_cls).context()
        rng = np.random.Generator(np.random.Philox(41))
        base = random_designs("lv_fishing", 71, 1, 5)[0]
        basis = ctx.fim_basis(base)
        assert np.all(basis[:10] == 0.0) and np.all(basis[70] == 0.0) and np.all(np.any(basis[10:70] != 0.0, axis=1))
        n = 3000
        des = np.tile(base, (n, 1))
        des[:, 10:70] = rng.random((n, 60))
        des[: n // 2, 10:70] = (des[: n // 2, 10:70] > 0.7).astype(np.float64)      # integer designs
        des[:, 70] = 1e-3 + 0.9 * rng.random(n)
        c, g, F = ctx.evaluate(des)                                                # full integration
        c2, g2, F2 = np.empty(n), np.empty((n, 71)), np.empty((n, 3))
        ctx.eval_fixed_controls(basis, des, c2, g2, F2)                            # host buffers
        assert rel(F2, F) < RTOL and rel(c2, c) < RTOL
        assert rel_rows(g2[:, 10:], g[:, 10:]) < RTOL and np.all(g2[:, :10] == 0.0)
        d = torch.from_numpy(des).cuda()                                           # device buffers
        dc = torch.empty(n, dtype=torch.float64, device="cuda"); dg = torch.empty_like(d)
        ctx.eval_fixed_controls(torch.from_numpy(basis).cuda(), d, dc, dg, None)
        ctx.sync()
        assert np.array_equal(dc.cpu().numpy(), c2) and np.array_equal(dg.cpu().numpy(), g2)


def test_two_contexts_sharing_a_stream_do_not_overlap(built_library):
    """A launch may only overlap the SAME context's previous launch (the one its buffers were
    checked against): two contexts alternating on one stream must stay strictly ordered."""
    import torch
    a = make_problem("lv_fishing", D.FisherDCriterion).context()
    b = make_problem("lv_fishing", D.FisherACriterion).context()
    st = torch.cuda.current_stream().cuda_stream
    a.set_stream(st); b.set_stream(st)
    ks = a.kernel_stats(True)
    n = ks["grid"] * ks["block"]
    d1 = torch.from_numpy(random_designs("lv_fishing", 71, n, 1)).cuda()
    ga = [torch.empty_like(d1) for _ in range(2)]
    gb = [torch.empty_like(d1) for _ in range(2)]
    ca = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    cb = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    for k in range(4):                                   # rotating buffers: only the stream check can say no
        a.eval_batch(d1, ca[k % 2], ga[k % 2], None, async_=True)
        b.eval_batch(ga[k % 2], cb[k % 2], gb[k % 2], None, async_=True)   # b consumes a's output
    torch.cuda.synchronize()
    assert a.kernel_stats(True)["overlapped_launches"] == 0 and b.kernel_stats(True)["overlapped_launches"] == 0
    cb_ref = torch.empty_like(cb[1]); gb_ref = torch.empty_like(gb[1])
    b.eval_batch(ga[1].clone(), cb_ref, gb_ref, None)
    assert torch.equal(cb[1].nan_to_num(), cb_ref.nan_to_num()) and torch.equal(gb[1].nan_to_num(), gb_ref.nan_to_num())


def test_nan_and_extreme_designs_are_data_not_errors(oracles, built_library):
    """NaN / Inf in a design row poison that row only (no error, no effect on neighbours); NaN never
    wins the arg-min; weights outside [0, 1] and tau = eps() (the reference's lower bound,
    discretize.jl:168) evaluate like any other number and match the oracle."""
    ctx = make_problem("lv_fishing", D.ACriterion).context()
    orc = oracles("lv_fishing", 3, "rk4", 4)
    des = random_designs("lv_fishing", 71, 700, 77)
    clean = ctx.evaluate(des)
    bad = des.copy()
    bad[3, 20] = np.nan            # a weight
    bad[130, 2] = np.inf           # a control
    bad[699, 70] = np.nan          # tau
    c, g, F = ctx.evaluate(bad)
    ok = np.ones(700, dtype=bool); ok[[3, 130, 699]] = False
    assert np.array_equal(c[ok], clean[0][ok]) and np.array_equal(g[ok], clean[1][ok])
    assert np.isnan(c[3]) and not np.isfinite(c[130]) and np.isnan(c[699])
    i, v = ctx.argmin()
    assert ok[i] and v == np.min(c[ok])
    ext = des[:64].copy()
    ext[:, 10:70] = 3.0 * ext[:, 10:70] - 1.0          # weights in [-1, 2)
    ext[:, 70] = np.finfo(float).eps
    ext[0, 10:70] = 0.0                                 # F = 0, M = eps I: 1/eps-sized criteria
    c, g, F = ctx.evaluate(ext)
    oc, og, oF = orc.evaluate(ext, grad=True)
    assert rel(F, oF) < RTOL
    assert float(np.max(np.abs(c - oc) / np.abs(oc))) < 1e-9
    assert rel_rows(g[1:], og[1:]) < 1e-8               # cond(F + eps I) amplifies rounding


@pytest.mark.parametrize("name,method", [("lv_fishing", "rk4"), ("lv_fishing", "tsit5"), ("lv_ic", "rk4"),
                                          ("readme", "tsit5")])
def test_forward_only_gradient_mode(name, method, oracles, built_library):
    """DYNOED_GRAD_NO_CONTROLS: criterion + d/dw + d/dtau (+ d/dx0 by dual parts) from a forward sweep
    with per-interval quadratures; control entries are NaN; everything else equals the adjoint
    gradient / the oracle."""
    ctx = make_problem(name, D.DCriterion, method).context()
    orc = oracles(name, 4, method, 4)
    des = random_designs(name, ctx.n_var, 2500, 13)
    c, g, F = ctx.evaluate(des, flags=runtime.GRAD_NO_CONTROLS)
    oc, og, oF = orc.evaluate(des, grad=True)
    nc = ctx.info.n_ctrl
    lay = make_problem(name, D.DCriterion, method).layout
    ctrl_cols = [o + k for o, l in zip(lay.control_offsets, lay.control_lengths) for k in range(l)] if nc else []
    keep = [k for k in range(ctx.n_var) if k not in ctrl_cols]
    assert rel(F, oF) < RTOL and rel(c, oc) < RTOL
    assert np.all(np.isnan(g[:, ctrl_cols]))
    assert rel_rows(g[:, keep], og[:, keep]) < RTOL
    assert ctx.argmin()[0] == int(np.argmin(oc))


def test_stiff_full_size_properties(built_library):
    """Robertson (config 4) at batch size 65,536 through device buffers: determinism, F linear in
    the weights with the SAME basis for every design (no controls / ICs => x, G do not depend on
    the design), d/dtau = tr(Lambda) + 1, arg-min, and the fixed-controls shortcut reproducing the
    Rosenbrock integration."""
    import torch
    n = 65536
    edges = D.graded_edges(10, 40, 1e-6, 10)
    prob = D.OEDProblem(D.OEDSystem(models.robertson()), D.DCriterion(), alg=D.ROS3P(edges=edges))
    ctx = prob.context()
    des = stiff_designs("robertson", ctx.n_var, n, 5)
    d = torch.from_numpy(des).cuda()
    runs = []
    for _ in range(2):
        c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
        F = torch.empty((n, 1), dtype=torch.float64, device="cuda")
        ctx.eval_batch(d, c, g, F)
        runs.append((c.cpu().numpy(), g.cpu().numpy(), F.cpu().numpy()))
    assert all(np.array_equal(a, b) for a, b in zip(*runs))
    c, g, F = runs[0]
    assert ctx.argmin() == (int(np.argmin(c)), float(c.min()))
    basis = ctx.fim_basis(des[0])                          # [n_var, 1]; the same for every design here
    assert rel(F[:, 0], des @ basis[:, 0]) < 1e-12
    M = F[:, 0] + des[:, 10]
    assert rel(c, 1.0 / M + des[:, 10]) < 1e-13            # DCriterion, np = 1
    assert rel(g[:, 10], -1.0 / M ** 2 + 1.0) < 1e-12
    assert rel_rows(g[:, :10], (-1.0 / M ** 2)[:, None] * basis[None, :10, 0]) < 1e-11
    c2, g2 = np.empty(n), np.empty((n, 11))
    ctx.eval_fixed_controls(basis, des, c2, g2, None)
    assert rel(c2, c) < 1e-12 and rel_rows(g2, g) < 1e-11


def test_pinned_host_buffers_take_the_zero_copy_result_path(built_library):
    """Page-locked result arrays: crit / fim are stored by the kernel straight into host memory;
    numbers must equal the pageable (copy) path bit for bit."""
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    des = random_designs("lv_fishing", 71, 20000, 8)
    c, g, F = ctx.evaluate(des)                                    # pageable numpy arrays
    hp = torch.from_numpy(des).pin_memory()
    hc = torch.full((len(des),), float("nan"), dtype=torch.float64).pin_memory()
    hg = torch.empty((len(des), 71), dtype=torch.float64).pin_memory()
    hf = torch.full((len(des), 3), float("nan"), dtype=torch.float64).pin_memory()
    ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
    assert np.array_equal(hc.numpy(), c) and np.array_equal(hg.numpy(), g) and np.array_equal(hf.numpy(), F)
    assert ctx.argmin()[0] == int(np.argmin(c))


def test_plain_c_consumer_of_the_abi(tmp_path, built_library):
    """examples/c_abi_demo.c — no Python between the caller and libdynoed_cuda.so: the grid and the
    layout are written out by hand in C.  Its numbers must equal the Python host's, and design 0 is
    the reference test's LV optimum (F = [1.2025, 0.3721, 2.6536])."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.dirname(runtime.LIB_PATH)
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
                    "-o", exe, "-L", libdir, "-ldynoed_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    rows = [[float(x) for x in ln.split()[3::2][:1]] + [float(v) for v in ln.split()[5:8]] +
            [float(ln.split()[9]), float(ln.split()[11]), float(ln.split()[13])] for ln in out[:4]]
    des = np.zeros((4, 71))
    des[0, 0] = 0.62; des[0, 36:40] = 1.0; des[0, 66:70] = 1.0; des[0, 70] = 1.0
    for j in range(1, 4):
        des[j] = [((j * 37 + k * 11) % 17) / 17.0 for k in range(71)]
        des[j, 70] = 0.1 * j
    c, g, F = make_problem("lv_fishing", D.FisherDCriterion).context().evaluate(des)
    for j, r in enumerate(rows):
        assert r[0] == c[j] and r[1:4] == list(F[j]) and r[4:] == [g[j, 0], g[j, 39], g[j, 70]]
    assert np.allclose(F[0], [1.20249794, 0.37211035, 2.6535698], rtol=2e-6)     # RK4 m=4 vs exact: 6e-8
    assert out[4].split()[1] == str(int(np.argmin(c)))


@pytest.mark.parametrize("crit_cls", CRITS)
def test_four_parameter_model_exercises_the_register_jacobi_path(crit_cls, built_library):
    """np = 4 (three rate constants + one unknown initial condition): the in-kernel criteria use the
    cyclic-Jacobi eigen-decomposition in registers instead of the np <= 2 closed forms.  All six
    criteria and their gradients (controls, IC, weights, tau) against the numpy oracle, for the
    adjoint kernel and the forward-only kernel."""
    from dynamicoed.jl_b200.symbolic import (Differential, Eq, ODESystem, independent_variable,
                                             observed_variables, parameters, variables)
    from oracle.oed_oracle import Oracle
    t = independent_variable("t")
    a = variables("a", 1.0, tunable=True, bounds=(0.5, 1.5))           # unknown IC on the leading state
    b, c3 = variables("b c", [0.5, 0.2])
    k = parameters("k[1] k[2] k[3]", [0.7, 1.3, 0.4], tunable=True)
    v = parameters("v", 0.0, input=True, measurement_rate=0.5, bounds=(0.0, 1.0))
    obs = observed_variables("o[1] o[2] o[3]", measurement_rate=0.25)
    Dt = Differential(t)
    sysm = ODESystem([Eq(Dt(a), -k[0] * a * b + v), Eq(Dt(b), k[0] * a * b - k[1] * b * c3),
                      Eq(Dt(c3), k[1] * b * c3 - k[2] * c3 + 0.1 * a)],
                     t, tspan=(0.0, 2.0), observed=[Eq(obs[0], a), Eq(obs[1], b * c3), Eq(obs[2], c3)], name="np4")
    prob = D.OEDProblem(D.OEDSystem(sysm), crit_cls(), alg=D.RK4(2))
    ctx = prob.context()
    assert ctx.info.np == 4 and ctx.info.nF == 10 and ctx.info.n_ic == 1
    rng = np.random.Generator(np.random.Philox(9))
    des = rng.random((160, ctx.n_var))
    des[:, -1] = 1e-2 + 0.9 * des[:, -1]
    des[:, prob.layout.ic_offset] = 0.5 + des[:, prob.layout.ic_offset]
    des[:, prob.layout.w_offsets[0]:prob.layout.tau_offset] = 0.2 + 0.8 * des[:, prob.layout.w_offsets[0]:prob.layout.tau_offset]
    c, g, F = ctx.evaluate(des)
    oc, og, oF = Oracle(sysm, crit_cls().id, "rk4", 2).evaluate(des, grad=True)
    den = np.maximum(np.abs(oc), 1e-300)
    assert rel(F, oF) < RTOL and float(np.max(np.abs(c - oc) / den)) < 1e-9
    assert rel_rows(g, og) < 1e-8                       # eigen-criteria: cond(F + tau I) amplifies rounding
    c2, g2, _ = ctx.evaluate(des, flags=runtime.GRAD_NO_CONTROLS)
    lay = prob.layout
    keep = [k_ for k_ in range(ctx.n_var) if not (lay.control_offsets[0] <= k_ < lay.control_offsets[0] + lay.control_lengths[0])]
    assert float(np.max(np.abs(c2 - oc) / den)) < 1e-9 and rel_rows(g2[:, keep], og[:, keep]) < 1e-8


@pytest.mark.parametrize("crit_cls", CRITS)
def test_reference_integer_optimum_by_exhaustive_batch(crit_cls, built_library):
    """test/references/1D.jl:68-107 ("Integer": Juniper + Ipopt): x' = p x, 10 measurement decisions,
    0.05 sum(w) <= 0.2.  The batch evaluator makes branch-and-bound unnecessary at this size: ALL
    feasible {0,1}^10 designs are evaluated in one call (full integration and the fixed-controls
    shortcut) and the arg-min is the reference's optimum w* = [0,0,0,1,1,1,1,0,0,0], objective
    -1.2823515846720533e-2 (Fisher*) or 2.0 (A/D/E) within the reference's tolerances."""
    import itertools
    prob = D.OEDProblem(D.OEDSystem(models.one_d()), crit_cls(), alg=D.Tsit5(8))
    ctx = prob.context()
    W = np.array([w for w in itertools.product((0.0, 1.0), repeat=10) if 0.05 * sum(w) <= 0.2 + 1e-12])
    des = np.concatenate([W, np.ones((len(W), 1))], axis=1)          # tau at its upper bound
    c, g, F = ctx.evaluate(des)
    best = int(np.argmin(c))
    assert ctx.argmin()[0] == best
    assert np.array_equal(W[best], [0, 0, 0, 1, 1, 1, 1, 0, 0, 0])
    golden = -1.2823515846720533e-2 if crit_cls().id <= 2 else 2.0
    assert np.isclose(c[best], golden, atol=1e-2, rtol=1e-1)
    assert abs(F[best, 0] - 0.012813544228654607) < 1e-9              # exact int_0.3^0.7 t^2 e^{-4t} dt
    basis = ctx.fim_basis(des[0])                                     # no controls: one basis for all
    c2 = np.empty(len(W))
    ctx.eval_fixed_controls(basis, des, c2, None, None)
    assert int(np.argmin(c2)) == best and rel(c2, c) < 1e-12


def test_long_design_vectors_pick_a_smaller_tile(oracles, built_library):
    """README model with measurement_rate 1/400: n_var = 401 (> 128 rows x 227 KB of shared memory),
    so the default tile drops below 128; results still match the oracle, TMA staging included."""
    from dynamicoed.jl_b200.symbolic import (Differential, Eq, ODESystem, independent_variable,
                                             observed_variables, parameters, variables)
    from oracle.oed_oracle import Oracle
    t = independent_variable("t")
    x = variables("x", 1.0)
    p = parameters("p[1]", -2.0, tunable=True)
    obs = observed_variables("obs", measurement_rate=400)
    sysm = ODESystem([Eq(Differential(t)(x), p * x)], t, tspan=(0.0, 1.0), observed=[Eq(obs, x ** 2)],
                     name="simple_system")
    prob = D.OEDProblem(D.OEDSystem(sysm), D.ACriterion(), alg=D.RK4(1))
    ctx = prob.context()
    assert ctx.n_var == 401 and ctx.kernel_stats(True)["block"] == 64
    rng = np.random.Generator(np.random.Philox(3))
    des = rng.random((333, 401)); des[:, -1] = 0.05 + 0.9 * des[:, -1]
    c, g, F = ctx.evaluate(des)
    oc, og, oF = Oracle(sysm, 3, "rk4", 1).evaluate(des, grad=True)
    assert rel(c, oc) < RTOL and rel(F, oF) < RTOL and rel_rows(g, og) < RTOL


def test_readme_example_runs_end_to_end(built_library):
    """examples/readme_example.py — the reference's README (README.md:25-63) with D.solve standing in
    for Ipopt: at most 3 of 10 measurements of x^2; FisherA puts them where (2 x G)^2 is largest."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "readme_example.py")],
                         capture_output=True, text=True, check=True, timeout=300).stdout
    assert "success True" in out
    w = np.array([float(v) for v in out.split("w*  = [")[1].split("]")[0].split()])
    assert abs(w.sum() - 3.0) < 1e-6 and np.all(w > -1e-9) and np.all(w < 1 + 1e-9)
    # the integrand 4 t^2 e^{-8t} peaks at t = 0.25: intervals 2, 3, 4 (0-based 1..3) carry the weight
    assert np.allclose(np.sort(np.argsort(-w)[:3]), [1, 2, 3])


@pytest.mark.parametrize("name,method,crit_cls", [("robertson", "ros3p", D.DCriterion), ("chem6", "rodas3", D.FisherECriterion)])
def test_stiff_golden_fixtures(name, method, crit_cls, built_library):
    """Committed fixtures of the stiff configuration (tests/golden/make_golden.py::main_stiff): no
    oracle at run time."""
    z = np.load(os.path.join(GOLD, f"{name}_{method}_graded.npz"))
    edges = D.graded_edges(10, 30, 1e-6, 6)
    ctx = D.OEDProblem(D.OEDSystem(STIFF[name][0]()), crit_cls(), alg=ROS[method](edges=edges)).context()
    c, g, F = ctx.evaluate(z["designs"])
    den = np.maximum(np.abs(z["crit"]), np.max(np.abs(z["F"]), axis=1))
    assert rel(F, z["F"]) < RTOL and float(np.max(np.abs(c - z["crit"]) / den)) < RTOL
    assert rel_rows(g, z["grad"]) < RTOL


def test_pipelined_host_calls_match_blocking_ones(built_library):
    """dynoed_eval_batch_host_async / dynoed_wait: a single host thread keeps two host-buffer calls in
    flight (page-locked arrays); every call's results equal the blocking call's bit for bit, in any
    interleaving with device-path launches."""
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    n, K = 30000, 6
    des = [random_designs("lv_fishing", 71, n, 50 + k) for k in range(K)]
    ref = [ctx.evaluate(d) for d in des]
    sets = [(torch.empty((n, 71), dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
             torch.empty((n, 71), dtype=torch.float64).pin_memory(), torch.empty((n, 3), dtype=torch.float64).pin_memory())
            for _ in range(2)]
    ids = [None, None]
    for k in range(K + 1):
        b = k % 2
        if k >= 2 or (k >= 1 and False):
            pass
        if ids[b] is not None:                       # the call that used this buffer set two steps ago
            ctx.wait(ids[b])
            hp, hc, hg, hf = sets[b]
            c, g, F = ref[k - 2]
            assert np.array_equal(hc.numpy(), c) and np.array_equal(hg.numpy(), g) and np.array_equal(hf.numpy(), F)
        if k < K:
            hp, hc, hg, hf = sets[b]
            hp.numpy()[:] = des[k]
            hc.fill_(float("nan")); hg.fill_(float("nan"))
            ids[b] = ctx.eval_batch_host_async(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
            if k == 3:                               # a device-path launch in between must not break ordering
                d = torch.from_numpy(des[0]).cuda(); dc = torch.empty(n, dtype=torch.float64, device="cuda")
                ctx.eval_batch(d, dc, None, None)
                assert np.array_equal(dc.cpu().numpy(), ref[0][0])
    ctx.wait(ids[(K - 1) % 2])
    hp, hc, hg, hf = sets[(K - 1) % 2]
    assert np.array_equal(hc.numpy(), ref[K - 1][0]) and np.array_equal(hg.numpy(), ref[K - 1][1])
    with pytest.raises(runtime.DynoedError):
        ctx.wait(10 ** 6)


# ---- Hessians (dynoed_hessian_batch): what Ipopt's exact-Hessian mode asks Optimization.jl for ----
def _fd_hessian_of(gradfun, x, rel_step):
    """The same central-difference formula, on the CPU, from any gradient function."""
    nv = x.size
    h = rel_step * np.maximum(1.0, np.abs(x))
    P = np.vstack([x + np.diag(h), x - np.diag(h)])
    G = gradfun(P)
    den = (P[np.arange(nv), np.arange(nv)] - P[nv + np.arange(nv), np.arange(nv)])[:, None]
    A = (G[:nv] - G[nv:]) / den
    return 0.5 * (A + A.T)


def _phi_second(crit, M, A, B):
    """d^2 phi(M)[A, B] for the criteria that are smooth functions of det / inverse."""
    Mi = np.linalg.inv(M)
    tA, tB, tAB = np.trace(Mi @ A), np.trace(Mi @ B), np.trace(Mi @ A @ Mi @ B)
    d = np.linalg.det(M)
    if crit == 0:
        return 0.0                                    # -tr M
    if crit == 1:
        return -d * (tA * tB - tAB)                   # -det M
    if crit == 3:
        return 2.0 * np.trace(Mi @ A @ Mi @ B @ Mi)   # tr M^-1
    if crit == 4:
        return (tA * tB + tAB) / d                    # 1 / det M
    raise ValueError(crit)


@pytest.mark.parametrize("crit", [0, 1, 3, 4])
def test_hessian_weight_block_vs_closed_form(crit, oracles, built_library):
    # F is linear in the weights (augment.jl:223-226): the (w, tau) block of the Hessian is
    # d^2 phi(F + tau I)[B_a, B_b] with B_k = dF/dw_k from the ORACLE (tau direction: I)
    ctx = make_problem("lv_fishing", CRITS[crit]).context()
    orc = oracles("lv_fishing", crit, "rk4", 4)
    x = random_designs("lv_fishing", 71, 1, 77)[0]
    x[-1] = 0.3
    c, g, H = ctx.hessian(x)
    oc, og, oF = orc.evaluate(x[None], grad=True)
    assert rel(np.array([c]), oc) < RTOL and rel_rows(g[None], og) < RTOL
    assert np.array_equal(H, H.T)
    wcols = np.arange(10, 70)
    units = np.tile(x, (60, 1)); units[:, 10:70] = np.eye(60)
    _, _, Bp = orc.evaluate(units, grad=False)
    sym = lambda p: np.array([[p[0], p[1]], [p[1], p[2]]])
    dirs = [sym(b) for b in Bp] + [np.eye(2)]
    M = sym(oF[0]) + x[-1] * np.eye(2)
    cols = list(wcols) + [70]
    Hw = np.array([[_phi_second(crit, M, A, B) for B in dirs] for A in dirs])
    scale = max(np.abs(Hw).max(), np.abs(og).max())
    assert np.abs(H[np.ix_(cols, cols)] - Hw).max() < 2e-8 * scale, np.abs(H[np.ix_(cols, cols)] - Hw).max() / scale


@pytest.mark.parametrize("name,crit", [("lv_fishing", 1), ("lv_fishing", 5), ("lv_ic", 3), ("readme", 3)])
def test_hessian_full_vs_oracle_differences(name, crit, oracles, built_library):
    ctx = make_problem(name, CRITS[crit]).context()
    orc = oracles(name, crit, "rk4", 4)
    xs = random_designs(name, ctx.n_var, 3, 91 + crit)
    rs = 6.0554544523933395e-06
    c, g, H = ctx.hessian(xs)                                   # three points: one batch of 3 (2 nv + 1) rows
    oc, og, _ = orc.evaluate(xs, grad=True)
    assert rel(c, oc) < RTOL and rel_rows(g, og) < RTOL
    for k in range(3):
        Ho = _fd_hessian_of(lambda P: orc.evaluate(P, grad=True)[1], xs[k], rs)
        scale = max(np.abs(Ho).max(), np.abs(og[k]).max())
        assert np.abs(H[k] - Ho).max() < 1e-7 * scale, (name, crit, np.abs(H[k] - Ho).max() / scale)
        assert np.array_equal(H[k], H[k].T)
    # truncation error: a 10x larger step moves the result by O(h^2)
    _, _, H10 = ctx.hessian(xs[0], rel_step=10 * rs)
    assert np.abs(H10 - H[0]).max() < 1e-5 * max(np.abs(H[0]).max(), 1.0)


def test_hessian_device_pointers_chunks_and_errors(built_library):
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    nv = ctx.n_var
    n = 1000                                                     # > 131072 / 143 = 916: two chunks
    xs = random_designs("lv_fishing", nv, n, 5)
    c, g, H = ctx.hessian(xs)
    d = torch.from_numpy(xs).cuda()
    Hd = torch.empty((n, nv, nv), dtype=torch.float64, device="cuda")
    cd = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.hessian_batch(d, Hd, cd, None)
    assert np.array_equal(Hd.cpu().numpy(), H) and np.array_equal(cd.cpu().numpy(), c)
    # one point alone is a small batch (direction-parallel kernel): same numbers up to rounding
    c1, g1, H1 = ctx.hessian(xs[917])
    assert rel(H1, H[917]) < 1e-8 and rel(g1, g[917]) < 1e-12 and abs(c1 - c[917]) < 1e-13 * abs(c1)
    c0, g0, _ = ctx.evaluate(xs)
    assert rel(c0, c) < 1e-13 and rel_rows(g0, g) < 1e-12
    with pytest.raises(runtime.DynoedError):
        ctx.hessian_batch(xs, Hd)                                # host designs, device Hessian
    with pytest.raises(runtime.DynoedError):
        ctx.hessian(xs[0], rel_step=0.5)


def test_exact_hessian_solve_reaches_the_reference_optimum(built_library):
    # /root/reference/test/references/LotkaVolterra.jl:8-155 solved the way the reference's tests do
    # — `solve(opt_prob, Ipopt.Optimizer())`, i.e. with exact Hessians — here scipy trust-constr
    # with the GPU Hessian in Ipopt's place
    prob = D.OEDProblem(D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing())),
                        D.FisherACriterion())
    v = D.states(prob)
    dts = {k: np.array([b - a for a, b in g]) for k, g in D.get_timegrids(prob).items()}
    cons = D.ConstraintsSystem([
        D.geq(0.2 - 0.5 * sum(dts["w₁"] * v.measurements["w₁"]), 0),
        D.geq(0.2 - 0.5 * sum(dts["w₁"] * v.measurements["w₂"]), 0),
        D.geq(1.0, sum(dts["u"] * v.controls["u"]))])
    op = D.OptimizationProblem(prob, None, constraints=cons, integer_constraints=False)
    r = D.solve(op, method="trust-constr", gtol=1e-9, xtol=1e-12, maxiter=500)
    assert abs(r.objective - (-4.85)) < 1e-2, r.objective
    assert np.allclose(r.u[:10], [0.62] + [0] * 9, atol=1e-2), r.u[:10]
    w_ref = np.array([0.0] * 26 + [1.0] * 4)
    assert np.allclose(r.u[10:40], w_ref, atol=1e-2) and np.allclose(r.u[40:70], w_ref, atol=1e-2)
    H = op.f.hess(None, r.u)
    assert H.shape == (71, 71) and np.array_equal(H, H.T) and np.abs(H[:10, :10]).max() > 0


# ---- small batches: the direction-parallel kernel (optimiser callback, Hessian rows) ----
@pytest.mark.parametrize("name,method,m", [("lv_fishing", "rk4", 4), ("lv_fishing", "tsit5", 2), ("lv_ic", "rk4", 4),
                                           ("lv_ic", "tsit5", 4)])
def test_small_batches_use_the_direction_parallel_kernel_and_match(name, method, m, oracles, built_library):
    for crit in (1, 2, 3):
        ctx = make_problem(name, CRITS[crit], method, m).context()
        ks = ctx.kernel_stats(True)
        nmax = ks["small_batch_max"]
        assert nmax >= 64
        big = random_designs(name, ctx.n_var, nmax + 300, 400 + crit)      # too large: adjoint kernel
        cb, gb, Fb = ctx.evaluate(big)
        assert ctx.kernel_stats(True)["small_batch_launches"] == ks["small_batch_launches"]
        for n in (1, 7, 64, nmax):
            before = ctx.kernel_stats(True)["small_batch_launches"]
            c, g, F = ctx.evaluate(big[:n])
            assert ctx.kernel_stats(True)["small_batch_launches"] > before
            assert rel(c, cb[:n]) < 1e-12 and rel(F, Fb[:n]) < 1e-12 and rel_rows(g, gb[:n]) < 1e-11
            i, v = ctx.argmin()
            assert i == int(np.argmin(c)) and v == c[i]
        oc, og, oF = oracles(name, crit, method, m).evaluate(big[:64], grad=True)
        c, g, F = ctx.evaluate(big[:64])
        assert rel(c, oc) < RTOL and rel(F, oF) < RTOL and rel_rows(g, og) < RTOL


@pytest.mark.parametrize("name,method,sub", [("lv_fishing", "rk4", 4), ("lv_fishing", "tsit5", 2), ("lv_ic", "rk4", 4),
                                             ("lv_ic", "ros3p", 8), ("lv_fishing", "ros2", 4),
                                             ("chem6", "rodas3", D.graded_edges(10, 8, 1e-6, 3)),
                                             ("chem6", "ros3p", D.graded_edges(10, 30, 1e-6, 6))])   # tables spill to global scratch
def test_time_parallel_kernel_matches_the_other_gradient_kernels_and_the_oracle(name, method, sub, built_library, monkeypatch):
    """The smallest gradient batches run one CTA per design (tp_eval_kernel: x-only sweep, per-step affine
    maps, chunked chains).  Same numbers as the direction-parallel kernel (DYNOED_NO_TIMEPAR), the large-batch
    kernels and the C++ oracle; arg-min and NaN handling included."""
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    ctor = {"lv_fishing": models.lotka_volterra_fishing, "lv_ic": models.lotka_volterra_ic, "chem6": models.chem6}[name]
    algs = dict(ALGS, **ROS)
    alg = algs[method](edges=sub) if not isinstance(sub, int) else algs[method](sub)
    crit_cls, crit_id = (D.FisherECriterion, 2) if name == "chem6" else (D.ACriterion, 3)
    out = {}
    for variant, env in (("tp", {}), ("tp_spill", {"DYNOED_TIMEPAR_SPILL": "1"}), ("dir", {"DYNOED_NO_TIMEPAR": "1"}),
                         ("large", {"DYNOED_NO_DIRPAR": "1"})):
        for k in ("DYNOED_NO_TIMEPAR", "DYNOED_NO_DIRPAR", "DYNOED_TIMEPAR_SPILL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = D.OEDProblem(D.OEDSystem(ctor()), crit_cls(), alg=alg).context()
        ks = ctx.kernel_stats(True)
        assert (ks["time_parallel_max"] > 0) == variant.startswith("tp"), (variant, ks)
        des = (stiff_designs if name == "chem6" else random_designs)(name, ctx.n_var, 40, 21)
        res = []
        for n in (1, 5, 40):
            c, g, F = ctx.evaluate(des[:n])
            i, v = ctx.argmin()
            assert i == int(np.argmin(c)) and v == c[i]
            res.append((c, g, F))
        small = ctx.kernel_stats(True)["small_batch_launches"]
        assert (small == 0) == (variant == "large")
        if variant.startswith("tp"):              # objective-only calls of this size: same kernel without its gradient phases
            c0, g0, F0 = ctx.evaluate(des[:5], grad=False)
            assert g0 is None and rel(c0, res[1][0]) < 1e-12 and rel(F0, res[1][2]) < 1e-12
            i0, v0 = ctx.argmin()
            assert i0 == int(np.argmin(c0)) and v0 == c0[i0]
            if variant == "tp" and not (name == "chem6" and method == "ros3p"):     # (not with spilled tables)
                assert ctx.kernel_stats(True)["small_batch_launches"] == small + 1
                assert np.array_equal(c0, res[1][0]) and np.array_equal(F0, res[1][2])
        if variant == "tp":                       # a NaN design poisons its own row only and never wins
            bad = des[:5].copy(); bad[2, 3] = np.nan
            c, g, F = ctx.evaluate(bad)
            ok = np.array([0, 1, 3, 4])
            assert np.isnan(c[2]) and np.array_equal(c[ok], res[1][0][ok]) and np.array_equal(g[ok], res[1][1][ok])
            assert ctx.argmin()[0] == int(ok[np.argmin(c[ok])])
        out[variant] = res
        ctx.close()
    for k in ("DYNOED_NO_TIMEPAR", "DYNOED_NO_DIRPAR", "DYNOED_TIMEPAR_SPILL"):
        monkeypatch.delenv(k, raising=False)
    for (c, g, F), (c2, g2, F2) in zip(out["tp"], out["tp_spill"]):   # same algorithm, tables in global scratch (another
        assert rel(c, c2) < 1e-12 and rel(F, F2) < 1e-12 and rel_rows(g, g2) < 1e-11   # instantiation: FMA contraction may differ)
    orc = CppOracle(name, Oracle(ctor(), crit_id, method, list(sub) if not isinstance(sub, int) else sub))
    oc, og, oF = orc.evaluate(des, grad=True)
    den = np.maximum(np.abs(oc), np.max(np.abs(oF), axis=1))
    for variant, res in out.items():
        for (c, g, F), n in zip(res, (1, 5, 40)):
            assert rel(F, oF[:n]) < RTOL and float(np.max(np.abs(c - oc[:n]) / den[:n])) < RTOL, (variant, n)
            if not (variant == "large" and name == "lv_fishing" and method == "ros2"):   # large Rosenbrock batches: controls need DYNOED_GRAD controls mode
                assert rel_rows(g, og[:n]) < RTOL, (variant, n)
    for (c, g, F), (c2, g2, F2) in zip(out["tp"], out["dir"]):
        assert rel(F, F2) < 1e-11 and rel_rows(g, g2) < 1e-10


def test_small_batch_kernel_device_pointers_and_nan(built_library):
    import torch
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    des = random_designs("lv_fishing", 71, 200, 3)
    des[5, 12] = np.nan
    c, g, F = ctx.evaluate(des)
    d = torch.from_numpy(des).cuda()
    dc = torch.empty(200, dtype=torch.float64, device="cuda"); dg = torch.empty_like(d)
    dF = torch.empty((200, 3), dtype=torch.float64, device="cuda")
    ctx.eval_batch(d, dc, dg, dF)
    ctx.sync()
    ok = np.ones(200, dtype=bool); ok[5] = False
    assert np.array_equal(dc.cpu().numpy()[ok], c[ok]) and np.array_equal(dg.cpu().numpy()[ok], g[ok])
    assert np.isnan(c[5]) and np.isnan(dc.cpu().numpy()[5])
    i, v = ctx.argmin()
    assert i == int(np.nanargmin(c)) and v == c[i]
    # the stiff configuration: Rosenbrock small batches, gradient incl. unknown initial conditions
    prob = D.OEDProblem(D.OEDSystem(models.chem6()), D.FisherECriterion(),
                        alg=D.RODAS3(edges=D.graded_edges(10, 30, 1e-6, 6)))
    cx = prob.context()
    nmax = cx.kernel_stats(True)["small_batch_max"]
    dd = stiff_designs("chem6", cx.n_var, nmax + 50, 9)
    cb, gb, Fb = cx.evaluate(dd)                                   # too large: one dual pass per direction
    assert cx.kernel_stats(True)["small_batch_launches"] == 0
    c, g, F = cx.evaluate(dd[:32])
    assert cx.kernel_stats(True)["small_batch_launches"] >= 1
    den = np.maximum(np.abs(cb[:32]), np.max(np.abs(Fb[:32]), axis=1))
    assert float(np.max(np.abs(c - cb[:32]) / den)) < 1e-11 and rel(F, Fb[:32]) < 1e-11
    assert rel_rows(g, gb[:32]) < RTOL


def test_multistart_sweep_device_side_selection(built_library):
    """Chunked back-to-back launches, per-batch records reduced on the device by the library's
    pack / select kernels, one host read: the winner equals the arg-min of all criterion values and
    carries its own design / gradient rows."""
    import torch
    from dynamicoed.jl_b200 import multistart
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    n = 3 * 65536 + 1234
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    des = torch.rand((n, 71), dtype=torch.float64, device="cuda", generator=g)
    des[:, -1] = 1e-3 + 0.99 * des[:, -1]
    des[17, 3] = float("nan")
    crit = torch.empty(n, dtype=torch.float64, device="cuda"); grad = torch.empty_like(des)
    l0 = ctx.kernel_stats(True)["launches"]
    out = multistart.sweep_device(ctx, des, crit, grad, 1000, chunk=65536)
    assert ctx.kernel_stats(True)["launches"] - l0 == 4 + 2          # 4 evaluation launches + pack + select
    k = int(torch.argmin(torch.nan_to_num(crit, nan=float("inf"))))
    assert out["index"] == 1000 + k and out["value"] == float(crit[k]) and out["owner"] == 0
    assert torch.equal(out["design"], des[k].cpu()) and torch.equal(out["grad"], grad[k].cpu())
    # a second sweep over designs that a torch kernel rewrote in place: ordered, not stale
    des.mul_(0.5).add_(0.25)
    out2 = multistart.sweep_device(ctx, des, crit, grad, 0, chunk=65536)
    k2 = int(torch.argmin(torch.nan_to_num(crit, nan=float("inf"))))
    assert out2["index"] == k2 and out2["value"] == float(crit[k2])
    c_chk = torch.empty(64, dtype=torch.float64, device="cuda")
    ctx.eval_batch(des[k2:k2 + 64].contiguous(), c_chk, None, None)
    assert abs(float(c_chk[0]) - out2["value"]) <= 1e-12 * max(1.0, abs(out2["value"]))   # (64 designs: the time-parallel kernel)
    ctx.own_stream()


def test_pack_and_select_winner_kernels_match_their_host_mirror(built_library):
    """dynoed_pack_winner / dynoed_select_winner against multistart.pack_local_winner / gather_winner (the
    plain-torch specification the gloo tests run): NaN never wins, ties go to the smallest global index,
    empty and all-NaN shards, the ring's wrap-around, fabricated multi-rank gathers."""
    import torch
    from dynamicoed.jl_b200 import multistart
    ctx = make_problem("lv_fishing", D.FisherDCriterion).context()
    nv, chunk = 71, 512
    L = 2 + 2 * nv
    for nb in (1, 3, 64, 5):                                  # 64 + 5 batches wrap the 64-deep ring
        n = nb * chunk - (37 if nb > 1 else 0)
        des = torch.from_numpy(random_designs("lv_fishing", nv, n, 900 + nb)).cuda()
        if nb == 3:
            des[chunk + 5, 2] = float("nan")
        crit = torch.empty(n, dtype=torch.float64, device="cuda"); grad = torch.empty_like(des)
        for c in range(nb):
            a, b = c * chunk, min((c + 1) * chunk, n)
            ctx.eval_batch(des[a:b], crit[a:b], grad[a:b], None, async_=True)
        pay = torch.empty(L, dtype=torch.float64, device="cuda")
        ctx.pack_winner(nb, des, grad, chunk, 10_000, pay)
        ctx.sync()
        vals = torch.stack([torch.nan_to_num(crit[c * chunk:(c + 1) * chunk], nan=float("inf")).min() for c in range(nb)])
        idxs = torch.stack([torch.argmin(torch.nan_to_num(crit[c * chunk:(c + 1) * chunk], nan=float("inf"))) + c * chunk
                            for c in range(nb)])
        want = multistart.pack_local_winner(vals, idxs, 10_000, des, grad)
        assert torch.equal(pay, want)
    # empty shard
    pay0 = torch.empty(L, dtype=torch.float64, device="cuda")
    ctx.pack_winner(0, des[:0], grad[:0], chunk, 5, pay0, n_local=0)
    ctx.sync()
    assert float(pay0[0]) == float("inf") and float(pay0[1]) == -1.0 and float(pay0[2:].abs().sum()) == 0.0
    # fabricated gathers: world = 5 with a NaN rank, an empty rank and a tie
    g = torch.zeros((5, L), dtype=torch.float64, device="cuda")
    g[:, 2:] = torch.arange(5, dtype=torch.float64, device="cuda")[:, None] + 0.5
    g[:, 0] = torch.tensor([0.5, -2.0, float("nan"), -2.0, float("inf")], dtype=torch.float64)
    g[:, 1] = torch.tensor([7.0, 9000.0, 3.0, 4105.0, -1.0], dtype=torch.float64)
    out = torch.empty(L + 1, dtype=torch.float64, device="cuda")
    ctx.select_winner(g, 5, out)
    ctx.sync()
    assert float(out[-1]) == 3.0 and float(out[0]) == -2.0 and float(out[1]) == 4105.0   # tie -> smaller global index
    assert torch.equal(out[2:-1], g[3, 2:])
    g[:, 1] = -1.0
    ctx.select_winner(g, 5, out)
    ctx.sync()
    assert float(out[-1]) == -1.0 and float(out[0]) == float("inf") and float(out[1]) == -1.0
    # selection on an auxiliary stream: ordered after the evaluations by an event, results identical
    side = torch.cuda.Stream()
    ctx.set_aux_stream(side.cuda_stream)
    pay2 = torch.empty(L, dtype=torch.float64, device="cuda")
    ctx.pack_winner(5, des, grad, chunk, 10_000, pay2)
    side.synchronize()
    assert torch.equal(pay2, pay)
    ctx.set_aux_stream(0)
    ctx.own_stream()


def test_repeated_smallest_eigenvalue_gives_a_valid_subgradient(built_library):
    """FisherE / E at a REPEATED lambda_min (SURVEY.md section 7 hard part): the criterion is not
    differentiable there, the value is still exact and the kernel must return one valid element of the
    generalised gradient, -v v^T (FisherE) or -v v^T / lambda^2 (E) with v a unit vector of the smallest
    eigenvalue's eigenspace (closed 2x2 form: v = e_1; register Jacobi: the first minimal column)."""
    from oracle.oed_oracle import criterion, criterion_matrix_gradient
    rng = np.random.default_rng(11)
    mats = [np.diag([2.0, 2.0]), 3.5 * np.eye(2), np.diag([2.0, 2.0, 5.0]), np.diag([7.0, 1.5, 1.5]),
            np.diag([4.0, 4.0, 4.0, 9.0]), 0.75 * np.eye(4)]
    for n, k in ((3, 2), (4, 2), (4, 3), (6, 2)):            # rotated: eigenspace not axis-aligned
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.concatenate([np.full(k, 1.25), 1.25 + 1.0 + rng.random(n - k)])
        mats.append((Q * lam) @ Q.T)
    for Mx in mats:
        n = Mx.shape[0]
        Mx = 0.5 * (Mx + Mx.T)
        lmin = float(np.linalg.eigvalsh(Mx)[0])
        for cls, crit in ((D.FisherECriterion, 2), (D.ECriterion, 5)):
            v = cls()(Mx, 0.0)
            assert np.isclose(v, criterion(crit, Mx, 0.0), rtol=1e-12)
            L = cls().gradient(Mx, 0.0)
            scale = 1.0 if crit == 2 else 1.0 / lmin ** 2
            assert np.allclose(L, L.T, atol=1e-14)
            P = -L / scale                                    # should be v v^T, |v| = 1, M v = lmin v
            assert abs(np.trace(P) - 1.0) < 1e-9
            assert np.linalg.norm(P @ P - P) < 1e-8           # a rank-one projector
            assert np.linalg.norm(Mx @ P - lmin * P) < 1e-8 * max(1.0, abs(lmin))
            if np.count_nonzero(Mx - np.diag(np.diag(Mx))) == 0:
                # axis-aligned eigenspace: the documented convention coincides with LAPACK's (first column)
                assert np.allclose(L, criterion_matrix_gradient(crit, Mx, 0.0), atol=1e-13)


def test_lv_against_the_oracle_with_its_own_constants(built_library):
    """The C++ oracle evaluated with the Lotka-Volterra / 1D constants it carries itself (written down in
    oracle/oed_oracle.cpp from /root/reference/test/references/LotkaVolterra.jl:11-17 and 1D.jl:15-16),
    not the ones the product's models.py feeds it: a wrong constant in models.py can no longer cancel."""
    from oracle.cpp_oracle import CppOracle
    from oracle.oed_oracle import Oracle
    for name, crit_cls, crit in (("lv_fishing", D.FisherDCriterion, 1), ("one_d", D.FisherACriterion, 0)):
        ctx = make_problem(name, crit_cls).context()
        des = random_designs(name, ctx.n_var, 777, 91)
        c, g, F = ctx.evaluate(des)
        orc = CppOracle(name, Oracle(CASES[name](), crit, "rk4", 4), own_constants=True)
        orc.set_criterion(crit)
        oc, og, oF = orc.evaluate(des, grad=True)
        assert rel(c, oc) < RTOL and rel(F, oF) < RTOL and rel_rows(g, og) < RTOL
