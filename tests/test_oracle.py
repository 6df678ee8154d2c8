"""Pins the CPU oracles: against the reference's own goldens (tight: test/criteria.jl; loose:
test/references/*.jl optima), against independent tight-tolerance scipy integration, and against
each other (numpy/sympy oracle with symbolic Jacobians vs C++ dual-number oracle with hand-written
Jacobians)."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp
from scipy.optimize import minimize

from dynamicoed.jl_b200 import models
from oracle import cpp_oracle
from oracle.oed_oracle import (CRIT_A, CRIT_D, CRIT_E, FISHER_A, FISHER_D, FISHER_E, Oracle,
                               criterion, criterion_matrix_gradient, symmetric_from_vector,
                               tableau, vector_from_symmetric)

# /root/reference/test/criteria.jl:15-24
F_KAT = np.array([[5.68145, -0.148312, 2.00382, 5.1841, 3.57814],
                  [-0.148312, 2.17276, -1.18576, 1.57651, -3.95544],
                  [2.00382, -1.18576, 3.8703, -0.14832, 2.69545],
                  [5.1841, 1.57651, -0.14832, 10.9617, 0.0774284],
                  [3.57814, -3.95544, 2.69545, 0.0774284, 24.8495]])
KAT = [(FISHER_A, -47.53571), (FISHER_D, -2180.4107236837117), (FISHER_E, -1.0326129713287733),
       (CRIT_A, 2.103928493261123), (CRIT_D, 0.0004586291881332073), (CRIT_E, 0.9684170427504832)]


def test_packing_roundtrip():   # test/criteria.jl:5-12
    rng = np.random.default_rng(1)
    for n in range(1, 21):
        A = rng.integers(0, 1000, (n, n)) / 10.0
        A = np.triu(A) + np.triu(A, 1).T
        v = vector_from_symmetric(A)
        assert v.shape == (n * (n + 1) // 2,)
        assert np.array_equal(symmetric_from_vector(v), A)
        # column-major upper triangle: x[j(j-1)/2 + i] (1-based)
        for j in range(1, n + 1):
            for i in range(1, j + 1):
                assert v[j * (j - 1) // 2 + i - 1] == A[i - 1, j - 1]


@pytest.mark.parametrize("crit,res", KAT)
def test_criteria_kat(crit, res):   # test/criteria.jl:26-29, isapprox default rtol = sqrt(eps)
    assert np.isclose(criterion(crit, F_KAT, 0.0), res, rtol=1.5e-8, atol=0)
    assert np.isclose(criterion(crit, F_KAT), res, rtol=1.5e-8, atol=0)
    v = cpp_oracle.lib().oracle_criterion(crit, 5, vector_from_symmetric(F_KAT).ctypes.data, 0.0, None)
    assert np.isclose(v, res, rtol=1.5e-8, atol=0)


def test_log_det_extension_against_the_reference_d_criteria():
    """-log det(F + tau I) (DYNOED_LOG_D, an extension: BASELINE's north_star names "FisherD logdet F",
    the reference has -det F and 1/det F): consistent with the reference's own D values on the KAT matrix of
    test/criteria.jl, in both oracles."""
    from oracle.oed_oracle import LOG_D
    fisher_d, crit_d = dict(KAT)[1], dict(KAT)[4]
    v = criterion(LOG_D, F_KAT)
    assert np.isclose(v, -np.log(-fisher_d), rtol=1e-9) and np.isclose(v, np.log(crit_d), rtol=1e-9)
    vc = cpp_oracle.lib().oracle_criterion(LOG_D, 5, vector_from_symmetric(F_KAT).ctypes.data, 0.0, None)
    assert np.isclose(vc, v, rtol=1e-12)
    Lc = np.empty(15)
    cpp_oracle.lib().oracle_criterion(LOG_D, 5, vector_from_symmetric(F_KAT).ctypes.data, 0.25, Lc.ctypes.data)
    assert np.allclose(symmetric_from_vector(Lc), criterion_matrix_gradient(LOG_D, F_KAT, 0.25), rtol=1e-10, atol=0)


@pytest.mark.parametrize("crit", range(7))
def test_criterion_gradient_vs_finite_differences(crit):
    rng = np.random.default_rng(crit)
    for n in (1, 2, 3, 5):
        A = rng.standard_normal((n, n + 2))
        F = A @ A.T
        tau = 0.3
        L = criterion_matrix_gradient(crit, F, tau)
        eps = 1e-6
        for i in range(n):
            for j in range(i, n):
                E = np.zeros((n, n)); E[i, j] = E[j, i] = eps
                fd = (criterion(crit, F + E, tau) - criterion(crit, F - E, tau)) / (2 * eps)
                an = L[i, j] * (1 if i == j else 2)
                assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (crit, n, i, j, fd, an)
        fd = (criterion(crit, F, tau + eps) - criterion(crit, F, tau - eps)) / (2 * eps)
        assert abs(fd - np.trace(L)) <= 1e-6 * max(1.0, abs(fd))


def _order_conditions(A, b, c, order):
    s = len(b)
    Am = np.zeros((s, s))
    for i, row in enumerate(A):
        Am[i, :len(row)] = row
    b, c = np.array(b), np.array(c)
    assert np.allclose(Am.sum(1), c, atol=1e-14)
    conds = [(b.sum(), 1), (b @ c, 1 / 2), (b @ c ** 2, 1 / 3), (b @ Am @ c, 1 / 6),
             (b @ c ** 3, 1 / 4), (b @ (c * (Am @ c)), 1 / 8), (b @ Am @ c ** 2, 1 / 12),
             (b @ Am @ Am @ c, 1 / 24)]
    if order >= 5:
        conds += [(b @ c ** 4, 1 / 5), (b @ Am @ c ** 3, 1 / 20), (b @ Am @ Am @ Am @ c, 1 / 120),
                  (b @ (c ** 2 * (Am @ c)), 1 / 10), (b @ ((Am @ c) ** 2), 1 / 20)]
    for got, want in conds:
        assert abs(got - want) < 2e-14, (got, want)


def test_tableaus_satisfy_order_conditions():
    _order_conditions(*tableau("rk4"), order=4)
    _order_conditions(*tableau("tsit5"), order=5)   # pins the Tsit5 coefficients


LV_OPT = np.zeros(71)
LV_OPT[0] = 0.62
LV_OPT[36:40] = 1.0
LV_OPT[66:70] = 1.0
LV_OPT[70] = 1.0


def _lv_exact_F(design):
    """Independent tight-tolerance integration (scipy DOP853) of the LV augmented system."""
    p1, p2, p3, p4, c1, c2 = 1.0, 1.0, 0.4, 0.6, 1.0, 1.0

    def rhs(t, z, u, w1, w2):
        x, y, g11, g21, g12, g22 = z[:6]
        J = np.array([[p1 - c1 * y - p3 * u, -c1 * x], [c2 * y, -p2 + c2 * x - p4 * u]])
        G = np.array([[g11, g12], [g21, g22]])
        dG = J @ G + np.array([[-x * y, 0], [0, x * y]])
        return [p1 * x - c1 * x * y - p3 * x * u, -p2 * y + c2 * x * y - p4 * y * u,
                dG[0, 0], dG[1, 0], dG[0, 1], dG[1, 1],
                w1 * g11 ** 2 + w2 * g21 ** 2, w1 * g11 * g12 + w2 * g21 * g22,
                w1 * g12 ** 2 + w2 * g22 ** 2]
    z = np.array([0.5, 0.7, 0, 0, 0, 0, 0, 0, 0], dtype=float)
    for j in range(30):
        u, w1, w2 = design[j // 3], design[10 + j], design[40 + j]
        sol = solve_ivp(rhs, (j / 10, (j + 1) / 10), z, method="DOP853", rtol=1e-12, atol=1e-14,
                        args=(u, w1, w2))
        z = sol.y[:, -1]
    return z[6:]


def test_lv_oracle_vs_scipy_and_reference_golden():
    """RK4 convergence to the exact solution (4th order) and the reference's loose golden:
    objective -4.85 (atol 1e-2) at the reference's optimum, LotkaVolterra.jl:71-154."""
    exact = _lv_exact_F(LV_OPT)
    errs = []
    for m in (1, 2, 4, 8):
        o = Oracle(models.lotka_volterra_fishing(), FISHER_A, "rk4", m)
        obj, _, F = o.evaluate(LV_OPT)
        errs.append(np.max(np.abs(F[0] - exact) / np.abs(exact)))
    assert errs[2] < 1e-6 and errs[3] < 1e-7
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(3)]
    assert all(3.5 < r < 4.6 for r in rates), rates
    o = Oracle(models.lotka_volterra_fishing(), FISHER_A, "tsit5", 4)
    obj, _, F = o.evaluate(LV_OPT)
    assert np.max(np.abs(F[0] - exact) / np.abs(exact)) < 1e-8
    assert abs(obj[0] - (-4.85)) < 1e-2                     # the reference's golden
    assert abs(obj[0] - (-(exact[0] + exact[2]) - 2 + 1)) < 1e-7   # -tr(F + I) + 1


def test_lv_integer_golden():
    """LotkaVolterra.jl:222-302: u = [1,0,...], w = last 4, DCriterion objective 0.34 (atol 1e-2)."""
    d = LV_OPT.copy()
    d[0] = 1.0
    d[70] = 0.0     # tau* ~ 0 in the reference's solution (res.u; u_opt adds u0's 1.0)
    o = Oracle(models.lotka_volterra_fishing(), CRIT_D, "rk4", 8)
    obj, _, _ = o.evaluate(d)
    assert abs(obj[0] - 0.34) < 1e-2


def test_lv_kat_final_state():
    """SURVEY.md Appendix B KAT (u = 0, w = 1, RK4 m=4)."""
    d = np.zeros(71); d[10:70] = 1.0; d[70] = 1.0
    o = Oracle(models.lotka_volterra_fishing(), FISHER_A, "rk4", 4)
    z, _, _ = o.integrate(d)
    want = [1.8757549575397647, 0.9235735726919355, -0.8056043454731636, -1.8434637962149398,
            -2.7405713888119, 1.1629300556336966, 3.0000579151345703, 0.9889221531612189,
            4.194031331196515]
    assert np.allclose([float(v[0]) for v in z], want, rtol=1e-13, atol=0)


def test_one_d_exact_solution():
    """x = e^{pt}, G = t e^{pt}; w = 1 on [0.3, 0.7]: F = int t^2 e^{-4t} = 0.012813544228654607
    (the reference golden -1.28235e-2 at 1D.jl:62 carries Tsit5 reltol-1e-3 error)."""
    d = np.zeros(11); d[3:7] = 1.0; d[10] = 1.0
    o = Oracle(models.one_d(), FISHER_A, "rk4", 16)
    obj, g, F = o.evaluate(d, grad=True)
    assert abs(F[0, 0] - 0.012813544228654607) < 1e-9
    assert abs(obj[0] - (-1.2823515846720533e-2)) < 1e-2 * 1.0     # reference tolerance
    o = Oracle(models.readme(), FISHER_A, "rk4", 16)
    d = np.ones(11)
    assert abs(o.evaluate(d)[2][0, 0] - 0.015410094253999954) < 1e-9


def test_one_d_optimum_matches_reference():
    """Solve the reference's relaxed 1-D problem (1D.jl:29-66) with the oracle's objective and
    gradient: w* = [0,0,0,1,1,1,1,0,0,0] (atol 1e-5 / rtol 1e-1 in the reference)."""
    for crit, want in [(FISHER_A, -1.2823515846720533e-2), (CRIT_D, 2.0)]:
        o = Oracle(models.one_d(), crit, "tsit5", 4)
        fun = lambda u: float(o.evaluate(u)[0][0])
        jac = lambda u: o.evaluate(u, grad=True)[1][0]
        cons = [{"type": "ineq", "fun": lambda u: 0.2 - 0.5 * 0.1 * np.sum(u[:10]),
                 "jac": lambda u: np.concatenate([-0.05 * np.ones(10), [0.0]])}]
        u0 = np.concatenate([np.full(10, 0.3), [1.0]])
        r = minimize(fun, u0, jac=jac, bounds=[(0, 1)] * 10 + [(np.finfo(float).eps, 1.0)],
                     constraints=cons, method="SLSQP", options=dict(ftol=1e-14, maxiter=300))
        assert np.allclose(r.x[:10], [0, 0, 0, 1, 1, 1, 1, 0, 0, 0], atol=2e-3), r.x
        assert abs(r.fun - want) < 1e-2 + 1e-1 * abs(want)


def test_lv_ic_golden():
    """LotkaVolterra.jl:357-373: x0* = 0.15, w* = last 2 of 10, objective ~ 1e-2 (atol 1e-2)."""
    d = np.zeros(12); d[0] = 0.15; d[9:11] = 1.0; d[11] = 0.0
    o = Oracle(models.lotka_volterra_ic(), CRIT_D, "rk4", 16)
    obj, _, F = o.evaluate(d)
    assert abs(obj[0] - 1e-2) < 1e-2
    assert np.allclose(F[0], [9.4102, -11.5351, 87.7490], rtol=1e-4)


def test_linearity_in_weights():
    """Fact 5: under a fixed grid F(t_f) is exactly linear in w."""
    rng = np.random.default_rng(5)
    o = Oracle(models.lotka_volterra_fishing(), FISHER_D, "rk4", 4)
    d1 = rng.random(71); d2 = d1.copy(); d2[10:70] = rng.random(60); d3 = d1.copy()
    d3[10:70] = 0.3 * d1[10:70] + 1.7 * d2[10:70]
    F = o.evaluate(np.stack([d1, d2, d3]))[2]
    assert np.allclose(F[2], 0.3 * F[0] + 1.7 * F[1], rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("name,ctor", [("lv_fishing", models.lotka_volterra_fishing),
                                       ("one_d", models.one_d), ("readme", models.readme),
                                       ("lv_ic", models.lotka_volterra_ic)])
@pytest.mark.parametrize("method", ["rk4", "tsit5"])
def test_cpp_oracle_matches_numpy_oracle(name, ctor, method):
    rng = np.random.default_rng(7)
    for crit in range(6):
        o = Oracle(ctor(), crit, method, 3)
        c = cpp_oracle.CppOracle(name, o)
        des = rng.random((5, o.n_var))
        des[:, -1] = 1e-3 + (1 - 1e-3) * des[:, -1]
        a = o.evaluate(des, grad=True)
        b = c.evaluate(des, grad=True)
        for x, y in zip(a, b):
            assert np.max(np.abs(x - y)) <= 1e-12 * np.max(np.abs(x))
    X = c.trajectory(des[0])
    Y = o.trajectory(des[0])
    assert np.allclose(X, Y, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name,ctor", [("lv_fishing", models.lotka_volterra_fishing), ("one_d", models.one_d)])
def test_models_py_constants_equal_the_constants_the_oracle_carries_itself(name, ctor):
    """oracle/oed_oracle.cpp holds its own copy of the LV / 1D constants, written from
    /root/reference/test/references/LotkaVolterra.jl:11-17 and 1D.jl:15-16.  Feeding the oracle the product's
    models.py values must change nothing, bit for bit: otherwise one of the two copies is wrong."""
    rng = np.random.default_rng(17)
    o = Oracle(ctor(), 1, "rk4", 4)
    fed, own = cpp_oracle.CppOracle(name, o), cpp_oracle.CppOracle(name, o, own_constants=True)
    des = rng.random((9, o.n_var))
    for x, y in zip(fed.evaluate(des, grad=True), own.evaluate(des, grad=True)):
        assert np.array_equal(x, y)


def test_forward_mode_gradient_vs_finite_differences():
    rng = np.random.default_rng(11)
    o = Oracle(models.lotka_volterra_fishing(), CRIT_A, "rk4", 2)
    d = rng.random(71)
    _, g, _ = o.evaluate(d, grad=True)
    for k in (0, 4, 9, 12, 45, 70):
        e = np.zeros(71); e[k] = 1e-6
        fd = (o.evaluate(d + e)[0][0] - o.evaluate(d - e)[0][0]) / 2e-6
        assert abs(fd - g[0, k]) < 1e-7 * max(1, abs(fd))


# ---------------------------------------------------------------------------------------------
# stiff configuration: fixed-grid Rosenbrock oracles
# ---------------------------------------------------------------------------------------------
def _graded(n_first=40, m=10, n_intervals=10):
    import dynamicoed.jl_b200 as D
    return list(D.graded_edges(n_intervals, n_first, 1e-6, m))


def test_rosenbrock_tableaus_order_and_stability():
    """Empirical order of ROS2 / ROS3P on a nonlinear test problem and |R(-inf)| on y' = lambda y,
    through the numpy oracle's step (so the coefficients themselves are checked)."""
    from oracle import rosenbrock_oracle as R

    class Toy:   # y1' = -y1 y2 + 0.3 y2^2, y2' = y1^2 - y2
        def __init__(self, method): self.method = method
        def rhs(self, t, y, u, w): return [-y[0] * y[1] + 0.3 * y[1] * y[1], y[0] * y[0] - y[1]], None
    f = lambda t, y: [-y[0] * y[1] + 0.3 * y[1] ** 2, y[0] ** 2 - y[1]]
    ref = solve_ivp(f, (0, 2), [1.0, 0.5], rtol=1e-13, atol=1e-14, method="DOP853").y[:, -1]
    for method, order in (("ros2", 2), ("ros3p", 3), ("rodas3", 3)):
        errs = []
        for n in (40, 80, 160):
            y = [np.array([1.0]), np.array([0.5])]
            for _ in range(n):
                y = R.step(Toy(method), 0.0, y, 2.0 / n, [], [])
            errs.append(np.hypot(y[0][0] - ref[0], y[1][0] - ref[1]))
        rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
        assert np.all(np.abs(rates - order) < 0.15), (method, rates)

    class Lin:
        def __init__(self, method): self.method = method
        def rhs(self, t, y, u, w): return [-1e8 * y[0]], None
    assert abs(R.step(Lin("ros2"), 0.0, [np.array([1.0])], 1.0, [], [])[0][0]) < 1e-6      # L-stable
    assert abs(R.step(Lin("ros3p"), 0.0, [np.array([1.0])], 1.0, [], [])[0][0]) < 0.74     # A-stable
    assert abs(R.step(Lin("rodas3"), 0.0, [np.array([1.0])], 1.0, [], [])[0][0]) < 1e-6    # L-stable


def test_robertson_rosenbrock_vs_radau_and_reference_golden():
    """Robertson DAE (test/references/Rober.jl:9-21), w = last 3 of 10: tight scipy Radau gives
    F = 740.59; the fixed-grid Rosenbrock oracles converge to it and the
    DCriterion objective matches the reference's golden 0 +- 1e-2 (Rober.jl:50-66)."""
    k1, k2, k3 = 0.04, 3e7, 1e4

    def rhs(t, s):
        y1, y2, g1, g2, F = s
        y3 = 1 - y1 - y2
        a11, a12 = -k1 - k3 * y2, k3 * y3 - k3 * y2
        a21, a22 = k1 + k3 * y2, -k3 * y3 + k3 * y2 - 2 * k2 * y2
        return [-k1 * y1 + k3 * y2 * y3, k1 * y1 - k3 * y2 * y3 - k2 * y2 ** 2,
                a11 * g1 + a12 * g2 - y1, a21 * g1 + a22 * g2 + y1, (t >= 70.0) * g1 * g1]
    s1 = solve_ivp(rhs, (0, 70), [1, 0, 0, 0, 0], method="Radau", rtol=1e-11, atol=1e-14)
    s2 = solve_ivp(rhs, (70, 100), s1.y[:, -1], method="Radau", rtol=1e-11, atol=1e-14)
    F_exact = s2.y[4, -1]
    assert abs(F_exact - 740.59) < 0.05
    des = np.zeros((1, 11)); des[0, 7:10] = 1.0; des[0, 10] = 1e-3
    errs = {}
    for method, n0, m in (("ros2", 160, 40), ("ros3p", 40, 10), ("ros3p", 160, 40), ("rodas3", 40, 10)):
        o = Oracle(models.robertson(), CRIT_D, method, _graded(n0, m))
        c, _, F = o.evaluate(des)
        errs[(method, n0)] = abs(F[0, 0] - F_exact) / F_exact
        assert abs(c[0] - (1.0 / (F_exact + 1e-3) + 1e-3)) < 1e-2          # reference golden tolerance
    assert errs[("ros2", 160)] < 2e-3 and errs[("ros3p", 40)] < 2e-3 and errs[("ros3p", 160)] < 1e-4
    assert errs[("rodas3", 40)] < 2e-3


@pytest.mark.parametrize("name,ctor,crit", [("robertson", models.robertson, CRIT_D),
                                            ("chem6", models.chem6, FISHER_E)])
@pytest.mark.parametrize("method", ["ros2", "ros3p", "rodas3"])
def test_rosenbrock_cpp_oracle_matches_numpy_oracle_and_fd(name, ctor, crit, method):
    """Flat dense Rosenbrock two ways (numpy: dual-number Jacobian + LAPACK; C++: nested duals +
    own pivoted LU) agree to rounding; the C++ forward-mode gradient matches central differences."""
    o = Oracle(ctor(), crit, method, _graded(30, 6))
    co = cpp_oracle.CppOracle(name, o)
    rng = np.random.Generator(np.random.Philox(7))
    des = rng.random((3, o.n_var))
    des[:, -1] = 1e-3 + 0.9 * des[:, -1]
    if name == "chem6":
        des[:, 0] = 0.5 + des[:, 0]
        des[:, 1] = 0.1 + 0.9 * des[:, 1]
    c, _, F = o.evaluate(des)
    c2, g2, F2 = co.evaluate(des, grad=True)
    assert np.max(np.abs(F - F2) / np.max(np.abs(F2))) < 1e-11
    assert np.max(np.abs(c - c2) / np.abs(c2)) < 1e-11
    eps = 1e-6
    for k in range(o.n_var):
        dp, dm = des.copy(), des.copy()
        dp[:, k] += eps; dm[:, k] -= eps
        fd = (co.evaluate(dp, grad=False)[0] - co.evaluate(dm, grad=False)[0]) / (2 * eps)
        assert np.allclose(fd, g2[:, k], rtol=2e-5, atol=1e-9 * max(1.0, np.max(np.abs(g2)))), (k, fd, g2[:, k])


@pytest.mark.parametrize("name,method", [("robertson", "ros3p"), ("chem6", "rodas3")])
def test_stiff_golden_fixtures_match_the_cpp_oracle(name, method):
    """The committed stiff fixtures are reproducible from the oracle (guards against drift)."""
    import os
    import dynamicoed.jl_b200 as D
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}_{method}_graded.npz"))
    ctor = {"robertson": models.robertson, "chem6": models.chem6}[name]
    o = Oracle(ctor(), int(z["criterion"]), method, list(D.graded_edges(10, 30, 1e-6, 6)))
    c, g, F = cpp_oracle.CppOracle(name, o).evaluate(z["designs"], grad=True)
    assert np.array_equal(c, z["crit"]) and np.array_equal(g, z["grad"]) and np.array_equal(F, z["F"])


@pytest.mark.parametrize("crit", [1, 3, 4])
def test_central_differences_of_the_exact_gradient_give_the_hessian(crit):
    """The Hessian recipe of `dynoed_hessian_batch` (central differences of the exact gradient, step
    cbrt(eps) * max(1, |x_i|), symmetrised) applied to the ORACLE's gradient: its (w, tau) block
    must equal the closed form d^2 phi(F + tau I)[B_a, B_b] (F is linear in w, augment.jl:223-226),
    which bounds the method's own error independently of the GPU."""
    o = Oracle(models.lotka_volterra_fishing(), crit, "rk4", 4)
    c = cpp_oracle.CppOracle("lv_fishing", o)
    rng = np.random.default_rng(11)
    x = rng.random(71); x[-1] = 0.3
    nv = x.size
    h = 6.0554544523933395e-06 * np.maximum(1.0, np.abs(x))
    P = np.vstack([x + np.diag(h), x - np.diag(h)])
    G = c.evaluate(P, grad=True)[1]
    den = (P[np.arange(nv), np.arange(nv)] - P[nv + np.arange(nv), np.arange(nv)])[:, None]
    A = (G[:nv] - G[nv:]) / den
    H = 0.5 * (A + A.T)
    units = np.tile(x, (60, 1)); units[:, 10:70] = np.eye(60)
    Bp = c.evaluate(units, grad=False)[2]
    _, g0, F0 = c.evaluate(x[None], grad=True)
    sym = lambda p: np.array([[p[0], p[1]], [p[1], p[2]]])
    dirs = [sym(b) for b in Bp] + [np.eye(2)]
    M = sym(F0[0]) + x[-1] * np.eye(2)
    Mi, d = np.linalg.inv(M), np.linalg.det(M)

    def second(A_, B_):
        tA, tB, tAB = np.trace(Mi @ A_), np.trace(Mi @ B_), np.trace(Mi @ A_ @ Mi @ B_)
        if crit == 1:
            return -d * (tA * tB - tAB)                      # -det M
        if crit == 3:
            return 2.0 * np.trace(Mi @ A_ @ Mi @ B_ @ Mi)    # tr M^-1
        return (tA * tB + tAB) / d                           # 1 / det M
    Hw = np.array([[second(a, b) for b in dirs] for a in dirs])
    cols = list(range(10, 71))
    scale = max(np.abs(Hw).max(), np.abs(g0).max())
    assert np.abs(H[np.ix_(cols, cols)] - Hw).max() < 2e-8 * scale
