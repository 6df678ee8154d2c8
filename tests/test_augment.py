"""Symbolic augmentation against /root/reference/test/mtk_extensions.jl (exact, symbolic)."""
import sympy as sp

import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
from dynamicoed.jl_b200.symbolic import (getdefault, is_fisher_state, is_information_gain,
                                         is_measurement_function)


def test_mtk_extensions_simple_system():
    base = models.mtk_simple(1, False)
    x, = base.states
    p, = base.ps
    oed = D.OEDSystem(base)
    st = oed.get_states()
    assert [is_fisher_state(s) for s in st] == [False, False, True, False]      # :26
    assert [is_information_gain(s) for s in st] == [False, False, False, True]  # :27
    assert len(st) == 4
    ps = oed.parameters()
    assert [is_measurement_function(q) for q in ps] == [False, True]            # :31
    eqs = oed.equations()
    assert len(eqs) == 4 and len(oed.observed) == len(base.observed)
    G, F, Q = st[1], st[2], st[3]
    w = ps[1]
    dx, dG, dF = eqs[0].lhs, None, eqs[2].lhs
    # groundtruth (:43-46): D(x) ~ p x ; 0 ~ x - D(G) + p G ; D(F) ~ G^2 w ; Q ~ G
    assert sp.simplify(eqs[0].rhs - p * x) == 0 and eqs[0].lhs.of == x
    assert eqs[1].lhs == 0
    dGsym = [s for s in eqs[1].rhs.free_symbols if getattr(s, "role", "") == "derivative"]
    assert len(dGsym) == 1 and dGsym[0].of == G
    assert sp.simplify(eqs[1].rhs - (x - dGsym[0] + p * G)) == 0
    assert eqs[2].lhs.of == F and sp.simplify(eqs[2].rhs - G ** 2 * w) == 0
    assert eqs[3].lhs == Q and sp.simplify(eqs[3].rhs - G) == 0
    red = D.structural_simplify(oed)
    assert len(red.equations()) == 3                                            # :54
    assert red.get_states() == st[:-1]                                          # :55
    # defaults (:60-66)
    assert getdefault(x) == 2.0 and getdefault(p) == -2.0 and getdefault(F) == 0.0
    assert getdefault(G) == 0.0 and getdefault(w) == 0.0
    assert red.u0 == [2.0, 0.0, 0.0]
    assert sp.simplify(red.dG[0, 0] - (x + p * G)) == 0


def test_lv_augmented_structure():
    s = D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing()))
    assert [v.name for v in s.states] == ["x", "y", "G[1,1]", "G[2,1]", "G[1,2]", "G[2,2]",
                                          "F[1,1]", "F[1,2]", "F[2,2]"]
    assert [p.name for p in s.p_tunable] == ["c[1]", "c[2]"]
    assert [c.name for c in s.controls] == ["u"] and [w.name for w in s.w] == ["w₁", "w₂"]
    x, y = s.x
    G = s.G
    w1, w2 = s.w
    # Appendix B of SURVEY.md / augment.jl:223-226
    assert sp.expand(s.dF[0] - (w1 * G[0, 0] ** 2 + w2 * G[1, 0] ** 2)) == 0
    assert sp.expand(s.dF[1] - (w1 * G[0, 0] * G[0, 1] + w2 * G[1, 0] * G[1, 1])) == 0
    assert sp.expand(s.dF[2] - (w1 * G[0, 1] ** 2 + w2 * G[1, 1] ** 2)) == 0
    th = {p.name: p for p in s.ps}
    J = sp.Matrix(s.f).jacobian([x, y])
    fp = sp.Matrix([[-x * y, 0], [0, x * y]])
    assert (J * G + fp - s.dG).applyfunc(sp.expand) == sp.zeros(2, 2)
    assert sp.expand(s.f[0] - (th["p[1]"] * x - th["c[1]"] * x * y - th["p[3]"] * x * th["u"])) == 0


def test_unknown_ic_G0_rule():
    """augment.jl:182-188: state i's unit entry goes to column np-(i-1)."""
    aug = D.OEDSystem(models.lotka_volterra_ic())
    assert [p.name for p in aug.p_tunable] == ["c[1]", "x₀"]
    assert aug.unknown_initial_conditions == [1]
    assert aug.G_init == sp.Matrix([[0, 1], [0, 0]])
    s = D.structural_simplify(aug)
    assert s.u0 == [0.49, 0.7, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0]


def test_robertson_tearing():
    """Rober.jl:9-23 after dae_index_lowering + structural_simplify: states [y1,y2,G1,G2,F]."""
    s = D.structural_simplify(D.dae_index_lowering(D.OEDSystem(models.robertson())))
    assert [v.name for v in s.states] == ["y₁", "y₂", "G[1,1]", "G[2,1]", "F[1,1]"]
    y1, y2 = s.x
    k = {p.name: p for p in s.ps}
    y3 = 1 - y1 - y2
    f1 = -k["k₁"] * y1 + k["k₃"] * y2 * y3
    f2 = k["k₁"] * y1 - k["k₃"] * y2 * y3 - k["k₂"] * y2 ** 2
    assert sp.expand(s.f[0] - f1) == 0 and sp.expand(s.f[1] - f2) == 0
    J = sp.Matrix([f1, f2]).jacobian([y1, y2])
    fp = sp.Matrix([f1, f2]).jacobian([k["k₁"]])
    assert (J * s.G + fp - s.dG).applyfunc(sp.expand) == sp.zeros(2, 1)
    assert sp.expand(s.dF[0] - s.w[0] * s.G[0, 0] ** 2) == 0
    names = {e.lhs.name: e.rhs for e in s.observed}
    assert sp.expand(names["G[3,1]"] + s.G[0, 0] + s.G[1, 0]) == 0
