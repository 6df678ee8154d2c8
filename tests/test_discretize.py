"""Grid compiler against the reference's exact goldens: /root/reference/test/discretize.jl."""
import numpy as np

import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
from dynamicoed.jl_b200.discretize import (Timegrid, _generate_timegrid, design_layout,
                                           generate_initial_variables, generate_variable_bounds,
                                           julia_float_range)


def _prob(n_obs, control):
    s = D.structural_simplify(D.OEDSystem(models.mtk_simple(n_obs, control)))
    return s, Timegrid.from_system(s)


def test_julia_ranges_are_rationalised():
    # (0.0:0.1:3.0)[4] == 0.3 in Julia (TwicePrecision), not 0.30000000000000004
    r = julia_float_range(0.0, 0.1, 3.0)
    assert len(r) == 31 and r[3] == 0.3 and r[-1] == 3.0
    assert julia_float_range(0.0, 0.3, 3.0) == [float(k * 3) / 10 for k in range(11)]
    assert julia_float_range(0.0, 0.3, 2.0)[-1] == 1.8      # tail handled by the caller
    assert julia_float_range(0.0, 10.0, 100.0) == [10.0 * k for k in range(11)]
    assert all(a == b for a, b in zip(julia_float_range(0.0, 0.1, 2.0), np.arange(21) / 10))


def test_generate_timegrid_tail_and_int_rate():
    assert _generate_timegrid(0.3, (0.0, 2.0)) == [0.0, 0.3, 0.6, 0.9, 1.2, 1.5, 1.8, 2.0]
    assert _generate_timegrid(10, (0.0, 1.0)) == [k / 10 for k in range(11)]
    assert _generate_timegrid(2, (0.0, 2.0)) == [0.0, 1.0, 2.0]


def test_univariate():   # test/discretize.jl:48-57
    _, tg = _prob(1, False)
    assert tg.variables == ("w₁",)
    assert len(tg.timegrids) == 1
    assert tg.timegrids[0] == [(0.0, 1.0), (1.0, 2.0)]
    assert tg.timespans == [(0.0, 1.0), (1.0, 2.0)]
    assert tg.get_tspan(1) == (0.0, 1.0)
    assert tg.get_variable_idx("w₁", 1) == 1
    assert tg.get_variable_idx("w₁", 2) == 2


def test_two_observed():   # test/discretize.jl:59-90
    _, tg1 = _prob(1, False)
    _, tg = _prob(2, False)
    assert tg.variables == ("w₁", "w₂")
    assert tg.timegrids[0] == tg1.timegrids[0]
    assert tg.timegrids[1] == [(0.0, 0.3), (0.3, 0.6), (0.6, 0.9), (0.9, 1.2), (1.2, 1.5),
                               (1.5, 1.8), (1.8, 2.0)]
    assert len(tg.timespans) == 8
    assert tg.timespans == [(0.0, 0.3), (0.3, 0.6), (0.6, 0.9), (0.9, 1.0), (1.0, 1.2), (1.2, 1.5),
                            (1.5, 1.8), (1.8, 2.0)]
    assert tg.get_variable_idx("w₁", 4) == 1
    assert tg.get_variable_idx("w₂", 4) == 4
    assert tg.get_variable_idx("w₁", 5) == 2
    assert tg.get_variable_idx("w₂", 5) == 4


def test_two_observed_and_control():   # test/discretize.jl:92-134
    _, tg2 = _prob(2, False)
    _, tg = _prob(2, True)
    assert tg.variables == ("c", "w₁", "w₂")
    assert len(tg.timegrids) == 3
    j = julia_float_range
    assert tg.timegrids[0] == list(zip(j(0.0, 0.1, 1.9), j(0.1, 0.1, 2.0)))
    assert tg.timegrids[1] == tg2.timegrids[0]
    assert tg.timegrids[2] == tg2.timegrids[1]
    assert len(tg.timespans) == 20
    expect = [(0.0, 0.1), (0.1, 0.2), (0.2, 0.3), (0.3, 0.4), (0.4, 0.5), (0.5, 0.6), (0.6, 0.7),
              (0.7, 0.8), (0.8, 0.9), (0.9, 1.0), (1.0, 1.1), (1.1, 1.2), (1.2, 1.3), (1.3, 1.4),
              (1.4, 1.5), (1.5, 1.6), (1.6, 1.7), (1.7, 1.8), (1.8, 1.9), (1.9, 2.0)]
    assert tg.timespans == expect
    for var, i, k in [("w₁", 4, 1), ("w₁", 10, 1), ("w₁", 11, 2), ("w₂", 1, 1), ("w₂", 10, 4),
                      ("w₂", 15, 5), ("c", 1, 1), ("c", 10, 10), ("c", 15, 15)]:
        assert tg.get_variable_idx(var, i) == k


def test_design_layout_lv():
    """discretize.jl:110-137 + DynamicOED.jl:19-21: controls | ics | measurements | tau."""
    s = D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing()))
    tg = Timegrid.from_system(s)
    lay = design_layout(s, tg)
    assert lay.n_var == 71 and lay.control_offsets == [0] and lay.w_offsets == [10, 40]
    assert lay.tau_offset == 70 and len(tg.timespans) == 30
    assert np.array_equal(tg.indicators[0], np.repeat(np.arange(1, 11), 3))
    assert np.array_equal(tg.indicators[1], np.arange(1, 31))
    u0 = generate_initial_variables(s, tg)
    assert u0.regularization == 1.0 and np.all(np.asarray(u0)[:70] == 0.0)
    lb, ub = generate_variable_bounds(s, tg, True), generate_variable_bounds(s, tg, False)
    assert lb.regularization == np.finfo(float).eps and ub.regularization == 1.0
    assert np.all(np.asarray(ub)[:70] == 1.0) and np.all(np.asarray(lb)[:70] == 0.0)
    assert np.all(u0.controls.u == 0) and len(u0.measurements["w₁"]) == 30


def test_design_layout_ic():
    s = D.structural_simplify(D.OEDSystem(models.lotka_volterra_ic()))
    tg = Timegrid.from_system(s)
    lay = design_layout(s, tg)
    assert lay.n_var == 12 and lay.ic_offset == 0 and lay.w_offsets == [1] and lay.tau_offset == 11
    u0 = generate_initial_variables(s, tg)
    assert u0.initial_conditions["x₀"] == 0.49
    lb, ub = generate_variable_bounds(s, tg, True), generate_variable_bounds(s, tg, False)
    assert lb.initial_conditions["x₀"] == 0.1 and ub.initial_conditions["x₀"] == 1.0
