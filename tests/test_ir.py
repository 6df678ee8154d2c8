"""The model interchange format (dynoed-ir/1): every built-in model survives
system -> JSON -> system with identical emitted CUDA code, grids and layout.  This is what a Julia
host hands over (julia/export_ir.jl writes the same JSON from a ModelingToolkit ODESystem)."""
import json

import numpy as np
import pytest

import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import ir, models
from dynamicoed.jl_b200.emit import emit_model


@pytest.mark.parametrize("name", list(models.BUILTIN) + ["mtk_simple"])
def test_roundtrip_gives_identical_kernel_source_and_layout(name):
    ctor = models.BUILTIN.get(name) or (lambda: models.mtk_simple(n_obs=2, control=True))
    sys0 = ctor()
    text = json.dumps(ir.system_to_ir(sys0), ensure_ascii=False)
    sys1 = ir.system_from_ir(json.loads(text))
    s0 = D.structural_simplify(D.OEDSystem(sys0))
    s1 = D.structural_simplify(D.OEDSystem(sys1))
    src0, info0 = emit_model("M", s0)
    src1, info1 = emit_model("M", s1)
    assert src0 == src1
    for k in ("x0", "G0", "theta", "ic_state", "nx", "np", "n_obs", "n_ctrl", "n_ic"):
        assert info0[k] == info1[k]
    p0 = D.OEDProblem(s0, D.FisherACriterion())
    p1 = D.OEDProblem(s1, D.FisherACriterion())
    assert np.array_equal(p0.timegrid.tpoints, p1.timegrid.tpoints)
    assert np.array_equal(p0.timegrid.indicators, p1.timegrid.indicators)
    assert np.array_equal(np.asarray(D.get_initial_variables(p0)), np.asarray(D.get_initial_variables(p1)))
    l0, l1 = p0.layout, p1.layout
    assert (l0.n_var, l0.control_offsets, l0.w_offsets, l0.ic_offset, l0.tau_offset) == \
           (l1.n_var, l1.control_offsets, l1.w_offsets, l1.ic_offset, l1.tau_offset)


def test_ir_of_lv_fishing_reads_like_the_reference_declaration():
    d = ir.system_to_ir(models.lotka_volterra_fishing())
    assert d["format"] == "dynoed-ir/1" and d["tspan"] == [0.0, 3.0]
    assert [s["name"] for s in d["states"]] == ["x", "y"]
    u = [p for p in d["parameters"] if p["input"]]
    assert len(u) == 1 and u[0]["measurement_rate"] == 0.3 and u[0]["bounds"] == [0.0, 1.0]
    assert [p["name"] for p in d["parameters"] if p["tunable"] and not p["input"]] == ["c[1]", "c[2]"]
    assert d["equations"][0]["lhs"] == "D(x0)" and [o["measurement_rate"] for o in d["observed"]] == [0.1, 0.1]


def test_build_from_ir_resolves_builtin_models_without_recompiling(tmp_path, built_library):
    """`python -m dynamicoed.jl_b200.ir build model.json` — the Julia host's hand-over: an exported
    built-in model maps back onto libdynoed_cuda.so (identical emitted source), no nvcc run."""
    from dynamicoed.jl_b200 import runtime
    path = tmp_path / "lv.json"
    path.write_text(json.dumps(ir.system_to_ir(models.lotka_volterra_fishing()), ensure_ascii=False))
    out = ir.build_from_ir(str(path))
    assert out["model"] == "lv_fishing" and out["library"] == runtime.LIB_PATH
    assert (out["nx"], out["np"], out["n_obs"], out["n_ctrl"], out["n_ic"]) == (2, 2, 2, 1, 0)


@pytest.mark.parametrize("name", list(models.BUILTIN))
def test_augmented_section_is_checked_against_the_hosts_own_augmentation(name):
    """The "augmented" section carries the reference's own `full_equations(structural_simplify(OEDSystem(sys)))`
    (julia/export_ir.jl); `build_from_ir` emits a kernel only if this host's augmentation reproduces every
    right-hand side.  Here the section is produced by the host itself, re-ordered and re-written in another
    expression form (expanded / factored, `^` for powers as Julia prints them), and then tampered with."""
    import sympy as sp
    simp = D.structural_simplify(D.OEDSystem(models.BUILTIN[name]()))
    sec = ir.augmented_to_ir(simp)
    assert len(sec["equations"]) == len(simp.states)
    other = {"equations": [dict(state=e["state"], rhs=str(sp.factor(sp.sympify(e["rhs"]))).replace("**", "^"))
                           for e in reversed(sec["equations"])]}
    assert ir.check_augmented(json.loads(json.dumps(other)), simp) == len(simp.states)
    bad = json.loads(json.dumps(sec))
    bad["equations"][-1]["rhs"] = "1.0000001*(" + bad["equations"][-1]["rhs"] + ")"
    with pytest.raises(ValueError, match="differs from the reference"):
        ir.check_augmented(bad, simp)
    short = {"equations": sec["equations"][:-1]}
    with pytest.raises(ValueError, match="differential states"):
        ir.check_augmented(short, simp)


def test_build_from_ir_runs_the_cross_check(tmp_path, built_library):
    from dynamicoed.jl_b200 import runtime
    sys0 = models.lotka_volterra_fishing()
    d = ir.system_to_ir(sys0)
    d["augmented"] = ir.augmented_to_ir(D.structural_simplify(D.OEDSystem(sys0)))
    path = tmp_path / "lv_aug.json"
    path.write_text(json.dumps(d, ensure_ascii=False))
    out = ir.build_from_ir(str(path))
    assert out["model"] == "lv_fishing" and out["augmented_equations_checked"] == 9
    d["augmented"]["equations"][2]["rhs"] += " + 1e-6*x0"
    path.write_text(json.dumps(d, ensure_ascii=False))
    with pytest.raises(ValueError):
        ir.build_from_ir(str(path))
