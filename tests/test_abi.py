"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/dynoed.h declares, its model registry agrees with the emitter, the generated sources are
up to date, and — without a GPU — compute entry points fail loudly instead of falling back."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dynoed.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dynoed_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_library):
    lib = C.CDLL(built_library)
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dynoed.h but not exported"
    assert set(names) == set(runtime.EXPORTS)


def test_abi_version_and_registry(built_library):
    L = runtime.lib()
    assert L.dynoed_abi_version() == 1
    names = runtime.model_names()
    assert names == list(models.BUILTIN.keys())
    infos = {i["name"]: i for i in json.load(open(os.path.join(
        ROOT, "dynamicoed.jl_b200", "csrc", "generated", "models.json")))}
    for n in names:
        mi = runtime.model_info(n)
        e = infos[n]
        assert (mi.nx, mi.np, mi.n_obs, mi.n_ctrl, mi.n_ic, mi.n_theta, mi.nz, mi.nF) == \
            (e["nx"], e["np"], e["n_obs"], e["n_ctrl"], e["n_ic"], e["n_theta"], e["nz"], e["nF"])
        assert list(mi.x0[:mi.nx]) == e["x0"]
        assert list(mi.G0[:mi.nx * mi.np]) == e["G0"]
        assert list(mi.theta[:mi.n_theta]) == e["theta"]
        assert mi.flops_primal == e["flops"]["primal"] and mi.flops_adjoint == e["flops"]["adjoint"]
    lv = runtime.model_info("lv_fishing")
    assert (lv.nx, lv.np, lv.n_obs, lv.n_ctrl, lv.nz, lv.nF) == (2, 2, 2, 1, 6, 3)


def test_unknown_model_and_bad_desc_are_errors(built_library):
    mi = runtime.ModelInfo()
    assert runtime.lib().dynoed_model_query(b"nope", C.byref(mi)) == -3
    assert b"unknown model" in runtime.lib().dynoed_last_error(None)
    d = runtime.ProblemDesc()
    d.struct_size = 4            # ABI mismatch must be rejected before anything else
    h = C.c_void_p()
    assert runtime.lib().dynoed_create(C.byref(d), 0, C.byref(h)) == -1


def test_generated_sources_are_up_to_date(tmp_path):
    """csrc/generated is exactly what the emitter prints for the reference's models."""
    from dynamicoed.jl_b200.emit import emit_builtin
    emit_builtin(str(tmp_path))
    gen = os.path.join(ROOT, "dynamicoed.jl_b200", "csrc", "generated")
    for f in os.listdir(tmp_path):
        assert open(os.path.join(gen, f)).read() == open(os.path.join(tmp_path, f)).read(), f


def test_builtin_model_resolution():
    from dynamicoed.jl_b200.problem import _resolve_model
    s = D.structural_simplify(D.OEDSystem(models.lotka_volterra_fishing(integer_control=True)))
    libpath, name, info = _resolve_model(s)
    assert libpath is None and name == "lv_fishing" and info["nz"] == 6


def test_no_cpu_fallback(built_library):
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    prob = D.OEDProblem(D.OEDSystem(models.one_d()), D.FisherACriterion())
    with pytest.raises(runtime.DynoedError):
        prob.context()
    with pytest.raises(runtime.DynoedError):
        D.FisherACriterion()(np.eye(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamicoed.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"(from|import)\s+oracle", r"liboed_oracle", r"oed_oracle", r"oracle/"):
                    assert not re.search(pat, src), (pat, os.path.join(dirpath, f))


def test_header_is_valid_c_and_matches_the_ctypes_mirror(tmp_path):
    """include/dynoed.h must compile as plain C (the Julia `ccall` / cgo-style consumers see a C
    ABI), and the struct sizes the ctypes mirror assumes must equal the compiler's."""
    import ctypes
    import shutil
    import subprocess
    from dynamicoed.jl_b200 import runtime
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "t.c"
    src.write_text('#include "dynoed.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu\\n",'
                   'sizeof(dynoed_model_info), sizeof(dynoed_problem_desc), sizeof(dynoed_kernel_stats));return 0;}\n')
    exe = tmp_path / "t"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(runtime.ModelInfo), ctypes.sizeof(runtime.ProblemDesc),
                     ctypes.sizeof(runtime.KernelStats)]
