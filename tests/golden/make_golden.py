"""Generates tests/golden/*.npz — seeded designs and the numpy/sympy ORACLE's outputs (objective,
gradient, F) for every compiled-in model, both tableaus and all six criteria.

The reference (Julia) cannot run in this environment, so these are oracle outputs, not reference
outputs; the oracle itself is pinned against the reference's goldens in tests/test_oracle.py.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from dynamicoed.jl_b200 import models            # noqa: E402
from oracle.oed_oracle import Oracle              # noqa: E402

CASES = {"one_d": models.one_d, "readme": models.readme, "lv_fishing": models.lotka_volterra_fishing,
         "lv_ic": models.lotka_volterra_ic}


def designs_for(name, n_var, n, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    d = rng.random((n, n_var))
    d[:, -1] = 1e-3 + (1 - 1e-3) * d[:, -1]          # tau in [1e-3, 1)
    if name == "lv_ic":
        d[:, 0] = 0.1 + 0.9 * d[:, 0]                # x0 within its bounds (0.1, 1.0)
    d[0, :-1] = 0.0                                   # edge: all-zero design (F = 0)
    d[1, :-1] = 1.0                                   # edge: everything on
    if name == "lv_ic":
        d[0, 0], d[1, 0] = 0.1, 1.0
    return d


def main():
    for name, ctor in CASES.items():
        for method in ("rk4", "tsit5"):
            out = {}
            o = Oracle(ctor(), 0, method, 4)
            des = designs_for(name, o.n_var, 12, 20260)
            out["designs"] = des
            for crit in range(6):
                o.crit = crit
                c, g, F = o.evaluate(des, grad=True)
                out[f"crit{crit}"], out[f"grad{crit}"] = c, g
            out["F"] = F
            out["X0"] = o.trajectory(des[2])
            np.savez_compressed(os.path.join(HERE, f"{name}_{method}_m4.npz"), **out)
            print(name, method, des.shape)


STIFF = {"robertson": (models.robertson, 4, "ros3p"), "chem6": (models.chem6, 2, "rodas3")}


def stiff_edges():
    import dynamicoed.jl_b200 as D
    return D.graded_edges(10, 30, 1e-6, 6)


def main_stiff():
    """Stiff configuration: values from BOTH oracles (numpy flat-dense Rosenbrock, C++ nested duals),
    gradients from the C++ oracle; the two value sets must agree before anything is written."""
    from oracle.cpp_oracle import CppOracle
    for name, (ctor, crit, method) in STIFF.items():
        o = Oracle(ctor(), crit, method, list(stiff_edges()))
        rng = np.random.Generator(np.random.Philox(20264))
        des = rng.random((10, o.n_var))
        des[:, -1] = 1e-3 + 0.9 * des[:, -1]
        if name == "chem6":
            des[:, 0] = 0.5 + des[:, 0]
            des[:, 1] = 0.1 + 0.9 * des[:, 1]
        c_py, _, F_py = o.evaluate(des, grad=False)
        c, g, F = CppOracle(name, o).evaluate(des, grad=True)
        assert np.max(np.abs(F - F_py)) <= 1e-11 * np.max(np.abs(F)) and np.allclose(c, c_py, rtol=1e-9, atol=0)
        np.savez_compressed(os.path.join(HERE, f"{name}_{method}_graded.npz"), designs=des, crit=c, grad=g, F=F,
                            criterion=crit)
        print(name, method, des.shape)


if __name__ == "__main__":
    main()
    main_stiff()
