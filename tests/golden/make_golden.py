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


if __name__ == "__main__":
    main()
