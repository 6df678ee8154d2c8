"""Single-design latency through the host API (the Ipopt callback path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
for name, ctor, crit, alg in (("readme", models.readme, D.FisherACriterion, D.Tsit5(4)),
                              ("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.RK4(4)),
                              ("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.Tsit5(4))):
    prob = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=alg)
    ctx = prob.context()
    u = np.asarray(D.get_initial_variables(prob), dtype=np.float64)[None].copy()
    u[0, :-1] = 0.5
    for grad in (True, False):
        for _ in range(20): ctx.evaluate(u, grad=grad, fim=False)
        t = time.perf_counter()
        for _ in range(300): ctx.evaluate(u, grad=grad, fim=False)
        dt = (time.perf_counter() - t) / 300
        print(f"{name:10s} {type(alg).__name__:6s} n=1 {'obj+grad' if grad else 'obj only'}: {dt*1e6:7.1f} us per call", flush=True)
    opt = D.OptimizationProblem(prob)
    t = time.perf_counter()
    for k in range(200):
        u[0, 0] = 0.3 + 1e-3 * k
        opt.f.f(u[0]); opt.f.grad(None, u[0])
    print(f"{name:10s} optimizer callbacks f + grad: {(time.perf_counter()-t)/200*1e6:7.1f} us per iterate", flush=True)
    H = np.empty((1, ctx.n_var, ctx.n_var)); c1 = np.empty(1); g1 = np.empty((1, ctx.n_var))
    for _ in range(10): ctx.hessian_batch(u, H, c1, g1)
    t = time.perf_counter()
    for _ in range(200): ctx.hessian_batch(u, H, c1, g1)
    print(f"{name:10s} Hessian ({2 * ctx.n_var + 1} designs, one launch): {(time.perf_counter()-t)/200*1e6:7.1f} us per call", flush=True)
