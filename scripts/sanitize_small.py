"""Small, fast tour of every kernel for compute-sanitizer memcheck (partial tiles, both pointer
kinds, all integrators)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

rng = np.random.default_rng(0)
def designs(ctx, n):
    d = rng.random((n, ctx.n_var)); d[:, -1] = 0.01 + 0.9 * d[:, -1]; return d
for alg in (D.RK4(2), D.Tsit5(2)):
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherECriterion(), alg=alg).context()
    des = designs(ctx, 300)
    c, g, F = ctx.evaluate(des)                                   # host path, partial tile
    c1, g1, _ = ctx.evaluate(des[:1])
    c2, g2, _ = ctx.evaluate(des, flags=runtime.GRAD_NO_CONTROLS)
    d = torch.from_numpy(des).cuda(); dc = torch.empty(300, dtype=torch.float64, device="cuda"); dg = torch.empty_like(d)
    ctx.eval_batch(d, dc, dg, None); ctx.eval_batch(d, dc, None, None)
    ks = ctx.kernel_stats(True); n = ks["grid"] * ks["block"] + 77  # > one wave: persistent loop + overlap
    big = torch.rand((n, 71), dtype=torch.float64, device="cuda")
    outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty_like(big)) for _ in range(2)]
    for k in range(3): ctx.eval_batch(big, outs[k % 2][0], outs[k % 2][1], None, async_=True)
    ctx.sync(); print("argmin", ctx.argmin())
    X = ctx.trajectory(des[:3]); Fg, P, Pi = ctx.information_gain(des[:5])
    b = ctx.fim_basis(des[0]); cc = np.empty(300); gg = np.empty((300, 71)); ctx.eval_fixed_controls(b, des, cc, gg, None)
    hist = torch.empty((2, 2), dtype=torch.int64, device="cuda"); ctx.argmin_history_async(2, hist)
    out = torch.empty((2, 3), dtype=torch.int64, device="cuda"); ctx.argmin_gathered(torch.stack([hist, hist]).contiguous(), 2, 300, out, groups=2)
    ctx.sync(); ctx.close()
for name, ctor, crit in (("robertson", models.robertson, D.DCriterion), ("chem6", models.chem6, D.FisherECriterion)):
    ctx = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=D.ROS3P(edges=D.graded_edges(10, 12, 1e-6, 3))).context()
    des = designs(ctx, 200)
    if name == "chem6": des[:, 0] += 0.5; des[:, 1] = 0.1 + 0.9 * des[:, 1]
    c, g, F = ctx.evaluate(des); c0, _, _ = ctx.evaluate(des, grad=False)
    print(name, float(c[0]), float(c0[0])); ctx.close()
v, L = runtime.criterion_batch(5, rng.random((10, 6)) + np.array([3, 0, 3, 0, 0, 3.0]), tau=0.1, with_gradient=True)
print("done", v[:2])
