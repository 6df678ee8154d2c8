"""Minimal driver for ncu: the forward-only gradient kernel (DYNOED_GRAD_NO_CONTROLS) and the Tsit5 adjoint kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime
n = 65536
d = torch.rand((n, 71), dtype=torch.float64, device="cuda"); d[:, -1] = 0.01 + 0.9 * d[:, -1]
c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
for name, alg, flags in (("rk4 forward-only", D.RK4(4), runtime.GRAD_NO_CONTROLS), ("tsit5 adjoint", D.Tsit5(4), 0)):
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=alg).context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        if it == 3: e0.record()
        ctx.eval_batch(d, c, g, None, async_=True, flags=flags)
    e1.record(); torch.cuda.synchronize()
    print(name, "n", n, "ms/launch", e0.elapsed_time(e1) / 3, "crit[0]", float(c[0]))
    ctx.close()
