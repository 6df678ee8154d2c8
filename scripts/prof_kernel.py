"""Minimal driver for ncu: a few device-resident launches of the LV-fishing evaluation kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
grad = (sys.argv[2] != "nograd") if len(sys.argv) > 2 else True
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
ctx = prob.context()
rng = np.random.Generator(np.random.Philox(20262))
h = rng.random((n, 71)); h[:, -1] = 1e-3 + (1 - 1e-3) * h[:, -1]
d = torch.from_numpy(h).cuda()
c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
f = torch.empty((n, 3), dtype=torch.float64, device="cuda")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    if it == 3: e0.record()
    ctx.eval_batch(d, c, g if grad else None, f, async_=True)
e1.record(); torch.cuda.synchronize()
print("n", n, "grad", grad, "ms/launch", e0.elapsed_time(e1) / 3, "crit[0]", float(c[0]))
