"""Minimal driver for ncu: a few launches of the Rosenbrock kernel (Robertson, ROS3P, gradient)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n = 65536
ctx = D.OEDProblem(D.OEDSystem(models.robertson()), D.DCriterion(), alg=D.ROS3P(edges=D.graded_edges(10, 40, 1e-6, 10))).context()
d = torch.rand((n, 11), dtype=torch.float64, device="cuda"); d[:, -1] = 0.01 + 0.9 * d[:, -1]
c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    if it == 3: e0.record()
    ctx.eval_batch(d, c, g, None, async_=True)
e1.record(); torch.cuda.synchronize()
print("robertson ROS3P n", n, "ms/launch", e0.elapsed_time(e1) / 3, "crit[0]", float(c[0]))
