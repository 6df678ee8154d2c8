"""Minimal driver for ncu: launches of the Rosenbrock kernels of BASELINE config 4 (chem6, ROS3P, FisherE):
objective-only launches first, then gradient launches (second-order sensitivity kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n = 65536
ctx = D.OEDProblem(D.OEDSystem(models.chem6()), D.FisherECriterion(), alg=D.ROS3P(edges=D.graded_edges(10, 40, 1e-6, 10))).context()
d = torch.rand((n, ctx.n_var), dtype=torch.float64, device="cuda"); d[:, -1] = 1e-3 + 0.9 * d[:, -1]
d[:, 0] = 0.5 + d[:, 0]; d[:, 1] = 0.1 + 0.9 * d[:, 1]
c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for grad in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(5):
        if it == 2: e0.record()
        ctx.eval_batch(d, c, g if grad else None, None, async_=True)
    e1.record(); torch.cuda.synchronize()
    print("chem6 ROS3P n", n, "grad", grad, "ms/launch", e0.elapsed_time(e1) / 3, "crit[0]", float(c[0]))
