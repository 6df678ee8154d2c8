"""Minimal driver for ncu: the direction-parallel small-batch kernel, one LV design and 143 (a Hessian)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for n in (1, 143):
    d = torch.rand((n, 71), dtype=torch.float64, device="cuda"); d[:, -1] = 0.01 + 0.9 * d[:, -1]
    c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        if it == 3: e0.record()
        ctx.eval_batch(d, c, g, None, async_=True)
    e1.record(); torch.cuda.synchronize()
    print("dir kernel n", n, "us/launch", e0.elapsed_time(e1) / 3 * 1e3, "crit[0]", float(c[0]), flush=True)
