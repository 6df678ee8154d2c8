"""Quick device-resident timing of the LV-fishing kernel (+ FP64 operand-pattern microbenchmarks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

if "--micro" in sys.argv:
    pk = runtime.fp64_peak()[0]
    print("fp64 peak TF", pk)
    for v, name in enumerate(["DFMA r*c+c", "DFMA r*r+r (3 regs)", "DFMA r*r+same (2 regs)", "DMUL r*r", "DADD r+r", "DFMA r*U+r", "DFMA B*r+r (B shared)", "DFMA r*C+r (C of 2 regs)", "DMMA m8n8k4 (FMA equiv.)", "DMMA + DFMA mix (FMA equiv.)"]):
        g = runtime.fp64_microbench(v)
        print(f"variant {v} {name:28s} {g:9.1f} Ginstr/s  = {g/ (pk*1e3/2):.3f} of DFMA-chain rate")
method = D.Tsit5(4) if "--tsit5" in sys.argv else D.RK4(4)
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=method)
ctx = prob.context()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for n in (56832, 65536, 8 * 56832):
    rng = np.random.Generator(np.random.Philox(20262))
    h = rng.random((n, 71)); h[:, -1] = 1e-3 + (1 - 1e-3) * h[:, -1]
    d = torch.from_numpy(h).cuda()
    c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
    f = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    for grad in (True, False):
        for _ in range(3): ctx.eval_batch(d, c, g if grad else None, f, async_=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ctx.eval_batch(d, c, g if grad else None, f, async_=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ks = ctx.kernel_stats(grad)
        fl = ks["flops_per_design_grad" if grad else "flops_per_design_eval"]
        print(f"n {n:7d} {'grad' if grad else 'eval'} {ms*1e3:8.1f} us {n/ms/1e3:8.2f} Mdesigns/s {fl*n/ms/1e9:6.2f} TF regs {ks['regs_per_thread']} bps {ks['blocks_per_sm']} crit0 {float(c[0]):.15g}", flush=True)
