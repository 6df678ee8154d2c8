import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime
print("fp64 peak", runtime.fp64_peak())
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
ctx = prob.context()
print(ctx.kernel_stats(True)); print(ctx.kernel_stats(False))
n = 65536
rng = np.random.Generator(np.random.Philox(20262))
des = rng.random((n, 71)); des[:, -1] = 1e-3 + (1-1e-3)*des[:, -1]
d = torch.from_numpy(des).cuda()
crit = torch.empty(n, dtype=torch.float64, device='cuda'); grad = torch.empty_like(d); fim = torch.empty((n,3), dtype=torch.float64, device='cuda')
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for with_grad in (True, False):
    for it in range(3):
        ctx.eval_batch(d, crit, grad if with_grad else None, fim, async_=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        ctx.eval_batch(d, crit, grad if with_grad else None, fim, async_=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    ks = ctx.kernel_stats(with_grad)
    fl = ks['flops_per_design_grad' if with_grad else 'flops_per_design_eval']
    print("grad" if with_grad else "eval", "ms", ms, "designs/s", n/ms*1e3, "TFLOP/s", fl*n/ms*1e-9)
# host path
c, gg, f = ctx.evaluate(des[:4096])
t = time.time(); c, gg, f = ctx.evaluate(des); print("host e2e pageable s", time.time()-t)
print("fp64 peak", runtime.fp64_peak())
