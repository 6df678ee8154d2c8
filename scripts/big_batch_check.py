"""1,048,576 designs through the host path and the device path: same numbers, sane memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n = 1 << 20
ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
rng = np.random.Generator(np.random.Philox(1)); des = rng.random((n, 71)); des[:, -1] = 1e-3 + 0.99 * des[:, -1]
c, g, F = ctx.evaluate(des)        # first call allocates slots / staging
t = time.perf_counter(); c, g, F = ctx.evaluate(des); th = time.perf_counter() - t
d = torch.from_numpy(des).cuda(); dc = torch.empty(n, dtype=torch.float64, device="cuda"); dg = torch.empty_like(d)
dF = torch.empty((n, 3), dtype=torch.float64, device="cuda")
t = time.perf_counter(); ctx.eval_batch(d, dc, dg, dF); td = time.perf_counter() - t
print("host path %.1f ms (pageable), device path %.2f ms" % (th * 1e3, td * 1e3))
print("equal:", np.array_equal(dc.cpu().numpy(), c), np.array_equal(dg.cpu().numpy(), g), np.array_equal(dF.cpu().numpy(), F))
print("argmin", ctx.argmin(), int(np.argmin(c)), "scratch MB", ctx.kernel_stats(True)["scratch_bytes"] / 1e6,
      "GPU mem used MB", torch.cuda.mem_get_info()[1] / 1e6 - torch.cuda.mem_get_info()[0] / 1e6)
hp = torch.from_numpy(des).pin_memory(); hc = torch.empty(n, dtype=torch.float64).pin_memory()
hg = torch.empty((n, 71), dtype=torch.float64).pin_memory(); hf = torch.empty((n, 3), dtype=torch.float64).pin_memory()
ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
t = time.perf_counter()
for _ in range(3): ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
print("host path pinned %.1f ms; equal: %s" % ((time.perf_counter() - t) / 3 * 1e3, np.array_equal(hg.numpy(), g)))
