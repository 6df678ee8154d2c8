"""e2e (host-buffer) rate of blocking dynoed_eval_batch calls on pinned buffers, 1 and 2 host threads, for the current
DYNOED_ZC_GRAD setting (gradient tiles stored by the kernel straight into page-locked host memory vs a D2H copy
stage), with a bitwise check of the gradient against the device-pointer path.  One JSON line."""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n, nv = 65536, 71
def mk():
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
    rng = np.random.default_rng(1)
    h = rng.random((n, nv)); h[:, -1] = 1e-3 + 0.99 * h[:, -1]
    return (ctx, torch.from_numpy(h).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
            torch.empty((n, nv), dtype=torch.float64).pin_memory(), torch.empty((n, 3), dtype=torch.float64).pin_memory())
sets = [mk(), mk()]
ctx, hp, hc, hg, hf = sets[0]
dd = hp.cuda(); dc = torch.empty(n, dtype=torch.float64, device="cuda"); dg = torch.empty_like(dd); df = torch.empty((n, 3), dtype=torch.float64, device="cuda")
ctx.eval_batch(dd, dc, dg, df); ctx.sync()
hg.zero_(); ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
out = {"zc_grad": os.environ.get("DYNOED_ZC_GRAD", "0"), "bitwise_equal_to_device_path": bool(torch.equal(hg, dg.cpu()) and torch.equal(hc, dc.cpu()) and torch.equal(hf, df.cpu()))}
steps = 60
def run(nt):
    def work(i):
        c, p, cc, g, f = sets[i]
        for _ in range(i, steps, nt):
            c.eval_batch(p.numpy(), cc.numpy(), g.numpy(), f.numpy())
    for i in range(nt):
        for _ in range(2): sets[i][0].eval_batch(sets[i][1].numpy(), sets[i][2].numpy(), sets[i][3].numpy(), sets[i][4].numpy())
    ths = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
    t = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    torch.cuda.synchronize()
    return n * steps / (time.perf_counter() - t)
out["one_thread"] = run(1); out["two_threads"] = run(2)
out["one_thread_b"] = run(1); out["two_threads_b"] = run(2)
print(json.dumps(out))
