"""Host-buffer path driven from 1, 2 and 3 host threads (one ctx per thread, blocking calls)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n, nv = 65536, 71
def mk():
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
    hp = torch.from_numpy(np.random.default_rng(0).random((n, nv))).pin_memory()
    hc = torch.empty(n, dtype=torch.float64).pin_memory(); hg = torch.empty((n, nv), dtype=torch.float64).pin_memory()
    hf = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    return ctx, hp, hc, hg, hf
for T in (1, 2, 3):
    sets = [mk() for _ in range(T)]
    for ctx, hp, hc, hg, hf in sets:
        for _ in range(3): ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
    steps = 60
    def work(i):
        ctx, hp, hc, hg, hf = sets[i]
        for k in range(i, steps, T):
            ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
            _ = float(hc[0])
    ths = [threading.Thread(target=work, args=(i,)) for i in range(T)]
    t = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = (time.perf_counter() - t) / steps
    print(f"threads {T}: {dt*1e3:.3f} ms/step {n/dt/1e6:.1f} Mdesigns/s", flush=True)
    for s in sets: s[0].close()
# single host thread, pipelined host calls (two in flight)
ctx, _, _, _, _ = mk()
bufs = [mk()[1:] for _ in range(2)]
ids = [None, None]
for k in range(4):
    hp, hc, hg, hf = bufs[k % 2]
    if ids[k % 2] is not None: ctx.wait(ids[k % 2])
    ids[k % 2] = ctx.eval_batch_host_async(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
for i in ids: ctx.wait(i)
ids = [None, None]; steps = 60
t = time.perf_counter()
for k in range(steps):
    hp, hc, hg, hf = bufs[k % 2]
    if ids[k % 2] is not None:
        ctx.wait(ids[k % 2]); _ = float(hc[0])
    ids[k % 2] = ctx.eval_batch_host_async(hp.numpy(), hc.numpy(), hg.numpy(), hf.numpy())
for i in ids: ctx.wait(i)
dt = (time.perf_counter() - t) / steps
print(f"1 thread, pipelined host calls: {dt*1e3:.3f} ms/step {n/dt/1e6:.1f} Mdesigns/s", flush=True)
