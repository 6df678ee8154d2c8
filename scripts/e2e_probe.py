"""Host-buffer (e2e) path timing vs. pipeline granularity, plus raw pinned PCIe copy rates."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
n, nv = 65536, 71
hp = torch.from_numpy(np.random.default_rng(0).random((n, nv))).pin_memory()
hc = torch.empty(n, dtype=torch.float64).pin_memory(); hg = torch.empty((n, nv), dtype=torch.float64).pin_memory()
hf = torch.empty((n, 3), dtype=torch.float64).pin_memory()
d = torch.empty((n, nv), dtype=torch.float64, device="cuda"); d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): hg.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 20
    print("pinned H2D" + (" + concurrent D2H" if both else ""), f"{n*nv*8/dt/1e9:.1f} GB/s per direction, {dt*1e3:.3f} ms")
for chunks in sys.argv[1:] or ["8"]:
    os.environ["DYNOED_HOST_CHUNKS"] = chunks
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4)).context()
    for fim in (hf, None):
        for _ in range(3): ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), None if fim is None else fim.numpy())
        t = time.perf_counter()
        for _ in range(30): ctx.eval_batch(hp.numpy(), hc.numpy(), hg.numpy(), None if fim is None else fim.numpy())
        dt = (time.perf_counter() - t) / 30
        print(f"chunks {chunks:>3s} fim {fim is not None} {dt*1e3:.3f} ms/step {n/dt/1e6:.1f} Mdesigns/s", flush=True)
    ctx.close()
