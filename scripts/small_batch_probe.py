"""Small gradient batches: direction-parallel kernel vs the adjoint kernel, host-call latency and
device-side duration, for n = 1 ... past the switch-over point."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
for name, ctor, crit, alg in (("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.RK4(4)),
                              ("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.Tsit5(4)),
                              ("lv_ic", models.lotka_volterra_ic, D.FisherDCriterion, D.RK4(4))):
    ctxs = {}
    for mode in ("dir", "adjoint"):
        if mode == "adjoint": os.environ["DYNOED_NO_DIRPAR"] = "1"
        else: os.environ.pop("DYNOED_NO_DIRPAR", None)
        os.environ["DYNOED_DIRPAR_MAX_LANES"] = "1000000"
        ctxs[mode] = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=alg).context()
    nv = ctxs["dir"].n_var
    nl = 148 * 4 * 32 // max(ctxs["dir"].kernel_stats(True)["small_batch_max"] * 0 + 1, 1)
    for n in (1, 8, 64, 143, 512, 1024, 1722, 3444, 6888):
        des = np.random.default_rng(1).random((n, nv)); des[:, -1] = 0.5
        d = torch.from_numpy(des).cuda(); c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
        row = f"{name:10s} {type(alg).__name__:5s} n={n:5d}"
        for mode, ctx in ctxs.items():
            S = torch.cuda.current_stream(); ctx.set_stream(S.cuda_stream)
            for _ in range(5): ctx.eval_batch(d, c, g, None, async_=True)
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(S); ctx.eval_batch(d, c, g, None, async_=True); e1.record(S); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            cc, gg = np.empty(n), np.empty((n, nv))
            for _ in range(5): ctx.eval_batch(des, cc, gg, None)
            t = time.perf_counter()
            for _ in range(50): ctx.eval_batch(des, cc, gg, None)
            host = (time.perf_counter() - t) / 50 * 1e6
            row += f" | {mode}: device {np.median(ts):7.1f} us, host call {host:7.1f} us"
        print(row, flush=True)
    for ctx in ctxs.values(): ctx.close()
