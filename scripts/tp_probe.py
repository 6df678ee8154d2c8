"""Smallest gradient batches: time-parallel kernel (one CTA per design) vs the direction-parallel kernel,
device-side duration and host-call latency, plus the largest gradient difference between the two."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
CASES = (("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.RK4(4)),
         ("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.Tsit5(4)),
         ("lv_ic", models.lotka_volterra_ic, D.FisherDCriterion, D.RK4(4)),
         ("lv_ic", models.lotka_volterra_ic, D.ACriterion, D.ROS3P(8)),
         ("lv_fishing", models.lotka_volterra_fishing, D.FisherDCriterion, D.RK4(16)),          # tables spill to global scratch
         ("chem6", models.chem6, D.FisherECriterion, D.ROS3P(edges=D.graded_edges(10, 30, 1e-6, 6))))
for name, ctor, crit, alg in CASES:
    ctxs = {}
    for mode in ("tp", "dir"):
        os.environ["DYNOED_TIMEPAR_MAX_N"] = "100000" if mode == "tp" else "0"
        ctxs[mode] = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=alg).context()
    os.environ.pop("DYNOED_TIMEPAR_MAX_N", None)
    nv = ctxs["tp"].n_var
    print(name, type(alg).__name__, "time_parallel_max", ctxs["tp"].kernel_stats(True)["time_parallel_max"],
          ctxs["dir"].kernel_stats(True)["time_parallel_max"], flush=True)
    for n in (1, 143):
        rng = np.random.default_rng(1)
        des = rng.random((n, nv)); des[:, -1] = 0.5
        if name == "chem6": des[:, :] = 0.3 + 0.4 * rng.random((n, nv))
        d = torch.from_numpy(des).cuda(); c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
        row = f"{name:10s} {type(alg).__name__:5s} n={n:5d}"
        res = {}
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
            res[mode] = (cc.copy(), gg.copy(), g.cpu().numpy().copy())
            row += f" | {mode}: device {np.median(ts):7.1f} us, host call {host:7.1f} us"
        sc = np.maximum(np.abs(res["dir"][1]).max(axis=1, keepdims=True), 1e-300)
        row += f" | max rel diff crit {np.abs(res['tp'][0] / res['dir'][0] - 1).max():.1e} grad rows {np.abs((res['tp'][1] - res['dir'][1]) / sc).max():.1e}"
        row += f" dev-ptr {np.abs((res['tp'][2] - res['dir'][2]) / sc).max():.1e}"
        print(row, flush=True)
    for ctx in ctxs.values(): ctx.close()
