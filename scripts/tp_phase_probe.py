"""Phase cycle counts of tp_eval_kernel from a DYNOED_TP_TIMING build (DYNOED_LIB=...): one LV design."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models
for alg in (D.RK4(4), D.Tsit5(4)):
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=alg).context()
    for n in (1, 143):
        des = np.random.default_rng(1).random((n, ctx.n_var)); des[:, -1] = 0.5
        d = torch.from_numpy(des).cuda(); c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
        S = torch.cuda.current_stream(); ctx.set_stream(S.cuda_stream)
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(S); ctx.eval_batch(d, c, g, None, async_=True); e1.record(S); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        cyc = g.cpu().numpy()[:, :7]
        print(os.environ.get("DYNOED_LIB", "default").split("/")[-1], type(alg).__name__, f"n={n} device {np.median(ts):.1f} us; cycles setup/phase1/phase1b/phase2/phase2b/phase3, ns start-to-end (median over designs):",
              np.median(cyc, axis=0).astype(int).tolist(), flush=True)
    ctx.close()
