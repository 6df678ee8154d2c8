"""Sub-metrics of SURVEY.md 8d on one GPU: substeps m = 1 / 4 / 16, criterion-only, Tsit5 tableau;
the stiff configuration's Rosenbrock kernels.  Writes profiles/<tag>_variants.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
peak = runtime.fp64_peak()[0]
rows = []
NB = 4


def timed(ctx, n, n_var, grad, nF, reps=30, lo=None):
    ins = []
    for b in range(NB):
        d = torch.rand((n, n_var), dtype=torch.float64, device="cuda")
        d[:, -1] = 1e-3 + 0.9 * d[:, -1]
        if lo is not None:
            for k, (a, s) in enumerate(lo): d[:, k] = a + s * d[:, k]
        ins.append(d)
    outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty((n, n_var), dtype=torch.float64, device="cuda"),
             torch.empty((n, nF), dtype=torch.float64, device="cuda")) for _ in range(NB)]
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for k in range(4): ctx.eval_batch(ins[k % NB], outs[k % NB][0], outs[k % NB][1] if grad else None, outs[k % NB][2], async_=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps): ctx.eval_batch(ins[k % NB], outs[k % NB][0], outs[k % NB][1] if grad else None, outs[k % NB][2], async_=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = 65536
for alg_name, alg in (("RK4 m=1", D.RK4(1)), ("RK4 m=4", D.RK4(4)), ("RK4 m=16", D.RK4(16)), ("Tsit5 m=4", D.Tsit5(4))):
    ctx = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=alg).context()
    for grad in (True, False):
        ms = timed(ctx, n, 71, grad, 3, reps=30 if "16" not in alg_name else 8)
        ks = ctx.kernel_stats(grad)
        fl = ks["flops_per_design_grad" if grad else "flops_per_design_eval"]
        rows.append({"model": "lv_fishing", "integrator": alg_name, "what": "criterion+gradient" if grad else "criterion only",
                     "n": n, "ms": ms, "designs_per_s": n / ms * 1e3, "tflops": fl * n / ms / 1e9,
                     "frac_fp64_peak": fl * n / ms / 1e9 / peak, "regs": ks["regs_per_thread"], "ctas_per_sm": ks["blocks_per_sm"]})
        print(rows[-1], flush=True)
    ctx.close()
for name, ctor, crit, lo in (("robertson", models.robertson, D.DCriterion, None),
                             ("chem6", models.chem6, D.FisherECriterion, [(0.5, 1.0), (0.1, 0.9)])):
    for meth in (D.ROS2, D.ROS3P):
        edges = D.graded_edges(10, 40, 1e-6, 10)
        prob = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=meth(edges=edges))
        ctx = prob.context()
        nn = 16384
        for grad in (True, False):
            ms = timed(ctx, nn, ctx.n_var, grad, ctx.info.nF, reps=5, lo=lo)
            ks = ctx.kernel_stats(grad)
            rows.append({"model": name, "integrator": meth.__name__ + " graded 130 steps", "what": "criterion+gradient" if grad else "criterion only",
                         "n": nn, "ms": ms, "designs_per_s": nn / ms * 1e3, "regs": ks["regs_per_thread"], "ctas_per_sm": ks["blocks_per_sm"]})
            print(rows[-1], flush=True)
        ctx.close()
json.dump({"fp64_peak_tflops": peak, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"{tag}_variants.json"), "w"), indent=1)
