"""Config 4 (stiff / DAE, Rosenbrock): device-resident throughput of the forward-mode gradient kernels and
their objective-only counterparts, for the library named by DYNOED_LIB.  One JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime

peak = runtime.fp64_peak()[0]
NB = 3
cases = [("chem6", models.chem6, D.FisherECriterion, [(0.5, 1.0), (0.1, 0.9)]), ("robertson", models.robertson, D.DCriterion, None)]
only = sys.argv[1:] or ["ROS2", "ROS3P", "RODAS3"]
for name, ctor, crit, lo in cases:
    for meth in (D.ROS2, D.ROS3P, D.RODAS3):
        if meth.__name__ not in only:
            continue
        ctx = D.OEDProblem(D.OEDSystem(ctor()), crit(), alg=meth(edges=D.graded_edges(10, 40, 1e-6, 10))).context()
        n = 65536
        ins = []
        for b in range(NB):
            d = torch.rand((n, ctx.n_var), dtype=torch.float64, device="cuda")
            d[:, -1] = 1e-3 + 0.9 * d[:, -1]
            for k, (a, s) in enumerate(lo or []):
                d[:, k] = a + s * d[:, k]
            ins.append(d)
        c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(ins[0])
        f = torch.empty((n, ctx.info.nF), dtype=torch.float64, device="cuda")
        row = {"lib": os.path.basename(runtime.LIB_PATH), "model": name, "method": meth.__name__, "n": n}
        for grad in (False, True):
            for k in range(2):
                ctx.eval_batch(ins[k % NB], c, g if grad else None, f, async_=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 6
            e0.record()
            for k in range(reps):
                ctx.eval_batch(ins[k % NB], c, g if grad else None, f, async_=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            ks = ctx.kernel_stats(grad)
            fl = ks["flops_per_design_grad" if grad else "flops_per_design_eval"]
            row["grad" if grad else "eval"] = {"ms": ms, "designs_per_s": n / ms * 1e3, "flops_per_design": fl,
                                               "frac_fp64": fl * n / ms / 1e9 / peak if fl else None,
                                               "regs": ks["regs_per_thread"], "ctas_per_sm": ks["blocks_per_sm"]}
        row["grad_over_eval"] = row["grad"]["ms"] / row["eval"]["ms"]
        print(json.dumps(row), flush=True)
        ctx.close()
