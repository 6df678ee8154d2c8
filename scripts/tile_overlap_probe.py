"""Back-to-back (overlapping) launches of the gradient kernel vs tile size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime
n, nv, NB = 65536, 71, 4
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
lay = prob.layout
ins = [torch.rand((n, nv), dtype=torch.float64, device="cuda") for _ in range(NB)]
outs = [(torch.empty(n, dtype=torch.float64, device="cuda"), torch.empty((n, nv), dtype=torch.float64, device="cuda")) for _ in range(NB)]
for tile in (128, 96, 64, 32):
    ctx = runtime.Context("lv_fishing", criterion=1, method=0, tgrid=prob.timegrid.tpoints,
                          indicators=prob.timegrid.indicators - 1, ctrl_offsets=lay.control_offsets,
                          w_offsets=lay.w_offsets, ic_offset=lay.ic_offset, tau_offset=lay.tau_offset,
                          n_var=lay.n_var, substeps=4, tile=tile)
    S = torch.cuda.current_stream(); ctx.set_stream(S.cuda_stream)
    for k in range(6): ctx.eval_batch(ins[k % NB], outs[k % NB][0], outs[k % NB][1], None, async_=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record(S)
    for k in range(K): ctx.eval_batch(ins[k % NB], outs[k % NB][0], outs[k % NB][1], None, async_=True)
    e1.record(S); torch.cuda.synchronize()
    ks = ctx.kernel_stats(True)
    print(f"tile {tile:3d} ctas/SM {ks['blocks_per_sm']} {e0.elapsed_time(e1)/K*1e3:7.1f} us/step overlapped {ks['overlapped_launches']}", flush=True)
    ctx.close()
