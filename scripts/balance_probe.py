"""Isolated-launch duration of the gradient kernel vs (tile, resident CTAs per SM): wave quantisation.
A lone launch of n designs runs ceil(n / (grid * tile)) rounds; the last round is mostly empty when
the grid fills every warp slot.  Fewer resident warps per SM with evenly filled rounds should be faster."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dynamicoed.jl_b200 as D
from dynamicoed.jl_b200 import models, runtime
nv, NB = 71, 3
prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
lay = prob.layout
NMAX = 131072
ins = [torch.rand((NMAX, nv), dtype=torch.float64, device="cuda") for _ in range(NB)]
outs = [(torch.empty(NMAX, dtype=torch.float64, device="cuda"), torch.empty((NMAX, nv), dtype=torch.float64, device="cuda")) for _ in range(NB)]
combos = [(128, 0), (128, 2), (64, 0), (64, 5), (64, 4), (64, 3), (32, 0), (32, 10), (32, 9), (32, 8), (32, 7), (32, 6)]
for n in (65536, 131072, 32768, 100000):
    for tile, cap in combos:
        if cap: os.environ["DYNOED_MAX_CTAS_PER_SM"] = str(cap)
        else: os.environ.pop("DYNOED_MAX_CTAS_PER_SM", None)
        ctx = runtime.Context("lv_fishing", criterion=1, method=0, tgrid=prob.timegrid.tpoints,
                              indicators=prob.timegrid.indicators - 1, ctrl_offsets=lay.control_offsets,
                              w_offsets=lay.w_offsets, ic_offset=lay.ic_offset, tau_offset=lay.tau_offset,
                              n_var=lay.n_var, substeps=4, tile=tile)
        S = torch.cuda.current_stream(); ctx.set_stream(S.cuda_stream)
        for k in range(3): ctx.eval_batch(ins[k % NB][:n], outs[k % NB][0][:n], outs[k % NB][1][:n], None, async_=True)
        torch.cuda.synchronize()
        ts = []
        for k in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(S)
            ctx.eval_batch(ins[k % NB][:n], outs[k % NB][0][:n], outs[k % NB][1][:n], None, async_=True)
            e1.record(S); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        e0.record(S)
        for k in range(K): ctx.eval_batch(ins[k % NB][:n], outs[k % NB][0][:n], outs[k % NB][1][:n], None, async_=True)
        e1.record(S); torch.cuda.synchronize()
        ks = ctx.kernel_stats(True)
        print(f"n {n:6d} tile {tile:3d} ctas/SM {ks['blocks_per_sm']:2d} warps/SM {ks['blocks_per_sm'] * tile // 32:2d}  isolated {np.median(ts):7.1f} us   back-to-back {e0.elapsed_time(e1)/K*1e3:7.1f} us (overlapped {ks['overlapped_launches']})", flush=True)
        ctx.close()
