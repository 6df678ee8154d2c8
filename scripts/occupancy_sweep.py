"""Throughput of the evaluation kernel vs. tile size and resident CTAs per SM (full waves only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicoed.jl_b200 import models, runtime
import dynamicoed.jl_b200 as D

def run(tile, cap, n=None, reps=20, grad=True):
    if cap: os.environ["DYNOED_MAX_CTAS_PER_SM"] = str(cap)
    else: os.environ.pop("DYNOED_MAX_CTAS_PER_SM", None)
    prob = D.OEDProblem(D.OEDSystem(models.lotka_volterra_fishing()), D.FisherDCriterion(), alg=D.RK4(4))
    lay = prob.layout
    ctx = runtime.Context("lv_fishing", criterion=1, method=0, tgrid=prob.timegrid.tpoints,
                          indicators=prob.timegrid.indicators - 1, ctrl_offsets=lay.control_offsets,
                          w_offsets=lay.w_offsets, ic_offset=lay.ic_offset, tau_offset=lay.tau_offset,
                          n_var=lay.n_var, substeps=4, tile=tile)
    ks = ctx.kernel_stats(grad)
    bps = ks["blocks_per_sm"]
    n = n or 148 * bps * tile
    rng = np.random.Generator(np.random.Philox(1))
    d = torch.from_numpy(rng.random((n, 71))).cuda()
    c = torch.empty(n, dtype=torch.float64, device="cuda"); g = torch.empty_like(d)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_profiling(True)
    for _ in range(3): ctx.eval_batch(d, c, g if grad else None, None, async_=True)
    torch.cuda.synchronize(); ctx.set_profiling(True)
    for _ in range(reps): ctx.eval_batch(d, c, g if grad else None, None, async_=True)
    torch.cuda.synchronize()
    ms, cnt = ctx.kernel_time()
    fl = ks["flops_per_design_grad" if grad else "flops_per_design_eval"]
    print(f"tile {tile:4d} cap {cap} bps {bps:2d} warps/SM {bps*tile//32:3d} regs {ks['regs_per_thread']} n {n:6d} "
          f"kernel {ms*1e3:8.1f} us  {n/ms/1e3:8.2f} Mdesigns/s  {fl*n/ms/1e9:6.2f} TF", flush=True)
    ctx.close()

if __name__ == "__main__":
    print("fp64 peak", runtime.fp64_peak())
    for tile in (128, 64, 32):
        for cap in ([1, 2, 0] if tile == 128 else [1, 2, 4, 0] if tile == 64 else [1, 2, 4, 7, 8, 0]):
            run(tile, cap)
    run(128, 0, 65536); run(64, 0, 65536); run(32, 0, 65536); run(32, 7, 65536); run(64, 0, 131072); run(64, 0, 1048576, reps=5)
    for tile in (128, 64):
        run(tile, 0, grad=False); run(tile, 0, 65536, grad=False)
